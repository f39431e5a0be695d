"""The reference-side binding a maintainer adds on a box with a JDK:
maelstrom_b200/csrc/ms_jni.c, integration/java/maelstrom/b200/Native.java and
integration/clojure/maelstrom/net.clj.  No JVM exists here, so these tests check what can be checked
without one: the JNI unit compiles to an empty object without <jni.h>, type-checks against a minimal
JNI header stand-in, covers every entry point of include/maelstrom_b200.h, and the record offsets
the Clojure side writes match the C structs."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JNI_C = os.path.join(ROOT, "maelstrom_b200", "csrc", "ms_jni.c")
HEADER = os.path.join(ROOT, "include", "maelstrom_b200.h")
NATIVE = os.path.join(ROOT, "integration", "java", "maelstrom", "b200", "Native.java")
NET_CLJ = os.path.join(ROOT, "integration", "clojure", "maelstrom", "net.clj")


def _cc(extra, out):
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fPIC", "-c", JNI_C, "-I" + os.path.join(ROOT, "include"),
           "-o", out] + extra
    subprocess.check_call(cmd)


def test_jni_unit_is_empty_without_a_jdk(tmp_path):
    obj = str(tmp_path / "ms_jni_empty.o")
    _cc([], obj)
    syms = subprocess.check_output(["nm", obj], text=True)
    assert "Java_maelstrom_b200_Native" not in syms


def test_jni_unit_type_checks_and_covers_the_abi(tmp_path):
    obj = str(tmp_path / "ms_jni.o")
    _cc(["-I" + os.path.join(ROOT, "tests", "native", "jni_stub")], obj)
    syms = subprocess.check_output(["nm", obj], text=True)
    natives = set(re.findall(r" T Java_maelstrom_b200_Native_(\w+)", syms))
    called = set(re.findall(r" U (ms_\w+)", syms))
    declared = set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", open(HEADER).read())) - {"ms_barrier_fn", "ms_journal_sink"}
    assert declared <= called, "ABI entries without a JNI native: %s" % sorted(declared - called)
    java = set(re.findall(r"public static native \S+ (\w+)\(", open(NATIVE).read()))
    assert natives == java, (sorted(natives - java), sorted(java - natives))


def test_clojure_namespace_mirrors_maelstrom_net():
    src = open(NET_CLJ).read()
    assert src.startswith("(ns maelstrom.net")
    for fn in ("net", "jepsen-net", "jepsen-os", "add-node!", "remove-node!", "send!", "recv!"):   # net.clj's public fns
        assert re.search(r"\(defn %s[\s\n]" % re.escape(fn), src), fn
    used = set(re.findall(r"Native/(\w+)", src))
    java = set(re.findall(r"public static native \S+ (\w+)\(", open(NATIVE).read()))
    assert used <= java, sorted(used - java)
    # parentheses balance (the file cannot be loaded here)
    depth, in_str, esc, in_comment = 0, False, False, False
    for ch in src:
        if in_comment:
            in_comment = ch != "\n"
        elif in_str:
            if esc:
                esc = False
            elif ch == "\\":
                esc = True
            elif ch == '"':
                in_str = False
        elif ch == ";":
            in_comment = True
        elif ch == '"':
            in_str = True
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
            assert depth >= 0
    assert depth == 0 and not in_str


def test_record_offsets_written_by_the_clojure_side():
    from maelstrom_b200._lib import Config, MSG_DTYPE, EVENT_DTYPE, JBODY_DTYPE
    assert C.sizeof(Config) == 136                                   # (direct 136) in ms-config
    off = {n: getattr(Config, n).offset for n, _ in Config._fields_}
    assert (off["n_nodes"], off["workload"], off["topology"], off["latency_dist"], off["latency_mean_ms"]) == (0, 4, 8, 12, 16)
    assert (off["seed_lo"], off["seed_hi"], off["p_loss"], off["n_values"], off["journal_level"]) == (20, 24, 32, 40, 64)
    m = MSG_DTYPE.fields
    assert [m[k][1] for k in ("id", "src", "dest", "msg_id", "in_reply_to", "type", "flags", "p0", "p1")] == \
        [0, 16, 20, 24, 28, 32, 34, 36, 40]                          # decode-message
    e, b = EVENT_DTYPE.fields, JBODY_DTYPE.fields
    assert [e[k][1] for k in ("event_id", "time_ns", "msg_id", "src", "dest")] == [0, 8, 16, 24, 28]
    assert b["type"][1] == 16                                        # drain-journal!
