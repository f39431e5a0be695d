"""The journal in the reference's on-disk format (net/journal.clj:55-127): ms_journal_open on a
path ending in ".fressian" writes `Event{id time type message}` Fressian objects.  No JVM exists here,
so the bytes are pinned two ways: (1) a fixture decoded by hand from journal.clj's write handlers and
the published Fressian encoding, (2) a reader written independently of the writer (fressian_reader.py)
that must give back the journal the oracle produced.  The reference's own reader has not seen these
files: "parity unpinned" for the file format."""
import numpy as np
import pytest

import oracle_lib as O
from fressian_reader import FressianReader, Keyword
from scenarios import make_pair

pytestmark = pytest.mark.usefixtures("engine_backend")

# Event 0 of the scenario below: c0 -> n1 {type "broadcast", msg_id 1, message 5}, journaled at time 0.
# By hand from journal.clj:70-92 + write-body! (:55-68):
HAND = bytes([
    0xEF, 0xDC]) + b"ev" + bytes([0x04,        # STRUCTTYPE "ev" (packed string, length 2) with 4 components
    0x00,                                      # id 0
    0x00,                                      # time 0
    0xCD, 0xCA, 0xF7, 0xCD, 0xDE]) + b"send" + bytes([   # :send -> PUT_CACHE, "key", ns nil, PUT_CACHE "send"
    0xEF, 0xDD]) + b"msg" + bytes([0x04,       # STRUCTTYPE "msg" with 4 components
    0x00,                                      # message id 0
    0xCD, 0xDC]) + b"c0" + bytes([             # src, cached
    0xCD, 0xDC]) + b"n1" + bytes([             # dest, cached
    0xC0, 0xED,                                # tag "map", BEGIN_CLOSED_LIST
    0xCD, 0xCA, 0xF7, 0xCD, 0xDE]) + b"type" + bytes([0xCD, 0xE3, 0x09]) + b"broadcast" + bytes([   # :type "broadcast" (value cached)
    0xCD, 0xCA, 0xF7, 0xCD, 0xE0]) + b"msg_id" + bytes([0x01,
    0xCD, 0xCA, 0xF7, 0xCD, 0xE1]) + b"message" + bytes([0x05,
    0xFD])                                     # END_COLLECTION


def test_fressian_journal_file(tmp_path):
    g, o = make_pair(4, topology="line", n_values=16)
    path = str(tmp_path / "0.fressian")

    def scenario(s, body):
        c = s.add_endpoint("c0")
        if hasattr(s, "journal_open"):
            s.journal_open(path)
        s.send(c, 1, body("broadcast", msg_id=1, p0=5))
        s.run(3_000_000)
        s.send(c, 2, body("read", msg_id=2))
        s.run(5_000_000)
        s.send(c, 0, body("add", msg_id=70000, p0=300))     # unknown to a broadcast node: error 10; wide ints
        s.run(7_000_000)
        if hasattr(s, "journal_close"):
            s.journal_close()

    scenario(g, __import__("maelstrom_b200").body)
    scenario(o, O.body)
    data = open(path, "rb").read()
    assert data[:len(HAND)] == HAND, (data[:len(HAND)].hex(), HAND.hex())
    events = FressianReader(data).read_all()
    ev_o, bd_o = o.journal()
    assert len(events) == len(ev_o) > 10
    names = {i: "n%d" % i for i in range(4)}
    names[4] = "c0"
    type_names = {v: k for k, v in O.T.items()}
    for e, eo, bo in zip(events, ev_o, bd_o):
        recv = bool(int(eo["event_id"]) >> 63)
        assert e["id"] == int(eo["event_id"]) & ((1 << 63) - 1) and e["time"] == int(eo["time_ns"])
        assert isinstance(e["type"], Keyword) and e["type"] == ("recv" if recv else "send")
        m = e["message"]
        assert m["id"] == int(eo["msg_id"]) and m["src"] == names[int(eo["src"])] and m["dest"] == names[int(eo["dest"])]
        body = m["body"]
        assert all(isinstance(k, Keyword) for k in body)
        assert body["type"] == type_names[int(bo["type"])]
        if int(bo["flags"]) & O.F_MSG_ID:
            assert body["msg_id"] == int(bo["msg_id"])
        if int(bo["flags"]) & O.F_REPLY:
            assert body["in_reply_to"] == int(bo["in_reply_to"])
        if body["type"] == "broadcast":
            assert body["message"] == int(bo["p0"])
        if body["type"] == "error":
            assert body["code"] == int(bo["p0"]) == 10
    assert any(m["message"]["body"].get("msg_id") == 70000 for m in events)     # a 3-byte packed int survived


def test_fressian_int_forms():
    # FressianWriter.writeInt: every packed width, both signs, through the writer's own file output
    import ctypes as C
    import subprocess, os, tempfile, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    vals = [0, 1, 63, 64, -1, -2, -4096, 4095, 4096, -4097, 524287, 524288, -524289, (1 << 25) - 1, 1 << 25, -(1 << 25) - 1,
            (1 << 33) - 1, 1 << 33, (1 << 41) - 1, 1 << 41, (1 << 49) - 1, 1 << 49, -(1 << 49) - 1, (1 << 62), -(1 << 63)]
    src = textwrap.dedent("""
        #include "%s/maelstrom_b200/csrc/ms_fressian.h"
        int main(int argc, char** argv) {
          FILE* f = fopen(argv[1], "wb");
          msf::Writer w(f);
          static const long long v[] = {%s};
          for (long long x : v) w.write_int(x);
          fclose(f);
        }""") % (root, ", ".join("%dLL" % v if v != -(1 << 63) else "(-9223372036854775807LL - 1)" for v in vals))
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cpp"), "w").write(src)
        subprocess.check_call(["g++", "-std=c++17", "-o", os.path.join(d, "t"), os.path.join(d, "t.cpp")])
        subprocess.check_call([os.path.join(d, "t"), os.path.join(d, "out")])
        data = open(os.path.join(d, "out"), "rb").read()
    r = FressianReader(data)
    assert [r.read_int() for _ in vals] == vals
    # widths: 1 byte up to 63, 2 bytes up to +-4096, 3 up to +-2^19, 4 up to +-2^25, 5 / 6 / 7, then INT + 8
    assert data[:4] == bytes([0x00, 0x01, 0x3F, 0x50]) and data[4] == 0x40
