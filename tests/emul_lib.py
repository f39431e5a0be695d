"""TEST INFRASTRUCTURE: builds the engine's own kernel + host sources (maelstrom_b200/csrc) with
g++ against tests/native/emul (a CPU SIMT emulator: fibers for threads, synchronous streams) and
binds the result with the same ctypes table the product uses.  This lets the CPU suite run the
real kernel logic against the oracle.  The product never loads this library."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "maelstrom_b200", "csrc")
EMUL = os.path.join(HERE, "native", "emul")
OUT_DIR = os.path.join(HERE, "native", "_build")
OUT = os.path.join(OUT_DIR, "libmaelstrom_b200_emul.so")
SOURCES = [os.path.join(CSRC, "ms_kernels.cu"), os.path.join(CSRC, "ms_engine.cu"),
           os.path.join(EMUL, "simt.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, "ms_device.cuh"), os.path.join(CSRC, "ms_raft.cuh"), os.path.join(CSRC, "ms_tree.h"),
                  os.path.join(CSRC, "ms_json.h"), os.path.join(CSRC, "ms_fressian.h"), os.path.join(EMUL, "cuda_runtime.h"),
                  os.path.join(ROOT, "include", "maelstrom_b200.h")]
_lib = None


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    tmp = OUT + ".tmp.%d" % os.getpid()
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DMS_EMUL", "-I" + EMUL,
           "-I" + os.path.join(ROOT, "include")]
    for s in SOURCES:
        cmd += ["-x", "c++", s]
    cmd += ["-o", tmp, "-lpthread"]
    subprocess.check_call(cmd)
    os.replace(tmp, OUT)
    return OUT


def load():
    global _lib
    if _lib is None:
        from maelstrom_b200._lib import SYMBOLS
        # MS_EMUL_LIB: a prebuilt variant of the same library (tools/emul_asan.sh: ASan + UBSan)
        L = C.CDLL(os.environ.get("MS_EMUL_LIB") or build())
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class use:
    """Context manager: maelstrom_b200.Sim objects created inside run on the emulator."""

    def __enter__(self):
        from maelstrom_b200 import _lib as product
        self.product = product
        self.saved = product._lib
        product._lib = load()
        return product._lib

    def __exit__(self, *exc):
        self.product._lib = self.saved
        return False
