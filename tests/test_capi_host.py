"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/maelstrom_b200.h declares, its struct layouts match the header
(compiled with gcc) and the Python mirror; pure host helpers agree with the
oracle and the reference's doc vectors.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "maelstrom_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as G
    G.build()
    from maelstrom_b200 import _lib
    return _lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    L = C.CDLL(lib.SO_PATH)
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libmaelstrom_b200.so does not export %s" % n
    # and the Python binding covers exactly the header
    assert sorted(lib.SYMBOLS) == names


def test_struct_layouts_match_header(lib, tmp_path):
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "maelstrom_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(ms_msg), sizeof(ms_body), sizeof(ms_event), sizeof(ms_jbody),
         sizeof(ms_op), sizeof(ms_config));
  printf("%zu %zu %zu %zu %zu\n", offsetof(ms_msg, src), offsetof(ms_msg, type), offsetof(ms_msg, p1),
         offsetof(ms_op, body), offsetof(ms_config, p_loss));
  printf("%zu %zu %zu\n", offsetof(ms_config, max_endpoints), offsetof(ms_config, device), offsetof(ms_config, reserved));
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = list(map(int, out))
    assert sizes[:5] == [48, 24, 32, 32, 40]
    assert sizes[5] == C.sizeof(lib.Config)
    assert sizes[6:11] == [16, 32, 40, 16, lib.Config.p_loss.offset]
    assert sizes[11:] == [lib.Config.max_endpoints.offset, lib.Config.device.offset, lib.Config.reserved.offset]
    assert lib.MSG_DTYPE.itemsize == 48 and lib.EVENT_DTYPE.itemsize == 32
    assert lib.JBODY_DTYPE.itemsize == 32 and lib.OP_DTYPE.itemsize == 40
    assert lib.MSG_DTYPE.fields["p1"][1] == 40 and lib.OP_DTYPE.fields["body"][1] == 16


def test_abi_version_and_loud_failure_without_gpu(lib):
    L = lib.lib()
    assert L.ms_abi_version() == 2
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    import maelstrom_b200 as mb
    with pytest.raises(mb.SimError) as e:
        mb.Sim(5)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("topo", ["grid", "line", "total", "tree2", "tree3", "tree4"])
@pytest.mark.parametrize("n", [1, 2, 5, 7, 25, 26, 64, 100])
def test_engine_topology_tables_match_oracle(lib, topo, n):
    from maelstrom_b200.engine import topology
    for k in range(n):
        assert topology(topo, n, k) == O.topology(topo, n, k)


def test_engine_topology_doc_vector(lib):
    from maelstrom_b200.engine import topology
    want = {0: [3, 1], 1: [4, 2, 0], 2: [1], 3: [0, 4], 4: [1, 3]}   # doc/03-broadcast/01-broadcast.md:302-306
    for k, nb in want.items():
        assert topology("grid", 5, k) == nb
    assert sum(len(topology("grid", 4096, k)) for k in range(4096)) == 16128


def test_type_codes_agree_between_header_engine_and_oracle(lib):
    from maelstrom_b200.engine import TYPES
    src = open(HEADER).read()
    for name, code in TYPES.items():
        m = re.search(r"MS_T_%s\s*=\s*(\d+)" % name.upper(), src)
        assert m and int(m.group(1)) == code == O.T[name]


def test_journal_decoders_refuse_malformed_batches():
    # host-only entry points: no device involved
    import ctypes as C
    import numpy as np
    from maelstrom_b200 import _lib
    from maelstrom_b200._lib import EVENT_DTYPE, JBatch
    L = _lib.lib()
    b = JBatch()
    b.first_event, b.n_events, b.n_rounds, b.format = 0, 1, 0, 8          # an event but no round row
    rounds = np.zeros(1, dtype=_lib.JROUND_DTYPE)
    events = np.zeros(1, dtype="<u8")
    out = np.zeros(1, dtype=EVENT_DTYPE)
    assert L.ms_journal_decode(C.byref(b), rounds.ctypes.data, events.ctypes.data, out.ctypes.data) < 0
    b.n_rounds, b.format = 1, 4                                          # MS_JFMT_4 needs the stream's history
    assert L.ms_journal_decode(C.byref(b), rounds.ctypes.data, events.ctypes.data, out.ctypes.data) < 0
    d = L.ms_jdecoder_create(8)
    assert d
    e4 = np.array([0x80000000 | 5], dtype="<u4")                        # a :recv of a message the decoder never saw sent
    rounds[0]["id_ref"] = 100
    assert L.ms_jdecoder_decode(d, C.byref(b), rounds.ctypes.data, e4.ctypes.data, out.ctypes.data) < 0
    assert b"window" in L.ms_jdecoder_error(d)
    e4[0] = (3 << 16) | 7                                               # a :send n3 -> n7: its id is the round's first id
    assert L.ms_jdecoder_decode(d, C.byref(b), rounds.ctypes.data, e4.ctypes.data, out.ctypes.data) == 0
    assert (int(out[0]["msg_id"]), int(out[0]["src"]), int(out[0]["dest"])) == (100, 3, 7)
    L.ms_jdecoder_destroy(d)
    assert L.ms_jdecoder_create(2) is None                              # window too small
