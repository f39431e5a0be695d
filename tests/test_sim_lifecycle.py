"""Lifecycle entry points of the C ABI against the reference calls they replace: j/journal ..
j/close! (net.clj:128-137), add-node! / remove-node! (net.clj:139-152), process/start-node! /
stop-node! (process.clj:168-256).  [emul] = kernel sources on the CPU SIMT emulator; [cuda] = B200."""
import numpy as np
import pytest

from scenarios import both, make_pair, random_broadcast_ops

pytestmark = pytest.mark.usefixtures("engine_backend")


def test_journal_file_lifecycle(tmp_path):
    # j/journal ... j/close! (net.clj:128-137): events stream to a file while the run goes on
    # (header "MSJ1", level, record sizes; then event [+ body] records in event-id order)
    import maelstrom_b200 as mb
    from maelstrom_b200._lib import EVENT_DTYPE, JBODY_DTYPE
    n = 25
    g, o = make_pair(n, topology="grid", n_values=512, ring_cap=1024, max_window=512, journal_cap_log2=14)
    path = str(tmp_path / "net-messages.msj")
    g.journal_open(path)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(2)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=8, per_tick=30, seed=9)
        s.schedule(ops)
        s.run(10_000_000)

    both(g, o, scenario)
    g.journal_close()
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw[:16], dtype="<u4")
    assert hdr.tolist() == [0x314A534D, 2, 32, 32]
    rec = np.frombuffer(raw[16:], dtype=np.dtype([("ev", EVENT_DTYPE), ("body", JBODY_DTYPE)]))
    ev_o, bd_o = o.journal()
    assert len(rec) == len(ev_o)
    for f in ("event_id", "time_ns", "msg_id", "src", "dest"):
        assert np.array_equal(rec["ev"][f], ev_o[f]), f
    for f in ("type", "flags", "msg_id", "in_reply_to", "p0", "p1"):
        assert np.array_equal(rec["body"][f], bd_o[f]), f
    assert np.array_equal(rec["body"]["id"], rec["ev"]["msg_id"])
    ev_left, _ = g.drain()
    assert len(ev_left) == 0                                  # everything went to the file


def test_endpoint_and_node_lifecycle():
    # add-node! / remove-node! (net.clj:139-152), process/start-node! / stop-node! (process.clj:168-256)
    import maelstrom_b200 as mb
    g, o = make_pair(4, topology="line", n_values=16)
    L, h = g.L, g.h
    c = g.add_endpoint("c0")
    assert g.endpoint_index("c0") == c and g.endpoint_index("n3") == 3
    assert g.endpoint_index("nobody") == -1                     # node-not-found, code 1 upstream (net.clj:159-164)
    assert L.ms_start_nodes(h, 1) == 0 and L.ms_start_nodes(h, 0) == -2      # the workload is fixed at ms_create
    g.send(c, 1, mb.body("broadcast", msg_id=1, p0=5))
    r = g.recv(c, 1_000_000_000)
    assert int(r["type"]) == mb.TYPES["broadcast_ok"]
    g.run(3_000_000)
    assert g.journal_written() == 2 * (2 + 3) == len(g.drain()[0])          # 2 client msgs + 3 gossip sends, x 2 events
    g.remove_endpoint(c)
    assert g.endpoint_index("c0") == -1
    assert g.send(c, 1, mb.body("read", msg_id=2)) == -1        # "Invalid source for message" (net.clj:172-173)
    assert L.ms_remove_endpoint(h, c) == -1                     # already gone: "No such node in network"
    c2 = g.add_endpoint("c1")
    assert c2 == c                                              # the slot of a removed client is recycled (lowest index first)
    assert g.endpoint_index("c1") == c and g.endpoint_index("c0") == -1
    oc = o.add_endpoint("c0"); o.remove_endpoint(oc)
    assert o.add_endpoint("c1") == oc == c                      # same rule in the oracle
    assert L.ms_stop_nodes(h) == 0
    assert g.send(c2, 0, mb.body("read", msg_id=1)) == -1       # "Invalid dest for message" (net.clj:174-175)


@pytest.mark.parametrize("name", ["flood_grid25", "echo_12_ops", "latency_loss_partition"])
def test_engine_reproduces_committed_journals(name):
    # tests/golden/journals.json, generated from the oracle (tests/golden/make_golden.py)
    import golden_cases as G
    assert name in G.CORE_CASES
    G.check_engine_against_fixture(name)
