"""txn-list-append on a persistent hash tree (demo/ruby/datomic_list_append.rb, MS_W_TXN_TREE): immutable tree
nodes in lww-kv, the root pointer in lin-kv.  Known answers for the tree arithmetic (csrc/ms_tree.h), then the
message-level behaviour of the CUDA engine against the oracle, journal bit for bit.  No Ruby exists in this image, so
neither side is pinned to the executed reference: what can be pinned (Zlib.crc32 key placement, range splitting as
:170-186 computes it) is pinned here."""
import zlib

import numpy as np
import pytest

import oracle_lib as O
from scenarios import assert_same_journal, both, make_pair, ops_array


def test_key_hash_is_zlib_crc32_of_the_decimal_string_mod_128():
    L = O.lib()
    for k in list(range(0, 300)) + [999, 1000, 4095, 9999, 16383]:
        assert L.or_tree_key_hash(k) == zlib.crc32(str(k).encode()) % 128, k      # datomic_list_append.rb:59-61


def pack_ops(ops):
    """[(f, key)] with f in {"r", "append"} -> the 64-bit payload of a txn (include/maelstrom_b200.h)"""
    w = 0
    for i, (f, k) in enumerate(ops):
        assert i < 4 and 0 <= k < 16384
        w |= (0x8000 | (0x4000 if f == "append" else 0) | k) << (16 * i)
    return w


def txn_ops(rng, n, clients, t0_ms, n_ticks, per_tick, n_keys, mids):
    rows = np.zeros(n_ticks * per_tick, dtype=O.OP_DTYPE)
    i = 0
    for t in range(n_ticks):
        for _ in range(per_tick):
            c = int(rng.integers(len(clients)))
            mids[c] += 1
            ops = [(("append", "r")[int(rng.integers(3)) == 0], int(rng.integers(n_keys))) for _ in range(int(rng.integers(1, 5)))]
            r = rows[i]
            i += 1
            r["time_ns"] = (t0_ms + t) * 1_000_000
            r["src"] = clients[c]
            r["dest"] = int(rng.integers(n))
            r["body"]["type"] = O.T["txn"]
            r["body"]["flags"] = O.F_MSG_ID
            r["body"]["msg_id"] = mids[c]
            r["body"]["p0"] = i
            r["body"]["p1"] = pack_ops(ops)
    return rows


@pytest.mark.parametrize("n,dist,mean,seed", [(1, "constant", 0, 1), (5, "constant", 0, 2), (7, "uniform", 2, 3), (12, "constant", 1, 4)])
def test_txn_tree_vs_oracle(engine_backend, n, dist, mean, seed):
    g, o = make_pair(n, workload="txn-list-append-tree", latency_dist=dist, latency_mean_ms=mean, max_endpoints=n + 16,
                     ring_cap=1024, max_window=512, server_ring_cap=256, server_max_window=64, rpc_table=256,
                     tree_ptrs=4096, journal_cap_log2=20, calendar_slots=32, calendar_cap=4096, seed=seed * 977)

    def scenario(s, body):
        s.add_endpoint("lin-kv", O.KIND_SERVICE)
        s.add_endpoint("lww-kv", O.KIND_SERVICE)
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(4)]
        s.schedule(ops_array([(0, cs[i % 4], i, "init", 1 + i // 4, 0) for i in range(n)]))
        rng = np.random.default_rng(seed)
        s.schedule(txn_ops(rng, n, cs, 40, 60, 3, 40, [100] * 4))
        s.run((400 + 1300 * mean) * 1_000_000)
        return s.client_replies()

    rg, ro = both(g, o, scenario)
    assert rg == ro == n + 180                                     # every init and every txn was answered
    ev, bd = assert_same_journal(g, o)
    sends = (ev["event_id"] >> np.uint64(63)) == 0
    oks = bd[(bd["type"] == O.T["txn_ok"]) & sends]
    errs = bd[(bd["type"] == O.T["error"]) & sends & (ev["src"] < n)]
    assert len(oks) > 40
    assert set(int(c) for c in errs["p0"]) <= {30}                 # conflicts only (no loss: nothing aborts)
    if n > 1:
        assert len(errs) > 0                                       # concurrent writers did collide on the root
    # committed writes form ONE chain of root pointers starting at "empty" (1)
    written = {}
    for p in oks["p1"]:
        rd, wr = int(p) & 0xFFFFFFFF, int(p) >> 32
        if rd != wr:
            assert rd not in written
            written[rd] = wr
    cur, steps = 1, 0
    while cur in written:
        cur = written[cur]
        steps += 1
    assert steps == len(written) > 20
    # the tree did grow past one leaf: some txn wrote nine nodes at once (a split) and later txns wrote paths
    writes = bd[(bd["type"] == O.T["write"]) & sends & (ev["src"] < n)]
    assert len(writes) > len(written) + 8


def test_txn_tree_promise_timeouts_under_loss(engine_backend):
    # 10 % of all messages are lost: a sync_rpc! whose request or reply is gone raises RPCError.timeout after 5 s
    # (promise.rb) -> error 0 to the client, the lock goes to the next waiter; replies that arrive late are ignored
    n = 3
    g, o = make_pair(n, workload="txn-list-append-tree", latency_dist="constant", latency_mean_ms=1, p_loss=0.1,
                     max_endpoints=n + 16, ring_cap=512, max_window=256, server_ring_cap=128, server_max_window=64, rpc_table=128,
                     tree_ptrs=2048, journal_cap_log2=20, calendar_slots=16, calendar_cap=2048, seed=99)

    def scenario(s, body):
        s.add_endpoint("lin-kv", O.KIND_SERVICE)
        s.add_endpoint("lww-kv", O.KIND_SERVICE)
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(3)]
        s.schedule(ops_array([(0, cs[i], i, "init", 1, 0) for i in range(n)]))
        rng = np.random.default_rng(5)
        s.schedule(txn_ops(rng, n, cs, 20, 12, 3, 30, [100] * 3))
        s.run(7_500_000_000)
        return s.client_replies()

    rg, ro = both(g, o, scenario)
    assert rg == ro
    ev, bd = assert_same_journal(g, o)
    sends = (ev["event_id"] >> np.uint64(63)) == 0
    errs = bd[(bd["type"] == O.T["error"]) & sends & (ev["src"] < n)]
    assert 0 in set(int(c) for c in errs["p0"])                    # somebody did time out
    assert int(((bd["type"] == O.T["txn_ok"]) & sends).sum()) > 0  # and somebody got through
    assert int(ev["time_ns"].max()) > 5_000_000_000


def test_mirror_serves_list_append_transactions(engine_backend):
    # what a Maelstrom client sees (workload/txn_list_append.clj): completed transactions, through maelstrom_b200.net /
    # client on top of the C ABI; the node needs its init first (the first node writes the empty tree and the root)
    import maelstrom_b200 as mb
    from maelstrom_b200 import client as C
    from maelstrom_b200.net import Net
    net = Net(mb.Sim(3, workload="txn-list-append-tree", max_endpoints=16, tree_ptrs=1024), mb.body).start_services(("lin-kv", "lww-kv"))
    c = C.Client(net)
    for i in range(3):
        assert c.rpc("n%d" % i, {"type": "init", "node_id": "n%d" % i, "node_ids": ["n0", "n1", "n2"]})["type"] == "init_ok"
    assert c.rpc("n0", {"type": "txn", "txn": [["r", 1, None]]})["txn"] == [["r", 1, None]]
    assert c.rpc("n1", {"type": "txn", "txn": [["append", 1, 10], ["r", 1, None]]})["txn"] == [["append", 1, 10], ["r", 1, [10]]]
    assert c.rpc("n2", {"type": "txn", "txn": [["append", 1, 11], ["append", 2, 5]]})["txn"] == [["append", 1, 11], ["append", 2, 5]]
    for k in range(3, 30):                                             # enough keys to split leaves and grow branches
        assert c.rpc("n%d" % (k % 3), {"type": "txn", "txn": [["append", k, k * 10]]})["type"] == "txn_ok"
    assert c.rpc("n0", {"type": "txn", "txn": [["r", 1, None], ["r", 2, None], ["r", 29, None], ["r", 77, None]]})["txn"] == \
        [["r", 1, [10, 11]], ["r", 2, [5]], ["r", 29, [290]], ["r", 77, None]]


@pytest.mark.gpu
def test_txn_tree_many_nodes_vs_oracle(engine_backend):
    # 200 nodes racing for one root: most cas calls lose (error 30), the winners' trees grow to three levels
    if engine_backend != "cuda":
        pytest.skip("B200 only (the emulator runs the small cases above)")
    n = 200
    g, o = make_pair(n, workload="txn-list-append-tree", latency_dist="constant", latency_mean_ms=1, max_endpoints=n + 16,
                     ring_cap=8192, max_window=4096, server_ring_cap=256, server_max_window=64, rpc_table=256,
                     tree_ptrs=2048, journal_cap_log2=23, calendar_slots=16, calendar_cap=1 << 15, seed=4242)

    def scenario(s, body):
        s.add_endpoint("lin-kv", O.KIND_SERVICE)
        s.add_endpoint("lww-kv", O.KIND_SERVICE)
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(8)]
        s.schedule(ops_array([(0, cs[i % 8], i, "init", 1 + i // 8, 0) for i in range(n)]))
        rng = np.random.default_rng(11)
        s.schedule(txn_ops(rng, n, cs, 60, 200, 20, 300, [1000] * 8))
        s.run(4_000_000_000)
        return s.client_replies()

    rg, ro = both(g, o, scenario)
    assert rg == ro == n + 4000
    ev, bd = assert_same_journal(g, o)
    sends = (ev["event_id"] >> np.uint64(63)) == 0
    assert int(((bd["type"] == O.T["txn_ok"]) & sends).sum()) > 300
    assert int(((bd["type"] == O.T["error"]) & sends & (ev["src"] < n) & (bd["p0"] == 30)).sum()) > 300


def tree_seeds():
    import os
    a, b = (int(x) for x in os.environ.get("MS_FUZZ_TREE_SEEDS", "0:4").split(":"))
    return list(range(a, b))


@pytest.mark.parametrize("seed", tree_seeds())
def test_txn_tree_random_scenarios(engine_backend, seed):
    # seeded: node count, latency law, key universe (few keys = deep splits of one hash range, many = wide trees),
    # request rate (queueing behind the txn lock), clients; journal equal to the oracle's
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(1, 20))
    dist = ("constant", "constant", "uniform")[int(rng.integers(3))]
    mean = 0 if dist == "constant" and rng.integers(2) else 1                 # (virtual seconds are slow on the emulator)
    n_keys = int(rng.choice([6, 30, 150, 2000]))
    per_tick = min(int(rng.integers(1, 7)), 2 * n)                              # (256 requests may wait for a node's txn lock)
    n_clients = int(rng.integers(1, 5))
    g, o = make_pair(n, workload="txn-list-append-tree", latency_dist=dist, latency_mean_ms=mean, max_endpoints=n + 16,
                     ring_cap=2048, max_window=1024, server_ring_cap=512, server_max_window=128, rpc_table=512,
                     tree_ptrs=8192, journal_cap_log2=21, calendar_slots=32, calendar_cap=8192, seed=int(rng.integers(1 << 40)))

    def scenario(s, body):
        s.add_endpoint("lin-kv", O.KIND_SERVICE)
        s.add_endpoint("lww-kv", O.KIND_SERVICE)
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(n_clients)]
        s.schedule(ops_array([(0, cs[i % n_clients], i, "init", 1 + i // n_clients, 0) for i in range(n)]))
        r2 = np.random.default_rng(seed)
        s.schedule(txn_ops(r2, n, cs, 30, 25, per_tick, n_keys, [100] * n_clients))
        s.run((300 + 1200 * mean) * 1_000_000)
        return s.client_replies()

    rg, ro = both(g, o, scenario)
    assert rg == ro and rg > n
    assert_same_journal(g, o)


def test_committed_golden_journal(engine_backend):
    # tests/golden/journals.json["txn_tree_four_nodes"], generated from the oracle (tests/golden/make_golden.py)
    import golden_cases as G
    G.check_engine_against_fixture("txn_tree_four_nodes")
