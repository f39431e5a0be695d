"""txn-list-append, the single-key variant ("simplest faithful variant
to build first"; demo/clojure/single_key_txn.clj:115-180): every node serves a txn by reading the
root from lin-kv, applying it and cas-ing the root; a lost race is error 30.  Engine vs oracle,
journal bit for bit.  [emul] = kernel sources on the CPU SIMT emulator; [cuda] = a B200."""
import numpy as np
import pytest

import oracle_lib as O
from scenarios import assert_same_journal, both, make_pair

pytestmark = pytest.mark.usefixtures("engine_backend")


def test_txns_conflicts_and_error_30():
    n = 3
    g, o = make_pair(n, workload="txn-list-append", max_endpoints=32, latency_dist="constant", latency_mean_ms=2)

    def scenario(s, body):
        kv = s.add_endpoint("lin-kv", O.KIND_SERVICE)
        cs = [s.add_endpoint("c%d" % i) for i in range(n)]
        for i in range(n):
            s.send(cs[i], i, body("init", msg_id=1))
        for i in range(n):
            assert int(s.recv(cs[i], 1_000_000_000)["type"]) == O.T["init_ok"]
        out = []

        def recv_all(c, k=1):
            got = []
            for _ in range(k):
                r = s.recv(c, 200_000_000)
                got.append(None if r is None else (int(r["type"]), int(r["in_reply_to"]), int(r["p0"]), int(r["p1"])))
            return got

        s.send(cs[0], 0, body("txn", msg_id=2, p1=101))                       # reads only, no root yet
        out += recv_all(cs[0])
        s.send(cs[0], 0, body("txn", msg_id=3, p1=102, appends=True))
        out += recv_all(cs[0])
        # three nodes race on the same root: one cas wins, the others get 22 from lin-kv -> error 30
        for i in range(n):
            s.send(cs[i], i, body("txn", msg_id=4, p1=200 + i, appends=True))
        for i in range(n):
            out += recv_all(cs[i])
        s.send(cs[1], 1, body("txn", msg_id=5, p1=300))                       # reads only: root unchanged
        out += recv_all(cs[1])
        s.send(cs[2], 2, body("broadcast", msg_id=6, p0=1))                   # unknown request type: error 10
        out += recv_all(cs[2])
        return out, kv

    (rg, kvg), (ro, kvo) = both(g, o, scenario)
    assert rg == ro
    T = O.T
    first, second = rg[0], rg[1]
    assert first[0] == T["txn_ok"] and first[3] == (0 | (1 << 32))             # nil -> {} (versions 0 -> 1)
    assert second[0] == T["txn_ok"] and (second[3] & 0xFFFFFFFF) == 1 and (second[3] >> 32) >= 2
    race = rg[2:5]
    assert sorted(r[0] for r in race) == sorted([T["txn_ok"], T["error"], T["error"]])
    assert [r[2] for r in race if r[0] == T["error"]] == [30, 30]              # single_key_txn.clj:170-172
    winner = [r for r in race if r[0] == T["txn_ok"]][0]
    assert (winner[3] & 0xFFFFFFFF) == second[3] >> 32
    ro_txn = rg[5]
    assert ro_txn[0] == T["txn_ok"] and (ro_txn[3] & 0xFFFFFFFF) == (ro_txn[3] >> 32) == winner[3] >> 32
    assert rg[6][0] == T["error"] and rg[6][2] == 10
    assert_same_journal(g, o)


def test_open_loop_txn_traffic_with_loss():
    n = 4
    g, o = make_pair(n, workload="txn-list-append", max_endpoints=32, ring_cap=512, max_window=512,
                     latency_dist="uniform", latency_mean_ms=2, p_loss=0.05)

    def scenario(s, body):
        kv = s.add_endpoint("lin-kv", O.KIND_SERVICE)
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(6)]
        rng = np.random.default_rng(21)
        rows = np.zeros(400, dtype=O.OP_DTYPE)
        for k in range(400):
            r = rows[k]
            r["time_ns"] = (k // 4) * 1_000_000
            r["src"] = cs[k % 6]
            r["dest"] = int(rng.integers(n))
            b = r["body"]
            b["type"] = O.T["txn"]
            b["flags"] = O.F_MSG_ID | (O.F_APPENDS if rng.integers(3) else 0)
            b["msg_id"] = k + 1
            b["p1"] = 1000 + k
        s.schedule(rows)
        s.run(160_000_000)
        return s.client_replies(), kv

    (rg, _), (ro, _) = both(g, o, scenario)
    assert rg == ro and rg > 200
    ev, bd = assert_same_journal(g, o)
    sends = (ev["event_id"] >> np.uint64(63)) == 0
    oks = bd[(bd["type"] == O.T["txn_ok"]) & sends]
    errs = bd[(bd["type"] == O.T["error"]) & sends & (ev["src"] < n)]          # what the nodes tell their clients
    assert len(oks) > 50 and set(int(c) for c in errs["p0"]) == {30}
    from_kv = bd[(bd["type"] == O.T["error"]) & sends & (ev["src"] >= n)]      # what lin-kv tells the nodes
    assert set(int(c) for c in from_kv["p0"]) == {20, 22}
    # the versions installed form one chain: every committed write starts from the previous one
    written = [(int(p) & 0xFFFFFFFF, int(p) >> 32) for p in oks["p1"] if (int(p) & 0xFFFFFFFF) != (int(p) >> 32)]
    assert len({w[0] for w in written}) == len(written) and len({w[1] for w in written}) == len(written)


def test_mirror_replays_apply_txn():
    # what a Maelstrom client sees (workload/txn_list_append.clj; doc/05-datomic): completed txns
    import maelstrom_b200 as mb
    from maelstrom_b200 import client as C
    from maelstrom_b200.net import Net
    net = Net(mb.Sim(2, workload="txn-list-append", max_endpoints=16), mb.body).start_services(("lin-kv",))
    c = C.Client(net)
    assert c.rpc("n0", {"type": "txn", "txn": [["r", 1, None]]})["txn"] == [["r", 1, None]]
    assert c.rpc("n0", {"type": "txn", "txn": [["append", 1, 10], ["r", 1, None]]})["txn"] == [["append", 1, 10], ["r", 1, [10]]]
    assert c.rpc("n1", {"type": "txn", "txn": [["append", 1, 11], ["append", 2, 5]]})["txn"] == [["append", 1, 11], ["append", 2, 5]]
    assert c.rpc("n0", {"type": "txn", "txn": [["r", 1, None], ["r", 2, None], ["r", 3, None]]})["txn"] == \
        [["r", 1, [10, 11]], ["r", 2, [5]], ["r", 3, None]]


def test_committed_golden_journal():
    # tests/golden/journals.json["txn_three_nodes"], generated from the oracle (tests/golden/make_golden.py)
    import golden_cases as G
    G.check_engine_against_fixture("txn_three_nodes")
