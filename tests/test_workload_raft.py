"""lin-kv workload served by Raft nodes (demo/python/raft.py): the
engine's Raft node program against the oracle's restatement, journal bit for bit, plus the
behaviour the reference documents for the algorithm (doc/06-raft): one leader per term,
proxying, error 11 without a leader, step-down and re-election behind a partition, committed
writes surviving a leader change.  [emul] = kernel sources on the CPU SIMT emulator; [cuda] = B200."""
import numpy as np
import pytest

import oracle_lib as O
from scenarios import assert_same_journal, both, make_pair

pytestmark = pytest.mark.usefixtures("engine_backend")
NASCENT, FOLLOWER, CANDIDATE, LEADER = 0, 1, 2, 3


def init_all(s, body, n):
    cs = []
    for i in range(n):
        c = s.add_endpoint("c%d" % i)
        cs.append(c)
        s.send(c, i, body("init", msg_id=1))
    return cs


def rpc(s, body, c, dest, mid, type, timeout_ns=5_000_000_000, **kw):
    s.send(c, dest, body(type, msg_id=mid, **kw))
    give_up = s.now + timeout_ns
    while True:
        r = s.recv(c, max(give_up - s.now, 0))
        if r is None:
            return None
        if int(r["in_reply_to"]) == mid:                       # late answers to abandoned requests are
            return int(r["type"]), int(r["p0"]), int(r["p1"]), int(r["in_reply_to"])   # dropped, client.clj:106-107


def leaders(s, n):
    return [i for i in range(n) if s.raft_state(i)["state"] == LEADER]


def test_election_replication_proxy_and_errors():
    n = 5
    g, o = make_pair(n, workload="lin-kv", max_endpoints=32, ring_cap=256, max_window=256)

    def scenario(s, body):
        init_all(s, body, n)
        c = s.add_endpoint("c9")
        out = [rpc(s, body, c, 2, 1, "read", p0=1)]                 # no leader known yet: error 11 (raft.py:565-570)
        s.run(4_500_000_000)                                        # election timeout is 2-4 s (raft.py:249-251)
        st = [s.raft_state(i) for i in range(n)]
        lead = leaders(s, n)
        assert len(lead) == 1 and all(x["term"] == st[lead[0]]["term"] for x in st)
        L, F = lead[0], (lead[0] + 1) % n
        out.append(rpc(s, body, c, L, 2, "write", p0=3, p1=7))
        out.append(rpc(s, body, c, L, 3, "read", p0=3))
        out.append(rpc(s, body, c, F, 4, "read", p0=3))             # proxied to the leader (raft.py:559-562)
        out.append(rpc(s, body, c, F, 5, "cas", p0=3, p1=7 | (9 << 32)))
        out.append(rpc(s, body, c, F, 6, "cas", p0=3, p1=7 | (9 << 32)))     # 22: expected 7 but had 9
        out.append(rpc(s, body, c, F, 7, "read", p0=44))                      # 20: not found
        s.run(s.now + 1_200_000_000)                                # a heartbeat later every log has caught up
        out.append([s.raft_state(i) for i in range(n)])
        return out

    rg, ro = both(g, o, scenario)
    assert rg == ro
    T = O.T
    assert rg[0][:2] == (T["error"], 11)
    assert [r[0] for r in rg[1:7]] == [T["write_ok"], T["read_ok"], T["read_ok"], T["cas_ok"], T["error"], T["error"]]
    assert rg[2][2] == 7 and rg[3][2] == 7 and rg[5][1] == 22 and rg[6][1] == 20
    final = rg[7]
    assert len({(x["log_size"], x["commit_index"]) for x in final}) == 1 and final[0]["log_size"] == 7
    assert_same_journal(g, o)


def test_partitioned_leader_steps_down_and_writes_survive():
    n = 5
    g, o = make_pair(n, workload="lin-kv", max_endpoints=32, ring_cap=256, max_window=256)

    def scenario(s, body):
        init_all(s, body, n)
        s.run(4_500_000_000)
        old = leaders(s, n)[0]
        c = s.add_endpoint("c9")
        out = [old, rpc(s, body, c, old, 1, "write", p0=1, p1=11)]
        s.partition([1 if i == old else 0 for i in range(n)])      # cut the leader off from its peers
        out.append(rpc(s, body, c, old, 2, "write", p0=2, p1=22, timeout_ns=500_000_000))   # cannot commit: no reply
        s.run(s.now + 4_500_000_000)                                # step-down after 2 s without acks; the others elect
        out.append([s.raft_state(i) for i in range(n)])
        s.heal()
        s.run(s.now + 4_500_000_000)
        lead = leaders(s, n)
        out.append(lead)
        out.append(rpc(s, body, c, lead[0], 3, "read", p0=1))       # the acknowledged write survived
        out.append(rpc(s, body, c, (lead[0] + 1) % n, 4, "read", p0=2))
        out.append([s.raft_state(i) for i in range(n)])
        return out

    rg, ro = both(g, o, scenario)
    assert rg == ro
    old, during, lead, after = rg[0], rg[3], rg[4], rg[7]
    assert rg[2] is None
    assert during[old]["state"] != LEADER                           # raft.py:371-376
    others = [i for i in range(n) if i != old and during[i]["state"] == LEADER]
    assert len(others) == 1 and during[others[0]]["term"] > during[old]["term"] - 2
    assert len(lead) == 1 and all(x["term"] == after[lead[0]]["term"] for x in after)
    assert rg[5][0] == O.T["read_ok"] and rg[5][2] == 11
    # the unacknowledged write is either committed by a later leader (22) or gone (error 20)
    assert rg[6][:3] in ((O.T["read_ok"], 0, 22), (O.T["error"], 20, 0))
    assert_same_journal(g, o)


def test_open_loop_clients_loss_and_services():
    # many requests per round through followers and the leader, 10 % loss, services running beside
    n = 3
    g, o = make_pair(n, workload="lin-kv", max_endpoints=32, ring_cap=512, max_window=512,
                     latency_dist="constant", latency_mean_ms=1)

    def scenario(s, body):
        sv = s.add_endpoint("lin-tso", O.KIND_SERVICE)
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(4)]
        for i in range(n):
            s.send(cs[0], i, body("init", msg_id=500 + i))
        rng = np.random.default_rng(11)
        rows = np.zeros(300, dtype=O.OP_DTYPE)
        for k in range(300):
            r = rows[k]
            r["time_ns"] = 4_200_000_000 + (k // 3) * 1_000_000
            r["src"] = cs[k % 4]
            b = r["body"]
            b["flags"] = O.F_MSG_ID
            b["msg_id"] = k + 1
            if k % 10 == 9:
                r["dest"] = sv
                b["type"] = O.T["ts"]
                continue
            r["dest"] = int(rng.integers(n))
            b["p0"] = int(rng.integers(4))
            kind = int(rng.integers(3))
            b["type"] = (O.T["read"], O.T["write"], O.T["cas"])[kind]
            if kind == 1:
                b["p1"] = int(rng.integers(4))
            elif kind == 2:
                b["p1"] = int(rng.integers(4)) | (int(rng.integers(4)) << 32)
        s.schedule(rows)
        s.run(4_100_000_000)
        s.set_loss(0.1)
        s.run(5_500_000_000)
        return [s.raft_state(i) for i in range(n)], s.client_replies()

    rg, ro = both(g, o, scenario)
    assert rg == ro and rg[1] > 100
    ev, bd = assert_same_journal(g, o)
    kinds = set(int(t) for t in bd["type"])
    assert {O.T["append_entries"], O.T["append_entries_res"], O.T["request_vote"], O.T["ts_ok"],
            O.T["read_ok"], O.T["write_ok"]} <= kinds


def test_committed_golden_journal():
    # tests/golden/journals.json["raft_three_nodes"], generated from the oracle (tests/golden/make_golden.py)
    import golden_cases as G
    G.check_engine_against_fixture("raft_three_nodes")
