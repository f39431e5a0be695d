"""Device-resident services (SURVEY.md section 8f NEXT-3; src/maelstrom/service.clj): lin-kv,
seq-kv, lww-kv and lin-tso as endpoints of the engine, against the oracle's restatement
(itself pinned to test/maelstrom/service_test.clj by tests/test_oracle_services.py), journal
bit for bit.  [emul] = kernel sources on the CPU SIMT emulator, [cuda] = a B200."""
import numpy as np
import pytest

import oracle_lib as O
from scenarios import assert_same_journal, both, make_pair

pytestmark = pytest.mark.usefixtures("engine_backend")
SVC = ("lin-kv", "seq-kv", "lww-kv", "lin-tso")


def start_services(s):
    return {name: s.add_endpoint(name, O.KIND_SERVICE) for name in SVC}


def rpc(s, body, c, dest, mid, type, **kw):
    s.send(c, dest, body(type, msg_id=mid, **kw))
    r = s.recv(c, 1_000_000_000)
    assert r is not None
    return int(r["type"]), int(r["in_reply_to"]), int(r["p0"]), int(r["p1"])


def test_lin_kv_and_tso_rpc_by_rpc():
    g, o = make_pair(3, workload="broadcast", n_values=8, max_endpoints=16)

    def scenario(s, body):
        sv = start_services(s)
        c = s.add_endpoint("c0")
        kv, tso = sv["lin-kv"], sv["lin-tso"]
        out = [rpc(s, body, c, kv, 1, "read", p0=5)]                                   # error 20
        out.append(rpc(s, body, c, kv, 2, "write", p0=5, p1=3))
        out.append(rpc(s, body, c, kv, 3, "read", p0=5))
        out.append(rpc(s, body, c, kv, 4, "cas", p0=5, p1=4 | (9 << 32)))              # error 22
        out.append(rpc(s, body, c, kv, 5, "cas", p0=5, p1=3 | (9 << 32)))
        out.append(rpc(s, body, c, kv, 6, "cas", p0=6, p1=0 | (1 << 32)))              # error 20
        out.append(rpc(s, body, c, kv, 7, "cas", p0=6, p1=0 | (1 << 32), create=True))
        out.append(rpc(s, body, c, kv, 8, "read", p0=6))
        out += [rpc(s, body, c, tso, 10 + k, "ts") for k in range(3)]
        s.send(c, kv, body("ts", msg_id=20))                                           # no clause: no reply
        assert s.recv(c, 5_000_000) is None
        return out

    rg, ro = both(g, o, scenario)
    assert rg == ro
    T = O.T
    assert [r[0] for r in rg] == [T["error"], T["write_ok"], T["read_ok"], T["error"], T["cas_ok"], T["error"],
                                  T["cas_ok"], T["read_ok"], T["ts_ok"], T["ts_ok"], T["ts_ok"]]
    assert [r[2] for r in rg][:6] == [20, 0, 0, 22, 0, 20] and [r[3] for r in rg][-3:] == [0, 1, 2]
    assert rg[2][3] == 3 and rg[7][3] == 1
    assert_same_journal(g, o)


def random_service_ops(n_clients, first_client, svc, n_ticks, per_tick, seed, n_keys=6):
    """open-loop traffic from simulated clients to the four services"""
    rng = np.random.default_rng(seed)
    rows = np.zeros(n_ticks * per_tick, dtype=O.OP_DTYPE)
    mid = [0] * n_clients
    k = 0
    for t in range(n_ticks):
        for _ in range(per_tick):
            c = int(rng.integers(n_clients))
            mid[c] += 1
            name = SVC[int(rng.integers(4))]
            r = rows[k]
            k += 1
            r["time_ns"] = t * 1_000_000
            r["src"] = first_client + c
            r["dest"] = svc[name]
            b = r["body"]
            b["flags"] = O.F_MSG_ID
            b["msg_id"] = mid[c]
            if name == "lin-tso":
                b["type"] = O.T["ts"]
                continue
            kind = int(rng.integers(10))
            b["p0"] = int(rng.integers(n_keys))
            if kind < 4:
                b["type"] = O.T["read"]
            elif kind < 7:
                b["type"] = O.T["write"]
                b["p1"] = int(rng.integers(5))
            else:
                b["type"] = O.T["cas"]
                b["p1"] = int(rng.integers(5)) | (int(rng.integers(5)) << 32)
                if kind == 9:
                    b["flags"] |= O.F_CREATE
    return rows


@pytest.mark.parametrize("workload,dist,mean", [("broadcast", "constant", 0), ("g-set", "uniform", 3)])
def test_mixed_service_traffic(workload, dist, mean):
    # many requests per window: ordering, one-at-a-time handling, seq-kv's timeline, lww-kv's replica pick
    g, o = make_pair(4, workload=workload, n_values=64, latency_dist=dist, latency_mean_ms=mean,
                     max_endpoints=32, ring_cap=512, max_window=512, gset_interval_ms=11)

    def scenario(s, body):
        sv = start_services(s)
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(6)]
        s.schedule(random_service_ops(6, cs[0], sv, n_ticks=30, per_tick=25, seed=5))
        if workload == "g-set":
            for i in range(4):
                s.send(cs[0], i, body("init", msg_id=900 + i))
        s.run(45_000_000)

    both(g, o, scenario)
    ev, bd = assert_same_journal(g, o)
    T = O.T
    kinds = set(int(t) for t in bd["type"])
    assert {T["read_ok"], T["write_ok"], T["cas_ok"], T["ts_ok"], T["error"]} <= kinds
    codes = set(int(p) for p in bd["p0"][bd["type"] == T["error"]])
    assert codes == {20, 22}


def test_services_pay_server_latency_clients_do_not():
    # net.clj:178-187: latency applies unless a client is involved; services are not clients
    g, o = make_pair(2, workload="echo", latency_dist="constant", latency_mean_ms=7, max_endpoints=16)

    def scenario(s, body):
        sv = start_services(s)
        h = s.add_endpoint("h0", O.KIND_HOST)            # a non-client endpoint talking to a service
        c = s.add_endpoint("c0")
        s.send(h, sv["lin-tso"], body("ts", msg_id=1))
        s.send(c, sv["lin-tso"], body("ts", msg_id=1))
        rc = s.recv(c, 1_000_000_000)
        t_client = s.now
        rh = s.recv(h, 1_000_000_000)
        return int(rc["p1"]), int(rh["p1"]), t_client, s.now

    rg, ro = both(g, o, scenario)
    assert rg == ro
    # client round trip: within the first tick; host <-> service: 7 ms each way
    assert rg[2] <= 1_000_000 and 14_000_000 <= rg[3] <= 15_000_000, rg
    assert_same_journal(g, o)


def test_mirror_rpcs_read_like_the_reference():
    # doc/services.md: clients talk to services with plain RPCs
    import maelstrom_b200 as mb
    from maelstrom_b200 import client as C
    from maelstrom_b200.net import Net
    net = Net(mb.Sim(1, workload="echo", max_endpoints=16), mb.body).start_services()
    c = C.Client(net)
    assert c.rpc("lin-kv", {"type": "write", "key": 1, "value": 4})["type"] == "write_ok"
    assert c.rpc("lin-kv", {"type": "read", "key": 1})["value"] == 4
    with pytest.raises(C.RPCError) as e:
        c.rpc("lin-kv", {"type": "cas", "key": 1, "from": 3, "to": 5})
    assert e.value.code == 22
    assert c.rpc("lin-kv", {"type": "cas", "key": 2, "from": 0, "to": 5, "create_if_not_exists": True})["type"] == "cas_ok"
    assert [c.rpc("lin-tso", {"type": "ts"})["ts"] for _ in range(3)] == [0, 1, 2]
    with pytest.raises(C.RPCError) as e:
        c.rpc("seq-kv", {"type": "read", "key": 9})
    assert e.value.code == 20


def test_committed_golden_journal():
    # tests/golden/journals.json["services_mixed"], generated from the oracle (tests/golden/make_golden.py)
    import golden_cases as G
    G.check_engine_against_fixture("services_mixed")
