"""Named scenarios behind the committed fixtures in tests/golden/journals.json.  Each runs
unchanged against the oracle (tests/oracle_lib.Sim), the engine on a B200 and the engine on the
CPU emulator; the fixture pins the journal each must produce."""
import hashlib

import numpy as np

import oracle_lib as O
from scenarios import ops_array, random_broadcast_ops

SEED = 0x4D41454C


def _flood_grid25(s, body):
    c = s.add_endpoint("c0")
    for v in range(4):
        s.send(c, (7 * v) % 25, body("broadcast", msg_id=v + 1, p0=v))
    s.run(2_000_000)
    s.send(c, 24, body("read", msg_id=9))
    s.run(3_000_000)


def _echo_12_ops(s, body):
    # doc/02-echo/index.md:379-383: 12 ops on one node = 26 messages with the init pair
    c = s.add_endpoint("c0")
    s.send(c, 0, body("init", msg_id=1))
    s.run(1_000_000)
    for k in range(12):
        s.send(c, 0, body("echo", msg_id=k + 2, p0=k, p1=1000 + k))
        s.run((k + 2) * 1_000_000)


def _latency_loss_partition(s, body):
    cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(3)]
    ops, _ = random_broadcast_ops(36, cs, n_ticks=12, per_tick=5, seed=3)
    s.schedule(ops)
    s.run(4_000_000)
    s.set_loss(0.2)
    s.partition([0] * 18 + [1] * 18)
    s.run(9_000_000)
    s.heal()
    s.slow()
    s.run(40_000_000)
    s.fast()
    s.drop(3, 4)
    s.run(80_000_000)


def _gset_five_nodes(s, body):
    cs = []
    for i in range(5):
        cs.append(s.add_endpoint("c%d" % i))
        s.send(cs[i], i, body("init", msg_id=1))
    s.run(1_000_000)
    for k, v in enumerate((3, 7, 11, 200)):
        s.send(cs[k], k, body("add", msg_id=2, p0=v))
    s.send(cs[4], 4, body("read", msg_id=2))
    s.run(70_000_000)
    for i in range(5):
        s.send(cs[i], i, body("read", msg_id=3))
    s.run(75_000_000)


def _services_mixed(s, body):
    sv = {name: s.add_endpoint(name, O.KIND_SERVICE) for name in ("lin-kv", "seq-kv", "lww-kv", "lin-tso")}
    cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(3)]
    rng = np.random.default_rng(77)
    rows = np.zeros(150, dtype=O.OP_DTYPE)
    names = list(sv)
    for k in range(150):
        r = rows[k]
        r["time_ns"] = (k // 10) * 1_000_000
        r["src"] = cs[k % 3]
        name = names[int(rng.integers(4))]
        r["dest"] = sv[name]
        b = r["body"]
        b["flags"] = O.F_MSG_ID
        b["msg_id"] = k + 1
        if name == "lin-tso":
            b["type"] = O.T["ts"]
            continue
        b["p0"] = int(rng.integers(4))
        kind = int(rng.integers(3))
        b["type"] = (O.T["read"], O.T["write"], O.T["cas"])[kind]
        b["p1"] = int(rng.integers(4)) | ((int(rng.integers(4)) << 32) if kind == 2 else 0)
        if kind == 2 and rng.integers(2):
            b["flags"] |= O.F_CREATE
    s.schedule(rows)
    s.run(20_000_000)


def _txn_three_nodes(s, body):
    s.add_endpoint("lin-kv", O.KIND_SERVICE)
    cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(3)]
    rows = np.zeros(60, dtype=O.OP_DTYPE)
    for k in range(60):
        r = rows[k]
        r["time_ns"] = (k // 3) * 1_000_000
        r["src"] = cs[k % 3]
        r["dest"] = (k * 7) % 3
        b = r["body"]
        b["type"] = O.T["txn"]
        b["flags"] = O.F_MSG_ID | (O.F_APPENDS if k % 4 else 0)
        b["msg_id"] = k + 1
        b["p1"] = 500 + k
    s.schedule(rows)
    s.run(40_000_000)


def _txn_tree_four_nodes(s, body):
    # datomic_list_append.rb: inits (the first node writes the empty tree and the root), then 90 txns of 1-3
    # micro-ops on 24 keys: leaves split, paths are copied, roots collide
    s.add_endpoint("lin-kv", O.KIND_SERVICE)
    s.add_endpoint("lww-kv", O.KIND_SERVICE)
    cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(3)]
    rows = np.zeros(4 + 90, dtype=O.OP_DTYPE)
    for i in range(4):
        r = rows[i]
        r["time_ns"] = 0
        r["src"] = cs[i % 3]
        r["dest"] = i
        r["body"]["type"] = O.T["init"]
        r["body"]["flags"] = O.F_MSG_ID
        r["body"]["msg_id"] = 1 + i
    for k in range(90):
        r = rows[4 + k]
        r["time_ns"] = (20 + k // 3) * 1_000_000
        r["src"] = cs[k % 3]
        r["dest"] = (k * 5) % 4
        b = r["body"]
        b["type"] = O.T["txn"]
        b["flags"] = O.F_MSG_ID
        b["msg_id"] = 100 + k
        b["p0"] = k
        w = 0
        for j in range(1 + k % 3):
            key = (k * 7 + j * 11) % 24
            w |= (0x8000 | (0 if (k + j) % 4 == 0 else 0x4000) | key) << (16 * j)
        b["p1"] = w
    s.schedule(rows)
    s.run(400_000_000)


def _raft_three_nodes(s, body):
    cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(2)]
    for i in range(3):
        s.send(cs[0], i, body("init", msg_id=100 + i))
    rows = np.zeros(24, dtype=O.OP_DTYPE)
    for k in range(24):
        r = rows[k]
        r["time_ns"] = 4_100_000_000 + (k // 2) * 1_000_000
        r["src"] = cs[k % 2]
        r["dest"] = k % 3
        b = r["body"]
        b["flags"] = O.F_MSG_ID
        b["msg_id"] = k + 1
        b["p0"] = k % 2
        b["type"] = (O.T["write"], O.T["read"], O.T["cas"])[k % 3]
        b["p1"] = (k % 5) | (((k + 1) % 5) << 32 if k % 3 == 2 else 0)
    s.schedule(rows)
    s.run(4_250_000_000)


CASES = {
    "flood_grid25": (dict(n_nodes=25, workload="broadcast", topology="grid", n_values=8), _flood_grid25),
    "echo_12_ops": (dict(n_nodes=1, workload="echo"), _echo_12_ops),
    "latency_loss_partition": (dict(n_nodes=36, workload="broadcast", topology="grid", n_values=128,
                                    latency_dist="exponential", latency_mean_ms=4), _latency_loss_partition),
    "gset_five_nodes": (dict(n_nodes=5, workload="g-set", n_values=256, gset_interval_ms=30), _gset_five_nodes),
    "services_mixed": (dict(n_nodes=2, workload="echo", latency_dist="uniform", latency_mean_ms=2), _services_mixed),
    "txn_three_nodes": (dict(n_nodes=3, workload="txn-list-append", latency_dist="constant", latency_mean_ms=1),
                        _txn_three_nodes),
    "raft_three_nodes": (dict(n_nodes=3, workload="lin-kv"), _raft_three_nodes),
    "txn_tree_four_nodes": (dict(n_nodes=4, workload="txn-list-append-tree", latency_dist="constant", latency_mean_ms=1),
                            _txn_tree_four_nodes),
}
# cases checked against the engine by tests/test_golden_fixtures.py; the others by their workload's test file
CORE_CASES = ("flood_grid25", "echo_12_ops", "latency_loss_partition")
ENGINE_SIZING = dict(ring_cap=256, max_window=256, journal_cap_log2=18, max_endpoints=64,
                     calendar_slots=1024, calendar_cap=2048)
W = {"echo": O.W_ECHO, "broadcast": O.W_BROADCAST, "g-set": O.W_GSET, "lin-kv": O.W_RAFT,
     "txn-list-append": O.W_TXN, "txn-list-append-tree": O.W_TXN_TREE}


def make_oracle(name):
    kw = dict(CASES[name][0])
    n = kw.pop("n_nodes")
    return O.Sim(n, workload=W[kw.pop("workload")], seed=SEED, **kw)


def make_engine(name):
    import maelstrom_b200 as mb
    kw = dict(CASES[name][0])
    n = kw.pop("n_nodes")
    return mb.Sim(n, seed=SEED, **kw, **ENGINE_SIZING)


def digest(events, bodies, stats, now, rnd):
    """What the fixture stores: sizes, sha256 of the packed event / body arrays, stats, clock."""
    ev = np.ascontiguousarray(events)
    bd = np.ascontiguousarray(bodies)
    return {
        "n_events": int(len(ev)),
        "events_sha256": hashlib.sha256(ev.tobytes()).hexdigest(),
        "bodies_sha256": hashlib.sha256(np.stack([bd[f].astype(np.uint64) for f in
                                                  ("type", "flags", "msg_id", "in_reply_to", "p0", "p1")]).tobytes()).hexdigest(),
        "first_events": [[int(e["event_id"] & np.uint64(0x7FFFFFFFFFFFFFFF)), int(e["event_id"] >> np.uint64(63)),
                          int(e["time_ns"]), int(e["msg_id"]), int(e["src"]), int(e["dest"])] for e in ev[:6]],
        "stats": stats,
        "now_ns": int(now),
        "rounds": int(rnd),
    }


def check_engine_against_fixture(name):
    """Run case `name` on the engine (whatever backend is active) and compare with journals.json."""
    import json
    import os
    import maelstrom_b200 as mb
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "journals.json")))
    g = make_engine(name)
    CASES[name][1](g, mb.body)
    ev, bd = g.drain()
    assert digest(ev, bd, g.stats(), g.now, g.round) == want[name]
