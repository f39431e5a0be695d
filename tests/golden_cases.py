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


CASES = {
    "flood_grid25": (dict(n_nodes=25, workload="broadcast", topology="grid", n_values=8), _flood_grid25),
    "echo_12_ops": (dict(n_nodes=1, workload="echo"), _echo_12_ops),
    "latency_loss_partition": (dict(n_nodes=36, workload="broadcast", topology="grid", n_values=128,
                                    latency_dist="exponential", latency_mean_ms=4), _latency_loss_partition),
    "gset_five_nodes": (dict(n_nodes=5, workload="g-set", n_values=256, gset_interval_ms=30), _gset_five_nodes),
}
ENGINE_SIZING = dict(ring_cap=256, max_window=256, journal_cap_log2=18, max_endpoints=64,
                     calendar_slots=1024, calendar_cap=2048)
W = {"echo": O.W_ECHO, "broadcast": O.W_BROADCAST, "g-set": O.W_GSET}


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
