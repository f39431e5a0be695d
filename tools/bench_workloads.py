#!/usr/bin/env python
"""First numbers for the node programs that are not the headline benchmark (bench.py stays the
contract for that one): g-set, services, txn-list-append and Raft on ONE GPU, each timed with the
engine's own CUDA-event timer around ms_run, journal level 0 (kernel path), inputs resident.

    python tools/bench_workloads.py                 # B200
    python tools/bench_workloads.py --emul --tiny   # script check on the CPU emulator (test infra)

Prints one JSON line per workload: delivered messages per second of wall time on the device,
rounds, virtual time covered, and the algorithmic bytes the design note assigns to the workload.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ops_array(n, dtype):
    return np.zeros(n, dtype=dtype)


def timed_run(sim, until_ns):
    sim.timer_begin()
    t0 = time.time()
    sim.run(until_ns)
    wall = time.time() - t0
    ms = sim.timer_end()
    c = sim.counters()
    return dict(device_ms=ms, wall_s=wall, rounds=c["rounds"], sends=c["sends"], recvs=c["recvs"],
                launches=c["launches"], msgs_per_s=(c["recvs"] / (ms / 1e3)) if ms > 0 else None)


def bench_gset(mb, n, interval_ms, n_values, ticks, adds_per_tick):
    from maelstrom_b200.engine import KIND_SIM_CLIENT, OP_DTYPE, TYPES, F_MSG_ID
    sim = mb.Sim(n, workload="g-set", n_values=n_values, gset_interval_ms=interval_ms, journal_level=0,
                 latency_dist="exponential", latency_mean_ms=5, ring_cap=max(1024, 2 * n), max_window=max(1024, 2 * n),
                 max_endpoints=n + 8, calendar_slots=256, calendar_cap=max(1 << 16, 8 * n * n // 16))
    c = sim.add_endpoint("c0", KIND_SIM_CLIENT)
    rows = ops_array(n + ticks * adds_per_tick, OP_DTYPE)
    for i in range(n):
        rows[i]["src"], rows[i]["dest"] = c, i
        rows[i]["body"]["type"], rows[i]["body"]["flags"], rows[i]["body"]["msg_id"] = TYPES["init"], F_MSG_ID, i + 1
    rng = np.random.default_rng(1)
    k = n
    for t in range(ticks):
        for _ in range(adds_per_tick):
            r = rows[k]
            k += 1
            r["time_ns"] = (1 + t) * 1_000_000
            r["src"], r["dest"] = c, int(rng.integers(n))
            r["body"]["type"], r["body"]["flags"], r["body"]["msg_id"] = TYPES["add"], F_MSG_ID, k
            r["body"]["p0"] = int(rng.integers(n_values))
    sim.schedule(rows)
    out = timed_run(sim, (ticks + 3 * interval_ms + 60) * 1_000_000)
    out.update(workload="g-set", nodes=n, interval_ms=interval_ms,
               algorithmic_bytes_per_replicate_full=272 + 4 * ((n_values + 31) // 32))
    return out


def bench_services(mb, per_tick, ticks):
    from maelstrom_b200.engine import KIND_SERVICE, KIND_SIM_CLIENT, OP_DTYPE, TYPES, F_MSG_ID
    sim = mb.Sim(1, workload="echo", journal_level=0, ring_cap=1 << 15, max_window=1 << 13, max_endpoints=64)
    sv = [sim.add_endpoint(name, KIND_SERVICE) for name in ("lin-kv", "seq-kv", "lww-kv", "lin-tso")]
    cs = [sim.add_endpoint("c%d" % i, KIND_SIM_CLIENT) for i in range(16)]
    rng = np.random.default_rng(2)
    rows = ops_array(per_tick * ticks, OP_DTYPE)
    rows["time_ns"] = (np.arange(len(rows)) // per_tick) * 1_000_000
    rows["src"] = np.asarray(cs)[rng.integers(len(cs), size=len(rows))]
    which = rng.integers(4, size=len(rows))
    rows["dest"] = np.asarray(sv)[which]
    kind = rng.integers(3, size=len(rows))
    rows["body"]["type"] = np.where(which == 3, TYPES["ts"], np.asarray([TYPES["read"], TYPES["write"], TYPES["cas"]])[kind])
    rows["body"]["flags"] = F_MSG_ID
    rows["body"]["msg_id"] = np.arange(len(rows)) + 1
    rows["body"]["p0"] = rng.integers(1024, size=len(rows))
    rows["body"]["p1"] = rng.integers(16, size=len(rows)) | (rng.integers(16, size=len(rows)) << 32)
    sim.schedule(rows)
    out = timed_run(sim, (ticks + 2) * 1_000_000)
    out.update(workload="services", requests_per_tick=per_tick, algorithmic_bytes_per_request=2 * 272)
    return out


def bench_txn(mb, n, per_tick, ticks):
    from maelstrom_b200.engine import KIND_SERVICE, KIND_SIM_CLIENT, OP_DTYPE, TYPES, F_MSG_ID, F_APPENDS
    sim = mb.Sim(n, workload="txn-list-append", journal_level=0, ring_cap=1 << 15, max_window=1 << 13,
                 max_endpoints=n + 32)
    sim.add_endpoint("lin-kv", KIND_SERVICE)
    cs = [sim.add_endpoint("c%d" % i, KIND_SIM_CLIENT) for i in range(16)]
    rng = np.random.default_rng(3)
    rows = ops_array(per_tick * ticks, OP_DTYPE)
    rows["time_ns"] = (np.arange(len(rows)) // per_tick) * 1_000_000
    rows["src"] = np.asarray(cs)[rng.integers(len(cs), size=len(rows))]
    rows["dest"] = rng.integers(n, size=len(rows))
    rows["body"]["type"] = TYPES["txn"]
    rows["body"]["flags"] = F_MSG_ID | np.where(rng.integers(2, size=len(rows)) == 1, F_APPENDS, 0)
    rows["body"]["msg_id"] = np.arange(len(rows)) + 1
    rows["body"]["p1"] = np.arange(len(rows)) + 1
    sim.schedule(rows)
    out = timed_run(sim, (ticks + 4) * 1_000_000)
    out.update(workload="txn-list-append", nodes=n, txns_per_tick=per_tick, messages_per_txn=6)
    return out


def bench_raft(mb, n, virtual_ms, writes_per_tick):
    from maelstrom_b200.engine import KIND_SIM_CLIENT, OP_DTYPE, TYPES, F_MSG_ID
    sim = mb.Sim(n, workload="lin-kv", journal_level=0, ring_cap=1 << 12, max_window=1 << 11, max_endpoints=n + 16)
    cs = [sim.add_endpoint("c%d" % i, KIND_SIM_CLIENT) for i in range(4)]
    t0 = 4200
    n_ops = n + (virtual_ms - t0) * writes_per_tick
    rows = ops_array(n_ops, OP_DTYPE)
    for i in range(n):
        rows[i]["src"], rows[i]["dest"] = cs[0], i
        rows[i]["body"]["type"], rows[i]["body"]["flags"], rows[i]["body"]["msg_id"] = TYPES["init"], F_MSG_ID, i + 1
    rng = np.random.default_rng(4)
    k = np.arange(n, n_ops)
    rows["time_ns"][n:] = (t0 + (k - n) // writes_per_tick) * 1_000_000
    rows["src"][n:] = np.asarray(cs)[rng.integers(4, size=len(k))]
    rows["dest"][n:] = rng.integers(n, size=len(k))
    rows["body"]["type"][n:] = np.asarray([TYPES["read"], TYPES["write"]])[rng.integers(2, size=len(k))]
    rows["body"]["flags"][n:] = F_MSG_ID
    rows["body"]["msg_id"][n:] = k + 1
    rows["body"]["p0"][n:] = rng.integers(64, size=len(k))
    rows["body"]["p1"][n:] = rng.integers(1000, size=len(k))
    sim.schedule(rows)
    out = timed_run(sim, virtual_ms * 1_000_000)
    out.update(workload="lin-kv (Raft)", nodes=n, virtual_ms=virtual_ms, leader=[i for i in range(n) if sim.raft_state(i)["state"] == 3])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emul", action="store_true", help="run on the CPU SIMT emulator (script check only)")
    ap.add_argument("--tiny", action="store_true")
    a = ap.parse_args()
    import maelstrom_b200 as mb
    ctx = None
    if a.emul:
        import emul_lib
        ctx = emul_lib.use()
        ctx.__enter__()
    try:
        if a.tiny:
            runs = [lambda: bench_gset(mb, 12, 8, 512, 10, 4), lambda: bench_services(mb, 40, 5),
                    lambda: bench_txn(mb, 3, 12, 5), lambda: bench_raft(mb, 3, 4260, 2)]
        else:
            runs = [lambda: bench_gset(mb, 1024, 100, 1 << 14, 200, 64), lambda: bench_services(mb, 1 << 14, 20),
                    lambda: bench_txn(mb, 256, 1 << 11, 20), lambda: bench_raft(mb, 5, 10_000, 4)]
        for r in runs:
            print(json.dumps(r(), sort_keys=True), flush=True)
    finally:
        if ctx:
            ctx.__exit__(None, None, None)


if __name__ == "__main__":
    main()
