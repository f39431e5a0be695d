#!/usr/bin/env python
"""bench.py -- simulated msgs/sec on the broadcast workload, 4096 nodes, grid
topology (BASELINE.json configs[1]) through the C ABI of maelstrom_b200.

A "step" is one virtual tick (1 ms): V broadcast requests are injected by
simulated clients at Philox-random nodes and the engine runs delta rounds until
the flood of every value has died out (12 033 server messages per value on the
64x64 grid) and virtual time advances.

  value      delivered messages / second, inputs (the op schedule) resident in HBM,
             journal written to HBM (32-B events), timed with CUDA events on the
             engine's stream (ms_timer_begin/end);
  e2e        same metric through the host-buffer path: every step uploads its ops
             from host memory (ms_schedule_ops), runs, and drains the full journal
             into pinned host memory (ms_journal_drain);
  roofline   round kernel only: (128*sends + 144*recvs) algorithmic bytes
              / sum of its launch durations measured with CUDA events;
  cpu_baseline  the CPU oracle (oracle/, a port of net.clj's rules) on a bounded
             sample of the same workload on the host cores.

`--impl reference` times the CPU restatement instead (the JVM reference cannot
run on this box: no java/lein), on all host cores as independent replicas.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES = 4096
TICK_NS = 1_000_000
SEED = 0x4D41454C          # "MAEL"
ALG_SEND_B, ALG_RECV_B = 128, 144
N_CLIENTS = 64


def philox_nodes(n, stream, offset=0):
    """Destination node of op i: Philox4x32-10(counter=i, key=(SEED, stream)) mod N (numpy restatement)."""
    c = np.zeros((n, 4), dtype=np.uint64)
    c[:, 0] = np.arange(offset, offset + n, dtype=np.uint64) & 0xFFFFFFFF
    k0 = np.full(n, SEED & 0xFFFFFFFF, dtype=np.uint64)
    k1 = np.full(n, stream, dtype=np.uint64)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    for _ in range(10):
        a = M0 * c0
        b = M1 * c2
        n0 = (b >> np.uint64(32)) ^ c1 ^ k0
        n2 = (a >> np.uint64(32)) ^ c3 ^ k1
        c1 = b & mask
        c3 = a & mask
        c0, c2 = n0 & mask, n2 & mask
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return (c0 % np.uint64(N_NODES)).astype(np.uint32)


def make_ops(op_dtype, first_tick, n_ticks, per_tick, client0, n_clients, type_code, flag_msg_id):
    """V broadcast requests per tick from simulated clients (round-robin), dense value ids."""
    n = n_ticks * per_tick
    ops = np.zeros(n, dtype=op_dtype)
    i = np.arange(n, dtype=np.uint64)
    g = i + np.uint64(first_tick * per_tick)        # global op index = broadcast value id
    ops["time_ns"] = ((g // np.uint64(per_tick)) * np.uint64(TICK_NS)).astype(np.int64)
    ops["src"] = (client0 + (g % np.uint64(n_clients))).astype(np.uint32)
    ops["dest"] = philox_nodes(n, 1, offset=first_tick * per_tick)
    ops["body"]["type"] = type_code
    ops["body"]["flags"] = flag_msg_id
    ops["body"]["msg_id"] = (g // np.uint64(n_clients) + np.uint64(1)).astype(np.uint32)
    ops["body"]["p0"] = g.astype(np.uint32)
    return ops


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        threading.Thread.__init__(self, daemon=True)
        self.rows = []
        self.proc = None
        self.index = index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [nm for j, nm in enumerate(names)
                   if any(len(r) > 3 + j and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU restatement
def _oracle_worker(args):
    per_tick, ticks, seed = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    s = O.Sim(N_NODES, workload=O.W_BROADCAST, topology="grid", n_values=per_tick * ticks + 1, seed=seed)
    c0 = None
    for i in range(N_CLIENTS):
        c = s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT)
        c0 = c if c0 is None else c0
    ops = make_ops(O.OP_DTYPE, 0, ticks, per_tick, c0, N_CLIENTS, O.T["broadcast"], O.F_MSG_ID)
    s.schedule(ops)
    t0 = time.perf_counter()
    s.run(ticks * TICK_NS)
    dt = time.perf_counter() - t0
    st = s.stats()["all"]
    return st["recv-count"], dt


def cpu_run(per_tick, ticks, procs):
    """`procs` independent replicas of the oracle (different seeds) on the host cores."""
    if procs == 1:
        res = [_oracle_worker((per_tick, ticks, SEED))]
        wall = res[0][1]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Pool(procs) as pool:
            t0 = time.perf_counter()
            res = pool.map(_oracle_worker, [(per_tick, ticks, SEED + i) for i in range(procs)])
            wall = time.perf_counter() - t0
        wall = max(r[1] for r in res)
    msgs = sum(r[0] for r in res)
    return msgs / wall, msgs, wall


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_tick = max(16, args.cpu_values // 24)     # per replica and step: keeps K+W steps within a few minutes
    # calibrate: 1 value = 12 035 messages; the single-thread port does a few M msgs/s
    vals, times = [], []
    for _ in range(args.warmup):
        cpu_run(per_tick, 1, cores)
    t_all = 0.0
    m_all = 0
    for _ in range(args.steps):
        v, msgs, wall = cpu_run(per_tick, 1, cores)
        t_all += wall
        m_all += msgs
    value = m_all / t_all
    line = {
        "impl": "reference", "metric": "simulated msgs/sec (broadcast, 4096 nodes)", "value": value,
        "unit": "msgs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_all / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "broadcast, 4096 nodes, grid 64x64, latency constant 0",
                   "values_per_tick": per_tick, "note": "CPU restatement of net.clj (oracle/), not the JVM: "
                   "no java/lein on this box; %d independent replicas, one per host core" % cores},
        "cpu_baseline": {"value": value, "unit": "msgs/s", "cores": cores, "kind": "port",
                         "sample": "%d replicas x %d values x 1 tick per step" % (cores, per_tick)},
        "e2e": {"value": value, "unit": "msgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def make_sim(mb, args, n_ticks_total, journal_discard, device, world=1):
    V = args.values_per_tick
    kw = dict(workload="broadcast", topology="grid", latency_dist="constant",
              latency_mean_ms=args.latency_ms, seed=SEED, n_values=V * n_ticks_total + 64,
              max_endpoints=N_NODES + N_CLIENTS, ring_cap=args.ring_cap, max_window=args.max_window,
              journal_level=1, journal_discard=1 if journal_discard else 0,
              journal_cap_log2=args.journal_cap_log2,
              threads_per_node=args.threads, calendar_cap=args.calendar_cap)
    if world > 1:
        # one shard per rank; cross-shard messages go over NVLink peer memory (maelstrom_b200/sharded.py)
        from maelstrom_b200.sharded import ShardedSim
        sim = ShardedSim(N_NODES, device=device, **kw)
    else:
        sim = mb.Sim(N_NODES, device=device, **kw)
    from maelstrom_b200.engine import KIND_SIM_CLIENT
    c0 = None
    for i in range(N_CLIENTS):
        c = sim.add_endpoint("c%d" % i, KIND_SIM_CLIENT)
        c0 = c if c0 is None else c0
    return sim, c0


def run_until_tick(sim, tick, drain=None):
    # ms_run returns 1 when the journal ring is half full: drain and continue
    while sim.run_raw(tick * TICK_NS) == 1:
        drain()


def gpu_arm(args, rank, world, local_rank):
    import torch
    import maelstrom_b200 as mb
    from maelstrom_b200.engine import TYPES, F_MSG_ID, OP_DTYPE

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    V = args.values_per_tick
    W, K = args.warmup, args.steps
    lat = args.latency_ms
    # with latency L ms a flood needs ~126*L ticks to die out; keep steps = ticks and let floods overlap
    total_ticks = W + K

    # ---- arm A: device-resident (value + roofline)
    sim, c0 = make_sim(mb, args, 3 * total_ticks + 4, True, local_rank, world)
    lstats = (lambda: sim.sim.stats()["all"]) if world > 1 else (lambda: sim.stats()["all"])   # this rank's endpoints
    ops = make_ops(OP_DTYPE, 0, 3 * total_ticks + 2, V, c0, N_CLIENTS, TYPES["broadcast"], F_MSG_ID)
    sim.schedule(ops)
    tick = 0
    for _ in range(W):
        tick += 1
        run_until_tick(sim, tick)
    before = lstats()
    c_before = sim.counters()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sim.timer_begin()
    for _ in range(K):
        tick += 1
        run_until_tick(sim, tick)
    ms_value = sim.timer_end()
    torch.cuda.synchronize()
    after = lstats()
    c_after = sim.counters()
    clocks = sampler.stop()
    recvs = after["recv-count"] - before["recv-count"]
    sends = after["send-count"] - before["send-count"]
    launches = c_after["launches"] - c_before["launches"]
    rounds = c_after["rounds"] - c_before["rounds"]

    if args.phases:
        sim.phase_cycles(True)
        for _ in range(K):
            tick += 1
            run_until_tick(sim, tick)
        import ctypes as _C
        _o = np.zeros(64, dtype=np.uint64)
        sim.L.ms_debug_phase_cycles(sim.h, 2, _o.ctypes.data)
        pc = _o.reshape(4, 16)
        names = ["fetch", "load", "order", "dedupe", "count+scan", "claims", "emit", "epilogue", "commit"]
        for c in range(4):
            nt_ = int(pc[c][15])
            if nt_:
                sys.stderr.write("class %d: %d tickets; cycles/ticket: %s\n" % (
                    c, nt_, ", ".join("%s=%d" % (nm, int(pc[c][k]) // nt_) for k, nm in enumerate(names))))
                sys.stderr.write("   fallback R>64: %d, verify-fail: %d, agg tickets: %d, non-agg gossip tickets: %d, blocks total: %d\n"
                                 % tuple(int(pc[c][k]) for k in (9, 10, 11, 12, 13)))
    # roofline pass: same work again (next K ticks) with CUDA events around every round-kernel launch
    sim.profile(True)
    sim.profile_read()
    b2 = lstats()
    for _ in range(K):
        tick += 1
        run_until_tick(sim, tick)
    a2 = lstats()
    k_ms, k_launches = sim.profile_read()
    sim.profile(False)
    alg_bytes = ALG_SEND_B * (a2["send-count"] - b2["send-count"]) + ALG_RECV_B * (a2["recv-count"] - b2["recv-count"])
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    max_window = c_after["max_window"]
    sim.close()

    # ---- arm B: end to end through host buffers
    e2e = None
    if not args.no_e2e and world > 1:
        # sharded e2e: host op buffers in every step, result read back every step (the 9 net-stats
        # counters); the full journal is not drained in sharded runs (per-shard rings, no global back-pressure)
        sim, c0 = make_sim(mb, args, total_ticks + 2, True, local_rank, world)
        host_ops = [make_ops(OP_DTYPE, t, 1, V, c0, N_CLIENTS, TYPES["broadcast"], F_MSG_ID)
                    for t in range(total_ticks)]
        tick = 0
        for step in range(total_ticks):
            if step == W:
                dist.barrier()
                torch.cuda.synchronize()
                s0 = sim.sim.stats()["all"]["recv-count"]
                t0 = time.perf_counter()
            sim.schedule(host_ops[step])
            tick += 1
            run_until_tick(sim, tick)
            sim.sim.stats()
        torch.cuda.synchronize()
        dist.barrier()
        t_e2e = time.perf_counter() - t0
        e2e = {"seconds": t_e2e, "msgs": sim.sim.stats()["all"]["recv-count"] - s0, "h2d": V * 40, "d2h": 72}
        sim.close()
    elif not args.no_e2e:
        sim, c0 = make_sim(mb, args, total_ticks + 2, False, local_rank)
        ev_cap = 1 << args.journal_cap_log2
        pinned = torch.empty(ev_cap * 32, dtype=torch.uint8, pin_memory=True)
        host_ops = [make_ops(OP_DTYPE, t, 1, V, c0, N_CLIENTS, TYPES["broadcast"], F_MSG_ID)
                    for t in range(total_ticks)]
        tick = 0
        d2h = 0
        t_e2e = 0.0
        msgs_e2e = 0
        for step in range(total_ticks):
            if step == W:
                if dist:
                    dist.barrier()
                torch.cuda.synchronize()
                s0 = sim.stats()["all"]["recv-count"]
                t0 = time.perf_counter()
            def drain_all():                                  # device -> host: the journal so far
                nonlocal d2h
                while True:
                    n = sim.drain_into(pinned.data_ptr(), ev_cap)
                    if step >= W:
                        d2h += n * 32
                    if n < ev_cap:
                        break
            sim.schedule(host_ops[step])                      # host -> device: this step's ops
            tick += 1
            run_until_tick(sim, tick, drain_all)
            drain_all()
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        msgs_e2e = sim.stats()["all"]["recv-count"] - s0
        sim.close()
        e2e = {"seconds": t_e2e, "msgs": msgs_e2e, "h2d": V * 40, "d2h": d2h // max(K, 1)}

    # ---- aggregate over ranks (max time, sum of work)
    if dist:
        t = torch.tensor([ms_value, e2e["seconds"] if e2e else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w = torch.tensor([recvs, sends, e2e["msgs"] if e2e else 0, launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        ms_value, e2e_s = float(t[0]), float(t[1])
        recvs, sends, e2e_msgs, launches = (int(x) for x in w.tolist())
    else:
        e2e_s = e2e["seconds"] if e2e else 0.0
        e2e_msgs = e2e["msgs"] if e2e else 0

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    value = recvs / (ms_value * 1e-3)
    line = {
        "metric": "simulated msgs/sec (broadcast, 4096 nodes)", "value": value, "unit": "msgs/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_value / K, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "broadcast, 4096 nodes, grid 64x64 (BASELINE.json configs[1])",
                   "latency": "constant %d ms" % lat, "values_per_step": V, "step": "1 virtual tick (1 ms)",
                   "delivered_msgs_per_step": recvs // max(K, 1), "rounds_per_step": rounds / max(K, 1),
                   "ring_cap": args.ring_cap, "max_window": args.max_window, "max_window_seen": max_window, "fallback_sorts": c_after["fallback_sorts"] - c_before["fallback_sorts"],
                   "l2_policy": "inputs larger than L2: inbox rings %.1f GB + seen bitmaps, streamed once per round"
                                % (N_NODES * args.ring_cap * 48 / 1e9),
                   "parallelism": ("%d shards by endpoint range, cross-shard messages written into peer inbox rings over NVLink, "
                                   "2 barriers per round" % world) if world > 1 else "single GPU",
                   "published_reference": "6e4 msgs/s, 48-way Xeon (README.md:39-42), different hardware"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of the four size-class launches of one
                     # round, ncu --set full at this configuration (profiles/r1e_summary.md)
                     "traffic": 532.5e6 if (V == 32768 and lat == 0 and world == 1) else None,
                     "traffic_unit": "bytes per round (one launch of each size class)",
                     "algorithmic_bytes_per_round": alg_bytes / max(k_launches, 1),
                     "kernel": "msd::k_round", "launches": k_launches, "avg_launch_us": 1e3 * k_ms / max(k_launches, 1),
                     "algorithmic_bytes_per_msg": ALG_SEND_B + ALG_RECV_B, "peak_source": peak_src},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = {"value": e2e_msgs / e2e_s, "unit": "msgs/s", "h2d_bytes_per_step": e2e["h2d"],
                       "d2h_bytes_per_step": e2e["d2h"],
                       "note": ("host op buffers in every step; the full journal (64 B/message) drained to pinned host "
                                "memory every step: PCIe-bound") if world == 1 else
                               ("host op buffers in every step; per step only the 9 net-stats counters are read back: "
                                "sharded runs drain the journal per shard, not measured here")}
    if world == 1 and not args.no_cpu:
        t0 = time.perf_counter()
        v, msgs, wall = cpu_run(args.cpu_values, 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": "msgs/s", "cores": 1, "kind": "port",
                                "sample": "%d values x 1 tick (%d msgs) in %.1f s, single-threaded oracle"
                                          % (args.cpu_values, msgs, wall)}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--values-per-tick", type=int, default=32768)
    ap.add_argument("--latency-ms", type=int, default=0)
    ap.add_argument("--ring-cap", type=int, default=8192)
    ap.add_argument("--max-window", type=int, default=4096)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--journal-cap-log2", type=int, default=26)
    ap.add_argument("--calendar-cap", type=int, default=1 << 20)
    ap.add_argument("--cpu-values", type=int, default=1536)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--phases", action="store_true", help="print per-phase cycle counts of the round kernel (stderr)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
