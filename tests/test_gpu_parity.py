"""GPU parity: the CUDA engine, driven through the C ABI, must produce a journal
(message ids, event ids, virtual times, delivery order, bodies) bit-identical
to the CPU oracle on the same seed and op sequence."""
import numpy as np
import pytest

import oracle_lib as O
from scenarios import assert_same_journal, both, make_pair, ops_array, random_broadcast_ops

# every test runs twice: [cuda] on the GPU box (gpu-marked) and [emul] in the CPU suite
pytestmark = pytest.mark.usefixtures("engine_backend")


@pytest.mark.gpu
def test_smoke_entry(engine_backend):
    if engine_backend != "cuda":
        pytest.skip("smoke() is the CUDA entry point")
    import __graft_entry__ as G
    G.smoke()


@pytest.mark.parametrize("topo,n,want", [
    ("grid", 5, 6), ("grid", 25, 56), ("line", 25, 24), ("total", 25, 576), ("tree4", 25, 24),
    ("grid", 1, 0), ("line", 1, 0), ("tree2", 7, 6), ("tree3", 40, 39),
])
def test_flood_counts_and_journal(topo, n, want):
    g, o = make_pair(n, topology=topo, n_values=8)

    def scenario(s, body):
        c = s.add_endpoint("c0")
        s.send(c, n // 3, body("broadcast", msg_id=1, p0=3))
        s.run(2_000_000)
        r = s.recv(c)
        return int(r["type"]), int(r["in_reply_to"]), int(r["id"])

    rg, ro = both(g, o, scenario)
    assert rg == ro
    assert_same_journal(g, o)
    assert g.stats()["servers"]["send-count"] == want
    for k in range(n):
        assert g.node_set(k).tolist() == [3]


def test_flood_4096_grid():
    n = 4096
    g, o = make_pair(n, topology="grid", n_values=64, max_endpoints=n + 8, ring_cap=256, max_window=256,
                     journal_cap_log2=20)

    def scenario(s, body):
        c = s.add_endpoint("c0")
        for v in range(3):
            s.send(c, (1234 * (v + 1)) % n, body("broadcast", msg_id=v + 1, p0=v))
        s.run(1_000_000)

    both(g, o, scenario)
    assert_same_journal(g, o)
    assert g.stats()["servers"]["send-count"] == 3 * 12033


def hot_broadcast_ops(n, clients, n_ticks, per_tick, hot, hot_permille, seed):
    """per_tick unique values per tick; hot_permille of them go to the few `hot` nodes, so the flood fronts that
    leave those nodes carry hundreds of values per round and the windows around them reach the upper classes."""
    rng = np.random.default_rng(seed)
    k = n_ticks * per_tick
    a = np.zeros(k, dtype=O.OP_DTYPE)
    i = np.arange(k)
    a["time_ns"] = (i // per_tick) * 1_000_000
    ci = rng.integers(0, len(clients), size=k)
    a["src"] = np.asarray(clients, dtype=np.uint32)[ci]
    is_hot = rng.integers(0, 1000, size=k) < hot_permille
    a["dest"] = np.where(is_hot, np.asarray(hot, dtype=np.uint32)[rng.integers(0, len(hot), size=k)],
                         rng.integers(0, n, size=k)).astype(np.uint32)
    a["body"]["type"] = O.T["broadcast"]
    a["body"]["flags"] = O.F_MSG_ID
    mid = np.zeros(k, dtype=np.uint32)
    for c in range(len(clients)):
        m = ci == c
        mid[m] = 1 + np.arange(int(m.sum()))
    a["body"]["msg_id"] = mid
    a["body"]["p0"] = i
    return a


@pytest.mark.gpu
def test_bench_topology_heavy_windows_vs_oracle(engine_backend):
    # The bench topology (4096 nodes, 64x64 grid) with windows in ALL size classes -- sender-block ordering with
    # its bitonic fallbacks, per-neighbor block claims -- against the oracle, journal bit for bit: 2 ticks x 1280
    # values = 30.8 M messages.  90 % of the values enter at three nodes (two of them diagonal neighbors), which
    # gives 23 726 windows of 513..2048 messages and 20 above 2048 (max 2342; counted on the oracle's journal).
    if engine_backend != "cuda":
        pytest.skip("61 M journal events: B200 only (the emulator covers the same paths at 16-25 nodes)")
    n, V, ticks = 4096, 1280, 2
    g, o = make_pair(n, topology="grid", n_values=V * ticks + 8, max_endpoints=n + 8, ring_cap=8192, max_window=4096,
                     journal_cap_log2=26, journal_level=1)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(4)]
        s.schedule(hot_broadcast_ops(n, cs, ticks, V, [10 * 64 + 10, 11 * 64 + 11, 40 * 64 + 50], 900, seed=77))
        s.run(ticks * 1_000_000)

    both(g, o, scenario)
    ev_g, _ = g.drain(bodies=False)
    ev_o, _ = o.journal()
    assert len(ev_g) == len(ev_o) == 2 * ticks * V * 12035
    for f in ("event_id", "time_ns", "msg_id", "src", "dest"):
        assert np.array_equal(ev_g[f], ev_o[f]), f
    assert g.stats() == o.stats() and g.round == o.round
    c = g.counters()
    assert c["max_window"] == 2342 and c["fallback_sorts"] > 0       # the big-window machinery did run


def test_echo_doc_counts():
    g, o = make_pair(1, workload="echo")

    def scenario(s, body):
        c0 = s.add_endpoint("c0")
        first = s.send(c0, 0, body("init", msg_id=1))
        r = s.recv(c0, 10_000_000_000)
        out = [first, int(r["type"]), int(r["id"]), int(r["msg_id"])]
        c1 = s.add_endpoint("c1")
        for i in range(12):
            s.send(c1, 0, body("echo", msg_id=i + 1, p0=i, p1=0xABCD0000 + i))
            r = s.recv(c1, 5_000_000_000)
            out.append((int(r["type"]), int(r["in_reply_to"]), int(r["p0"]), int(r["p1"]), int(r["msg_id"])))
        return out

    rg, ro = both(g, o, scenario)
    assert rg == ro
    assert_same_journal(g, o)
    assert g.stats()["all"] == {"send-count": 26, "recv-count": 26, "msg-count": 26}   # doc/02-echo/index.md:379-383


def test_many_values_duplicates_same_round():
    # several clients hit the same nodes with many values in the same round: exercises the
    # id sort, the first-sight hash and multi-chunk windows
    n = 25
    g, o = make_pair(n, topology="grid", n_values=4096, ring_cap=4096, max_window=2048, journal_cap_log2=21)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(4)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=3, per_tick=300, seed=11)
        s.schedule(ops)
        s.run(4_000_000)
        return nv

    both(g, o, scenario)
    assert_same_journal(g, o)
    for k in (0, 7, 24):
        assert g.node_set(k).tolist() == o.node_set(k).tolist() == list(range(900))


@pytest.mark.parametrize("dist,mean", [("constant", 1), ("constant", 10), ("uniform", 5), ("exponential", 5)])
def test_latency_distributions(dist, mean):
    n = 25
    g, o = make_pair(n, topology="grid", latency_dist=dist, latency_mean_ms=mean, n_values=256,
                     ring_cap=1024, max_window=512)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(2)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=20, per_tick=4, seed=5)
        s.schedule(ops)
        s.run(400_000_000)
        return nv

    both(g, o, scenario)
    ev, _ = assert_same_journal(g, o)
    assert len(ev) > 1000


def test_loss_flaky_slow_fast():
    n = 16
    g, o = make_pair(n, topology="grid", latency_dist="uniform", latency_mean_ms=3, n_values=512,
                     ring_cap=1024, max_window=512, p_loss=0.1)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(2)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=60, per_tick=3, seed=9)
        s.schedule(ops)
        s.run(15_000_000)
        s.flaky()
        s.run(30_000_000)
        s.slow()
        s.run(45_000_000)
        s.fast()
        s.set_loss(0.0)
        s.run(500_000_000)

    both(g, o, scenario)
    assert_same_journal(g, o)
    st = g.stats()
    assert st["all"]["send-count"] > st["all"]["recv-count"]


def test_partitions_drop_heal_bulk():
    n = 9
    g, o = make_pair(n, topology="grid", latency_dist="constant", latency_mean_ms=2, n_values=256,
                     ring_cap=512, max_window=256)

    def scenario(s, body):
        c = s.add_endpoint("c0")
        ops, nv = random_broadcast_ops(n, [c], n_ticks=40, per_tick=2, seed=3)
        s.schedule(ops)
        s.run(5_000_000)
        for a in range(0, 4):
            for b in range(4, 9):
                s.drop(a, b)
                s.drop(b, a)
        s.run(15_000_000)
        s.heal()
        s.run(25_000_000)
        s.partition([0, 0, 0, 1, 1, 1, 2, 2, 2])
        s.run(35_000_000)
        s.heal()
        s.run(200_000_000)

    both(g, o, scenario)
    assert_same_journal(g, o)
    assert g.counters()["partition_drops"] > 0


def test_errors_surface():
    import maelstrom_b200 as mb
    g = mb.Sim(4, topology="line", n_values=8, ring_cap=4, max_window=4)
    c = g.add_endpoint("c0")
    assert g.send(c, 99, mb.body("read", msg_id=1)) == -1          # node-not-found (net.clj:159-164)
    for v in range(8):
        g.send(c, 0, mb.body("broadcast", msg_id=v + 1, p0=v))     # 8 messages into a 4-slot ring
    with pytest.raises(mb.SimError):
        g.step(2)
    g.close()


def test_journal_backpressure_small_ring():
    # a 4096-record raw journal ring: the device stops running rounds whenever it is half full
    # (ms_run returns 1) and resumes after ms_journal_drain; the drained journal must be unchanged
    n = 25
    g, o = make_pair(n, topology="grid", n_values=2048, ring_cap=512, max_window=256, journal_cap_log2=12)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(3)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=40, per_tick=12, seed=21)
        s.schedule(ops)
        s.run(45_000_000)

    both(g, o, scenario)
    ev, _ = assert_same_journal(g, o)
    assert len(ev) > 40000


def test_sim_clients_and_block_path_counters():
    # simulated client sinks (bench configuration) + the fast ordering path must be taken
    from maelstrom_b200.engine import KIND_SIM_CLIENT
    n = 64
    g, o = make_pair(n, topology="grid", n_values=4096, ring_cap=2048, max_window=1024, journal_cap_log2=21)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i, KIND_SIM_CLIENT) for i in range(4)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=4, per_tick=400, seed=2)
        s.schedule(ops)
        s.run(6_000_000)
        return s.client_replies()

    rg, ro = both(g, o, scenario)
    assert rg == ro == 1600
    assert_same_journal(g, o)
    c = g.counters()
    assert c["fallback_sorts"] * 20 < c["rounds"] * (n + 4)   # bitonic fallback is the exception


@pytest.mark.parametrize("per_tick,max_window", [(1500, 2048), (3500, 4096)])
def test_large_windows_all_size_classes(per_tick, max_window):
    # windows of several thousand messages per node: exercises the 256- and 512-thread size
    # classes, multi-iteration emit loops and per-neighbor block claims with big counts
    n = 16
    g, o = make_pair(n, topology="grid", n_values=2 * per_tick + 8, ring_cap=4 * max_window,
                     max_window=max_window, journal_cap_log2=22, max_endpoints=n + 8)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(4)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=2, per_tick=per_tick, seed=31)
        s.schedule(ops)
        s.run(3_000_000)

    both(g, o, scenario)
    assert_same_journal(g, o)
    c = g.counters()
    assert c["max_window"] > 512


def test_drain_into_pinned_memory_in_pieces(engine_backend):
    if engine_backend != "cuda":
        pytest.skip("needs page-locked memory from the CUDA driver")
    # ms_journal_drain into a page-locked caller buffer (what bench.py's e2e leg does), in two
    # pieces that split a round; the concatenation must equal the oracle
    import torch
    n = 25
    g, o = make_pair(n, topology="grid", n_values=1024, ring_cap=1024, max_window=512, journal_cap_log2=20)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(2)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=5, per_tick=40, seed=4)
        s.schedule(ops)
        s.run(7_000_000)

    both(g, o, scenario)
    ev_o, _ = o.journal()
    total = len(ev_o)
    half = total // 2
    from maelstrom_b200._lib import EVENT_DTYPE
    pinned = torch.empty(half * 32, dtype=torch.uint8, pin_memory=True)
    got = g.drain_into(pinned.data_ptr(), half)
    assert got == half
    first = np.frombuffer(pinned.numpy().tobytes(), dtype=EVENT_DTYPE)
    rest, _ = g.drain(bodies=False)
    ev = np.concatenate([first, rest])
    assert len(ev) == total
    for f in ("event_id", "time_ns", "msg_id", "src", "dest"):
        assert np.array_equal(ev[f], ev_o[f]), f

