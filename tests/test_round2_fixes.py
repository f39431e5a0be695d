"""Round-2 regression tests for behaviours the round-1 review found broken:
  * slow! with a non-constant latency law must only make messages slower (net.clj:115-116), never
    kill the run: latencies beyond the timing wheel's span wait extra turns ("laps") in their slot;
  * a send to an endpoint that has been removed (a reply to a closed client, gossip to a stopped
    node) is a per-message drop, not a fatal error (the reference's assert, net.clj:166-176, throws
    only in the sending node's own thread, process.clj:148-150);
  * endpoint slots of closed clients are recycled (client.clj:55-59: Jepsen reopens clients after
    every indefinite op);
  * ms_schedule_ops rejects a bad batch as a whole.
Every scenario runs on the oracle and on the engine ([emul] in the CPU suite, [cuda] on the B200)
and the journals must be identical."""
import numpy as np
import pytest

import oracle_lib as O
from scenarios import assert_same_journal, both, make_pair, ops_array, random_broadcast_ops

pytestmark = pytest.mark.usefixtures("engine_backend")


@pytest.mark.parametrize("dist,mean,slots", [("exponential", 3, 16), ("constant", 7, 4)])
def test_slow_with_latency_beyond_the_wheel(dist, mean, slots):
    # 25-node grid, slow! twice (x100): latencies of hundreds of ticks on a wheel of `slots` slots
    n = 25
    g, o = make_pair(n, topology="grid", latency_dist=dist, latency_mean_ms=mean, n_values=256,
                     ring_cap=1024, max_window=512, calendar_slots=slots, calendar_cap=256)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(2)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=30, per_tick=3, seed=11)
        s.schedule(ops)
        s.run(10_000_000)
        s.slow()
        s.run(40_000_000)
        s.slow()
        s.run(400_000_000)
        s.fast()
        s.fast()
        s.run(4_500_000_000)
        return nv

    both(g, o, scenario)
    ev, _ = assert_same_journal(g, o)
    assert len(ev) > 2000 and g.now == 4_500_000_000
    for k in (0, 12, 24):
        assert g.node_set(k).tolist() == o.node_set(k).tolist()
    # everything sent was delivered in the end (no loss configured)
    st = g.stats()["all"]
    assert st["send-count"] == st["recv-count"]


def test_wheel_pool_exhaustion_is_reported():
    import maelstrom_b200 as mb
    g = mb.Sim(9, topology="total", latency_dist="constant", latency_mean_ms=50, n_values=4096,
               ring_cap=1024, max_window=512, calendar_slots=64, calendar_cap=16)
    c = g.add_endpoint("c0")
    for v in range(1500):                        # 1500 x 8 gossip messages in flight >> 64 x 16 + slack
        g.send(c, v % 9, mb.body("broadcast", msg_id=v + 1, p0=v))
    with pytest.raises(mb.SimError) as e:
        g.run(20_000_000)
    assert "timing wheel" in str(e.value)
    g.close()


def test_reply_to_a_closed_client_is_dropped_not_fatal():
    g, o = make_pair(3, workload="echo")

    def scenario(s, body):
        c0 = s.add_endpoint("c0")
        c1 = s.add_endpoint("c1")
        s.send(c0, 0, body("echo", msg_id=1, p1=77))
        s.step(1)                                  # the request is on its way to n0
        s.remove_endpoint(c0)                      # client/close! (client.clj:55-59) before the reply is sent
        s.run(2_000_000)                           # n0 answers a client that is gone
        s.send(c1, 1, body("echo", msg_id=1, p1=5))
        r = s.recv(c1, 1_000_000_000)              # the network is still alive
        return int(r["type"]), int(r["p1"]), s.undeliverable()

    rg, ro = both(g, o, scenario)
    assert rg == ro == (O.T["echo_ok"], 5, 1)
    ev, _ = assert_same_journal(g, o)
    # the dropped reply has a :send event and no :recv (ids stay dense: net.clj:197 runs before the assert)
    assert int(np.sum(ev["event_id"] >> 63 == 0)) == int(np.sum(ev["event_id"] >> 63 == 1)) + 1


def test_gossip_to_stopped_nodes_and_removed_hosts():
    n = 9
    g, o = make_pair(n, topology="grid", latency_dist="constant", latency_mean_ms=3, n_values=64,
                     ring_cap=256, max_window=128)

    def scenario(s, body):
        c = s.add_endpoint("c0")
        h = s.add_endpoint("h0", O.KIND_HOST)       # a non-client host endpoint: server latency applies
        for v in range(6):
            s.send(c, v, body("broadcast", msg_id=v + 1, p0=v))
        s.send(h, 4, body("read", msg_id=1))        # read_ok travels 3 ms back to h0 ...
        s.run(4_000_000)
        s.remove_endpoint(h)                        # ... and h0 leaves while it is in the wheel
        h2 = s.add_endpoint("h1", O.KIND_HOST)      # recycles the slot: must not inherit h0's mail
        assert h2 == h
        s.run(30_000_000)
        assert s.recv(h2, 0) is None
        s.send(h2, 4, body("read", msg_id=7))
        r = s.recv(h2, 1_000_000_000)
        return int(r["in_reply_to"]), int(r["p0"])

    rg, ro = both(g, o, scenario)
    assert rg == ro == (7, 6)
    assert_same_journal(g, o)


def test_schedule_rejects_a_bad_batch_as_a_whole():
    import maelstrom_b200 as mb
    n = 9
    g, o = make_pair(n, topology="grid", n_values=64)
    cg, co = g.add_endpoint("c0"), o.add_endpoint("c0")
    good1 = ops_array([(0, cg, 1, "broadcast", 1, 1), (1_000_000, cg, 2, "broadcast", 2, 2)])
    unsorted = ops_array([(5_000_000, cg, 3, "broadcast", 3, 3), (2_000_000, cg, 4, "broadcast", 4, 4)])
    unknown = ops_array([(6_000_000, cg, 3, "broadcast", 3, 3), (7_000_000, cg, 999, "broadcast", 4, 4)])
    good2 = ops_array([(3_000_000, cg, 5, "broadcast", 5, 5), (3_000_000, cg, 6, "broadcast", 6, 6)])
    g.schedule(good1)
    with pytest.raises(mb.SimError):
        g.schedule(unsorted)
    with pytest.raises(mb.SimError):
        g.schedule(unknown)
    g.schedule(good2)
    o.schedule(good1)
    o.schedule(good2)
    g.run(8_000_000)
    o.run(8_000_000)
    assert_same_journal(g, o)
    assert g.node_set(0).tolist() == [1, 2, 5, 6]


def test_many_incremental_schedules_prefix_table():
    # tick_off is rebuilt with one counting pass per call; 40 small appends
    n = 16
    g, o = make_pair(n, topology="grid", n_values=1024)
    cg, co = g.add_endpoint("c0"), o.add_endpoint("c0")
    rng = np.random.default_rng(3)
    t, mid = 0, 0
    for _ in range(40):
        rows = []
        for _ in range(int(rng.integers(1, 4))):
            t += int(rng.integers(0, 20)) * 1_000_000 + int(rng.integers(0, 2)) * 500
            mid += 1
            rows.append((t, cg, int(rng.integers(n)), "broadcast", mid, mid % 1024))
        a = ops_array(rows)
        g.schedule(a)
        o.schedule(a)
    g.run(t + 3_000_000)
    o.run(t + 3_000_000)
    assert_same_journal(g, o)


@pytest.mark.parametrize("latency", [0, 2])
def test_three_phase_commit_with_many_tickets(latency):
    # more than 16384 tickets: the round is committed by k_commit_a/b/c instead of the last ticket
    n = 25
    g, o = make_pair(n, topology="grid", n_values=512, max_endpoints=17000, ring_cap=256, max_window=256,
                     latency_dist="constant", latency_mean_ms=latency, journal_cap_log2=18)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i) for i in range(3)]
        ops, nv = random_broadcast_ops(n, cs, n_ticks=12, per_tick=6, seed=4)
        s.schedule(ops)
        s.run((14 + 60 * latency) * 1_000_000)

    both(g, o, scenario)
    ev, _ = assert_same_journal(g, o)
    assert len(ev) > 3000
