"""g-set workload (demo/ruby/g_set.rb:13-39): the engine's g-set node
program -- add / read / replicate_one / replicate_full and the `every 5 s` full-state replication
task -- against the oracle, journal bit for bit.  [emul] runs the kernel sources on the CPU SIMT
emulator (tests/native/emul); [cuda] is the same scenario on a B200."""
import numpy as np
import pytest

import oracle_lib as O
from scenarios import assert_same_journal, both, make_pair, ops_array

pytestmark = pytest.mark.usefixtures("engine_backend")


def init_all(s, body, n):
    """db.clj:46-69: one fresh client per node sends init and waits for init_ok."""
    cs = []
    for i in range(n):
        c = s.add_endpoint("c%d" % i)
        cs.append(c)
        s.send(c, i, body("init", msg_id=1))
    return cs


def test_doc_message_count_five_nodes():
    # doc/04-crdts/01-g-set.md:200-210: 5 nodes, full state every interval: 4 runs x 5 x 4 = 80
    # inter-server messages (interval shortened from 5 s to 50 ms, same number of runs)
    g, o = make_pair(5, workload="g-set", n_values=256, gset_interval_ms=50, ring_cap=64, max_window=64)

    def scenario(s, body):
        init_all(s, body, 5)
        s.run(200_000_000)

    both(g, o, scenario)
    assert_same_journal(g, o)
    assert g.stats()["servers"] == {"send-count": 80, "recv-count": 80, "msg-count": 80}


def test_add_read_replicate_and_error_replies():
    g, o = make_pair(5, workload="g-set", n_values=256, gset_interval_ms=40, ring_cap=64, max_window=64)

    def scenario(s, body):
        init_all(s, body, 5)
        s.run(1_000_000)
        c = s.add_endpoint("c9")
        out = []
        for k, v in enumerate((3, 7, 11)):
            s.send(c, k, body("add", msg_id=k + 1, p0=v))
            r = s.recv(c, 1_000_000_000)
            out.append((int(r["type"]), int(r["in_reply_to"]), int(r["id"])))
        s.send(c, 4, body("read", msg_id=10))
        r = s.recv(c, 1_000_000_000)
        out.append((int(r["type"]), int(r["p0"])))                  # nothing replicated yet: 0
        s.send(c, 4, body("replicate_one", p0=99))                   # g_set.rb:24-26: no reply
        s.send(c, 2, body("broadcast", msg_id=12, p0=1))             # no such handler: error 10
        r = s.recv(c, 1_000_000_000)
        out.append((int(r["type"]), int(r["p0"]), int(r["in_reply_to"])))
        s.run(100_000_000)
        s.send(c, 0, body("read", msg_id=13))
        r = s.recv(c, 1_000_000_000)
        out.append((int(r["type"]), int(r["p0"])))
        out.append([s.node_set(k).tolist() for k in range(5)])
        return out

    rg, ro = both(g, o, scenario)
    assert rg == ro
    assert rg[3] == (O.T["read_ok"], 0) and rg[4] == (O.T["error"], 10, 12) and rg[5] == (O.T["read_ok"], 4)
    assert rg[6] == [[3, 7, 11, 99]] * 5
    assert_same_journal(g, o)


def scheduled_adds_and_reads(n, n_clients, n_ticks, per_tick, seed, read_every=5):
    """ops from n_clients simulated clients: adds of fresh elements, every read_every-th op a read"""
    rng = np.random.default_rng(seed)
    rows, mid, v = [], [0] * n_clients, 0
    for t in range(n_ticks):
        for k in range(per_tick):
            c = int(rng.integers(n_clients))
            mid[c] += 1
            if (t * per_tick + k) % read_every == read_every - 1:
                rows.append((t * 1_000_000, n + c, int(rng.integers(n)), "read", mid[c], 0))
            else:
                rows.append((t * 1_000_000, n + c, int(rng.integers(n)), "add", mid[c], v))
                v += 1
    return ops_array(rows), v


@pytest.mark.parametrize("n,dist,mean,interval", [
    (16, "constant", 0, 7),        # merges, adds and reads meet in the same windows
    (40, "constant", 3, 10),       # timing wheel
    (70, "uniform", 4, 9),         # > 64 single-message sender blocks per window: bitonic ordering path
    (30, "exponential", 5, 16),
])
def test_mixed_traffic_reads_see_their_place_in_the_sequence(n, dist, mean, interval):
    g, o = make_pair(n, workload="g-set", n_values=2048, gset_interval_ms=interval, latency_dist=dist,
                     latency_mean_ms=mean, ring_cap=512, max_window=512, journal_cap_log2=20,
                     max_endpoints=n + 8)

    def scenario(s, body):
        cs = [s.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(4)]
        for i in range(n):                                            # init from the first sim client
            s.send(cs[0], i, body("init", msg_id=1000 + i))
        ops, nv = scheduled_adds_and_reads(n, 4, n_ticks=40, per_tick=6, seed=n)
        s.schedule(ops)
        s.run(120_000_000)
        return [len(s.node_set(k)) for k in range(n)]

    rg, ro = both(g, o, scenario)
    assert rg == ro and max(rg) > 0
    ev, bd = assert_same_journal(g, o)
    reads = bd[(bd["type"] == O.T["read_ok"])]
    assert len(reads) > 0 and reads["p0"].max() > 0                   # some read saw replicated state


def test_loss_partition_and_convergence():
    # workload/g_set.clj:52-62 (set-full): every acknowledged add is eventually everywhere
    n = 12
    g, o = make_pair(n, workload="g-set", n_values=512, gset_interval_ms=20, latency_dist="uniform",
                     latency_mean_ms=3, ring_cap=256, max_window=256, max_endpoints=2 * n + 8)

    def scenario(s, body):
        cs = init_all(s, body, n)
        s.run(10_000_000)
        s.set_loss(0.3)
        c = s.add_endpoint("c99")
        for v in range(30):
            s.send(c, v % n, body("add", msg_id=v + 1, p0=v))
        s.partition([0] * 5 + [1] * 7 + [0xFFFFFFFF] * (len(cs) + 1))
        s.run(150_000_000)
        mid = [len(s.node_set(k)) for k in range(n)]
        s.heal()
        s.set_loss(0.0)
        s.slow()
        s.run(600_000_000)
        s.fast()
        return mid, [s.node_set(k).tolist() for k in range(n)]

    (mg, sg), (mo, so) = both(g, o, scenario)
    assert mg == mo and sg == so
    assert all(x == sg[0] for x in sg) and len(sg[0]) >= 10
    assert_same_journal(g, o)


def test_out_of_range_element_is_an_error():
    g, _ = make_pair(3, workload="g-set", n_values=64, gset_interval_ms=10)
    import maelstrom_b200 as mb
    c = g.add_endpoint("c0")
    g.send(c, 0, mb.body("add", msg_id=1, p0=64))
    with pytest.raises(mb.SimError):
        g.run(3_000_000)


def test_committed_golden_journal():
    # tests/golden/journals.json["gset_five_nodes"], generated from the oracle
    import json
    import os
    import golden_cases as G
    import maelstrom_b200 as mb
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "journals.json")))
    g = G.make_engine("gset_five_nodes")
    G.CASES["gset_five_nodes"][1](g, mb.body)
    ev, bd = g.drain()
    assert G.digest(ev, bd, g.stats(), g.now, g.round) == want["gset_five_nodes"]
