"""Pins the CPU oracle to every known-answer vector the reference holds for the
hot path: topology vectors and flood counts from the
tutorial, the echo message count, id/ordering facts visible in the docs, and
the published Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
import numpy as np
import pytest

import oracle_lib as O


# ------------------------------------------------------------------ RNG
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_latency_distributions():
    # net.clj:73-77: constant -> mean; uniform -> integers on [0, 2*mean); exponential rate 1/mean
    assert O.latency_draw("constant", 100, 1, [1, 2, 3, 4]) == 100
    assert O.latency_draw("constant", 100, 10, [1, 2, 3, 4]) == 1000          # slow! x10, net.clj:115-116
    rng = np.random.default_rng(1)
    xs = rng.integers(0, 2 ** 32, size=(20000, 4), dtype=np.uint64).astype(np.uint32)
    u = np.array([O.latency_draw("uniform", 100, 1, x) for x in xs])
    assert u.min() >= 0 and u.max() < 200 and abs(u.mean() - 99.5) < 2.0
    e = np.array([O.latency_draw("exponential", 100, 1, x) for x in xs])
    # truncated exponential: mean of floor(Exp(100)) = 1/(e^(1/100)-1) ~= 99.5
    assert abs(e.mean() - 99.5) < 3.0
    assert abs(np.median(e) - 69) <= 3
    # exact fixed-point identities: u = 1 -> 0 ms ; u = 2^-64 -> floor(mean*64*ln2)
    assert O.latency_draw("exponential", 100, 1, [0, 0xffffffff, 0xffffffff, 0]) == 0
    assert O.latency_draw("exponential", 100, 1, [0, 0, 0, 0]) == int(100 * 64 * np.log(2))
    # u = 1/2 exactly: X+1 = 2^63 -> -log2 u = 1 -> floor(100 ln 2) = 69
    assert O.latency_draw("exponential", 100, 1, [0, 0xffffffff, 0x7fffffff, 0]) == 69
    assert O.latency_draw("exponential", 100, 10, [0, 0xffffffff, 0x7fffffff, 0]) == 693


def test_loss_threshold():
    L = O.lib()
    assert L.or_loss_threshold(0.0) == 0
    assert L.or_loss_threshold(0.5) == 1 << 31          # flaky!, net.clj:121-122
    assert L.or_loss_threshold(1.0) == 1 << 32
    assert L.or_loss_threshold(0.1) == int(0.1 * 2 ** 32)


# ------------------------------------------------------------------ topologies
def test_grid_topology_doc_vector():
    # doc/03-broadcast/01-broadcast.md:302-306, nodes n1..n5 = indices 0..4
    want = {0: [3, 1], 1: [4, 2, 0], 2: [1], 3: [0, 4], 4: [1, 3]}
    for k, nb in want.items():
        assert O.topology("grid", 5, k) == nb


def test_single_node_grid():
    assert O.topology("grid", 1, 0) == []                # 01-broadcast.md:265  {"n1" []}


def test_line_total_tree():
    assert O.topology("line", 1, 0) == []                # broadcast.clj:72-73
    assert O.topology("line", 5, 0) == [1]
    assert O.topology("line", 5, 2) == [1, 3]
    assert O.topology("line", 5, 4) == [3]
    assert O.topology("total", 4, 2) == [0, 1, 3]
    # SURVEY appendix C derived vector, b=2, 5 nodes
    want = {0: [1, 2], 1: [0, 3, 4], 2: [0], 3: [1], 4: [1]}
    for k, nb in want.items():
        assert O.topology("tree2", 5, k) == nb
    assert O.topology("tree4", 25, 0) == [1, 2, 3, 4]
    assert O.topology("tree4", 25, 5) == [1, 21, 22, 23, 24]
    assert O.topology("tree4", 25, 6) == [1]


def test_grid_4096_degree_sum():
    deg = sum(len(O.topology("grid", 4096, k)) for k in range(4096))
    assert deg == 16128                                  # 16128 - 4095 = 12033


# ------------------------------------------------------------------ flood counts
def flood(topo, n, start=0, n_values=4):
    s = O.Sim(n, workload=O.W_BROADCAST, topology=topo, n_values=n_values)
    c = s.add_endpoint("c0", O.KIND_CLIENT)
    s.send(c, start, O.body("broadcast", msg_id=1, p0=0))
    s.run(2_000_000)
    st = s.stats()
    reply = s.recv(c)
    assert reply is not None and reply["type"] == O.T["broadcast_ok"] and reply["in_reply_to"] == 1
    for k in range(n):
        assert s.node_set(k).tolist() == [0]
    return st


@pytest.mark.parametrize("topo,n,want", [
    ("grid", 5, 6),        # 02-performance.md:71-76   2.94 msgs/op, 50% broadcasts
    ("grid", 25, 56),      # 02-performance.md:87-92   27.8 msgs/op
    ("line", 25, 24),      # 02-performance.md:110-115 12.0 msgs/op
    ("total", 25, 576),    # 02-performance.md:233-237 290.6 msgs/op
    ("tree4", 25, 24),     # 02-performance.md:249-254 11.99 msgs/op
    ("grid", 4096, 12033), # BASELINE.md closed form
])
def test_flood_counts(topo, n, want):
    st = flood(topo, n, start=n // 3)
    assert st["servers"]["send-count"] == want
    assert st["servers"]["recv-count"] == want
    assert st["servers"]["msg-count"] == want
    assert st["clients"] == {"send-count": 2, "recv-count": 2, "msg-count": 2}


def test_doc_rates_from_flood_counts():
    # The doc's msgs-per-op are flood counts x the observed broadcast fraction
    # (998 broadcasts / 1991 ops, 02-performance.md:33-50).
    frac = 998 / 1991
    for per_value, doc in ((6, 2.94182), (56, 27.806324), (24, 12.005973), (576, 290.6)):
        assert abs(per_value * frac - doc) / doc < 0.03


# ------------------------------------------------------------------ echo
def test_echo_doc_counts_and_ids():
    # doc/02-echo/index.md:379-383: 12 ops on one node = 26 messages, all client.
    s = O.Sim(1, workload=O.W_ECHO)
    c0 = s.add_endpoint("c0", O.KIND_CLIENT)
    first = s.send(c0, 0, O.body("init", msg_id=1))
    assert first == 0                                    # :id=>0, 01-broadcast.md:246; net.clj:103
    r = s.recv(c0, 10_000_000_000)
    assert r["type"] == O.T["init_ok"] and r["in_reply_to"] == 1 and r["id"] == 1
    c1 = s.add_endpoint("c1", O.KIND_CLIENT)
    for i in range(12):
        s.send(c1, 0, O.body("echo", msg_id=i + 1, p0=i, p1=0xABCD0000 + i))
        r = s.recv(c1, 5_000_000_000)
        assert r["type"] == O.T["echo_ok"] and r["in_reply_to"] == i + 1
        assert r["p0"] == i and r["p1"] == 0xABCD0000 + i
        assert r["msg_id"] == i + 2                      # echo.rb:12 per-node counter (init_ok was 1)
    st = s.stats()
    assert st["all"] == {"send-count": 26, "recv-count": 26, "msg-count": 26}
    assert st["clients"] == st["all"]
    assert st["servers"] == {"send-count": 0, "recv-count": 0, "msg-count": 0}
    ev, _ = s.journal()
    assert (ev["event_id"] & ~np.uint64(O.RECV_BIT)).tolist() == list(range(52))   # journal.clj:321-337 dense
    assert sorted(set(ev["msg_id"].tolist())) == list(range(26))                   # journal.clj:264-274 dense


# ------------------------------------------------------------------ net rules
def test_client_latency_zero_server_latency_applies():
    # net.clj:185-186 + 02-performance.md:185-195: 100 ms/hop between servers only.
    s = O.Sim(25, topology="grid", latency_dist="constant", latency_mean_ms=100, n_values=4)
    c = s.add_endpoint("c0")
    s.send(c, 0, O.body("broadcast", msg_id=1, p0=0))
    s.run(2_000_000_000)
    ev, bd = s.journal()
    recv = ev[(ev["event_id"] & np.uint64(O.RECV_BIT)) != 0]
    cl = recv[recv["dest"] == c]
    assert cl["time_ns"].tolist() == [0]                 # reply to the client is immediate
    sv = recv[recv["dest"] < 25]
    # corner-to-corner on a 5x5 grid is 8 hops: last first-delivery at 800 ms
    first = {}
    for e in sv:
        first.setdefault(int(e["dest"]), int(e["time_ns"]))
    assert max(first.values()) == 800_000_000 and first[0] == 0 and len(first) == 25
    assert set((sv["time_ns"] % 100_000_000).tolist()) == {0}


def test_send_journaled_even_if_lost_and_partition_at_dequeue():
    # net.clj:207-215 (journal before the loss roll), net.clj:234 (partition at recv,
    # no :recv event); 02-performance.md:519-529 shows send-count > recv-count.
    s = O.Sim(5, topology="grid", n_values=64, p_loss=0.5)
    c = s.add_endpoint("c0")
    for v in range(32):
        s.send(c, v % 5, O.body("broadcast", msg_id=v + 1, p0=v))
    s.run(5_000_000)
    st = s.stats()
    assert st["all"]["send-count"] > st["all"]["recv-count"] > 0
    # partition: message sent before drop! but dequeued during it is cut
    s2 = O.Sim(2, topology="line", latency_dist="constant", latency_mean_ms=10, n_values=4)
    c = s2.add_endpoint("c0")
    s2.send(c, 0, O.body("broadcast", msg_id=1, p0=0))
    s2.run(5_000_000)                                    # n0 has forwarded to n1, in flight
    s2.drop(0, 1)                                        # n1 drops packets from n0
    s2.run(50_000_000)
    assert s2.node_set(1).tolist() == []
    assert s2.stats()["servers"] == {"send-count": 1, "recv-count": 0, "msg-count": 1}
    s2.heal()
    s2.send(c, 0, O.body("broadcast", msg_id=2, p0=1))
    s2.run(100_000_000)
    assert s2.node_set(1).tolist() == [1]


def test_unknown_dest_and_removed_endpoint():
    s = O.Sim(2, topology="line")
    c = s.add_endpoint("c0")
    assert s.send(c, 99, O.body("read", msg_id=1)) == -1   # node-not-found, net.clj:159-164
    s.remove_endpoint(c)
    assert s.send(c, 0, O.body("read", msg_id=1)) == -1


def test_read_snapshot_and_unknown_type_error():
    s = O.Sim(3, topology="line", n_values=16)
    c = s.add_endpoint("c0")
    s.send(c, 1, O.body("broadcast", msg_id=1, p0=7))
    s.run(1_000_000)
    s.recv(c)
    s.send(c, 2, O.body("read", msg_id=2))
    r = s.recv(c, 1_000_000_000)
    assert r["type"] == O.T["read_ok"] and r["p0"] == 1
    assert s.read_snapshot(int(r["id"])).tolist() == [7]
    s.send(c, 2, O.body("add", msg_id=3, p0=1))            # not a broadcast-node handler
    r = s.recv(c, 1_000_000_000)
    assert r["type"] == O.T["error"] and r["p0"] == 10 and r["in_reply_to"] == 3   # errors.edn code 10


# ------------------------------------------------------------------ g-set (oracle only; engine: next round)
def gset_cluster(n=5, **kw):
    s = O.Sim(n, workload=O.W_GSET, n_values=256, **kw)
    for i in range(n):                                   # db.clj:46-69: one fresh client per node sends init
        c = s.add_endpoint("c%d" % i)
        s.send(c, i, O.body("init", msg_id=1))
    return s


def test_gset_full_state_replication_doc_count():
    # doc/04-crdts/01-g-set.md:200-210: 5 nodes replicating the whole set every 5 s exchange
    # 80 inter-server messages over the 20 s of the run (4 rounds of 5 x 4 messages)
    s = gset_cluster()
    s.run(20_000_000_000)
    st = s.stats()
    assert st["servers"] == {"send-count": 80, "recv-count": 80, "msg-count": 80}
    assert st["clients"]["send-count"] == 10            # 5 init + 5 init_ok


def test_gset_add_read_and_convergence():
    s = gset_cluster()
    s.run(1_000_000)
    c = s.add_endpoint("c9")
    for k, v in enumerate((3, 7, 11)):
        s.send(c, k, O.body("add", msg_id=k + 1, p0=v))
        r = s.recv(c, 1_000_000_000)
        assert r["type"] == O.T["add_ok"] and r["in_reply_to"] == k + 1
    s.send(c, 4, O.body("read", msg_id=10))
    r = s.recv(c, 1_000_000_000)
    assert r["type"] == O.T["read_ok"] and r["p0"] == 0   # nothing replicated yet
    s.run(5_100_000_000)                                  # next replication round at t = 5 s
    for k in range(5):
        assert s.node_set(k).tolist() == [3, 7, 11]
    s.send(c, 4, O.body("read", msg_id=11))
    r = s.recv(c, 1_000_000_000)
    assert r["p0"] == 3 and s.read_snapshot(int(r["id"])).tolist() == [3, 7, 11]


def test_gset_converges_despite_loss_and_partition():
    # CRDT property the workload is about (workload/g_set.clj:52-62 set-full): every add is
    # eventually visible everywhere once the network heals
    s = gset_cluster(latency_dist="uniform", latency_mean_ms=20)
    s.run(100_000_000)                                    # every node initialised (init is not exempt from loss)
    s.set_loss(0.3)
    c = s.add_endpoint("c9")
    for v in range(20):
        s.send(c, v % 5, O.body("add", msg_id=v + 1, p0=v))
    s.partition([0, 0, 1, 1, 1])
    s.run(12_000_000_000)
    assert any(len(s.node_set(k)) < 20 for k in range(5))
    s.heal()
    s.set_loss(0.0)
    s.run(30_000_000_000)
    sets = [s.node_set(k).tolist() for k in range(5)]
    added = sorted(set(sum(sets, [])))
    assert all(x == added for x in sets) and len(added) >= 10   # lost client adds never reached a node
