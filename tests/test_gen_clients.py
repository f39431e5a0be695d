"""Closed-loop clients on the device (SURVEY.md 8f NEXT-2): maelstrom.client's discipline
(client.clj:41-172) and the broadcast / g-set generator (workload/broadcast.clj:187-241,
core.clj:67-80) as a per-client state machine inside the round kernel.  History and journal must equal
the oracle's; on top, the checks Jepsen's set-full checker would make on such a history."""
import numpy as np
import pytest

import oracle_lib as O
from scenarios import assert_same_journal, both, make_pair

pytestmark = pytest.mark.usefixtures("engine_backend")
INVOKE, OK, FAIL, INFO, TIMEOUT = 0, 1, 2, 3, 0xFFFF
BROADCAST, READ = 0, 1


def run_pair(n, workload, n_clients, scenario, **kw):
    g, o = make_pair(n, workload=workload, **kw)
    out = both(g, o, scenario)
    hg, ho = g.history(), o.history()
    assert len(hg) == len(ho) > 0
    for f in ("time_ns", "order", "client", "op", "type", "f", "error", "value"):
        assert np.array_equal(hg[f], ho[f]), f
    ev, _ = assert_same_journal(g, o)
    return g, o, hg, ev, out


def check_client_discipline(h, n_clients):
    # one outstanding op per client, msg ids / op counters from 1, every invocation completed at most once
    for c in np.unique(h["client"]):
        mine = h[h["client"] == c]
        open_op = None
        for r in mine:
            if r["type"] == INVOKE:
                assert open_op is None, "client %d invoked op %d while op %d was outstanding" % (c, r["op"], open_op)
                open_op = int(r["op"])
            else:
                assert open_op == int(r["op"])
                open_op = None
        ops = mine[mine["type"] == INVOKE]["op"]
        assert ops.tolist() == list(range(1, len(ops) + 1))


def test_broadcast_generator_history_and_set_full():
    n, n_clients = 25, 10

    def scenario(s, body):
        c0 = s.add_gen_clients(n_clients, interval_ns=3_000_000, time_limit_ns=150_000_000, read_permille=500,
                               timeout_ns=40_000_000, quiet_ns=60_000_000, first_name=0)
        assert c0 == n
        s.run(260_000_000)
        return [s.node_set(k).tolist() for k in range(n)]

    g, o, h, ev, (sets_g, sets_o) = run_pair(n, "broadcast", n_clients, scenario, topology="grid", n_values=1 << 14,
                                            latency_dist="uniform", latency_mean_ms=2, ring_cap=512, max_window=256)
    assert sets_g == sets_o
    check_client_discipline(h, n_clients)
    inv = h[h["type"] == INVOKE]
    assert 300 < len(inv) < 900 and 0.3 < np.mean(inv["f"] == READ) < 0.7      # a 50/50 mix, staggered around 3 ms
    # broadcast values are unique across clients (k + n_clients * j)
    vals = inv[inv["f"] == BROADCAST]["value"]
    assert len(set(vals.tolist())) == len(vals)
    # every client ends with one final read after the quiet period (broadcast.clj:237-240)
    for c in range(n, n + n_clients):
        mine = h[h["client"] == c]
        last_inv = mine[mine["type"] == INVOKE][-1]
        assert last_inv["f"] == READ and last_inv["time_ns"] >= 150_000_000 + 60_000_000
        assert mine[-1]["type"] == OK and mine[-1]["op"] == last_inv["op"]
    # set-full: every acknowledged broadcast is in every node's set by the final reads; a final read's
    # value (the node's set, read back after the run) has the size the read_ok carried
    acked = set(h[(h["type"] == OK) & (h["f"] == BROADCAST)]["value"].tolist())
    assert len(acked) > 100
    for k in range(n):
        assert acked <= set(sets_g[k])
    for c in range(n, n + n_clients):
        final = h[(h["client"] == c) & (h["type"] == OK)][-1]
        assert int(final["value"]) == len(sets_g[(c - n) % n])


def test_timeouts_errors_and_stale_replies():
    # a partition cuts some clients from their nodes mid-run: requests time out (:info for broadcast,
    # :fail for read, client.clj:160-164), the late replies that arrive after healing are stale and
    # dropped (client.clj:106-107), and the client carries on with fresh msg ids
    n, n_clients = 9, 6

    def scenario(s, body):
        c0 = s.add_gen_clients(n_clients, interval_ns=2_000_000, time_limit_ns=120_000_000, read_permille=400,
                               timeout_ns=15_000_000, quiet_ns=30_000_000, first_name=5)
        s.run(30_000_000)
        comp = np.zeros(n + n_clients, dtype=np.uint32)
        comp[[0, 1, 2]] = 1                              # nodes 0-2 are cut off from everything else ...
        comp[n:] = 0                                     # ... including every client (clients are listed here)
        s.partition(comp)
        s.run(70_000_000)
        s.heal()
        s.run(200_000_000)
        return c0

    g, o, h, ev, _ = run_pair(n, "broadcast", n_clients, scenario, topology="grid", n_values=1 << 14,
                              latency_dist="constant", latency_mean_ms=1, ring_cap=512, max_window=256)
    check_client_discipline(h, n_clients)
    t = h[h["error"] == TIMEOUT]
    assert len(t) > 5
    assert set(t[t["f"] == BROADCAST]["type"].tolist()) <= {INFO} and set(t[t["f"] == READ]["type"].tolist()) <= {FAIL}
    assert g.counters()["partition_drops"] > 0
    # clients 0-2 talk to the cut nodes: only they time out; their completions carry on after the heal
    cut_clients = {n + k for k in range(n_clients) if k % n in (0, 1, 2)}
    assert set(t["client"].tolist()) <= cut_clients
    late = h[(h["time_ns"] > 75_000_000) & (h["type"] == OK)]
    assert cut_clients <= set(late["client"].tolist())


def test_gset_generator_and_errors():
    # g-set: add / read (workload/g_set.clj:59-61); a read of a node never errors, an `add` to a node that is
    # told to stop answering is covered by the timeout path above; here: definite error replies -> :fail
    n, n_clients = 5, 4

    def scenario(s, body):
        s.add_gen_clients(n_clients, interval_ns=4_000_000, time_limit_ns=90_000_000, read_permille=300,
                          timeout_ns=20_000_000, quiet_ns=40_000_000)
        cs = s.add_endpoint("c900")
        for i in range(n):
            s.send(cs, i, body("init", msg_id=1 + i))
        s.run(180_000_000)

    g, o, h, ev, _ = run_pair(n, "g-set", n_clients, scenario, n_values=1 << 12, gset_interval_ms=25, ring_cap=256, max_window=128)
    check_client_discipline(h, n_clients)
    assert set(h[h["type"] != INVOKE]["type"].tolist()) == {OK}
    adds = set(h[(h["type"] == OK) & (h["f"] == BROADCAST)]["value"].tolist())
    for k in range(n):
        assert adds <= set(g.node_set(k).tolist())       # everything acknowledged has replicated everywhere
