"""The host-side mirror of maelstrom.net / maelstrom.client (maelstrom_b200/net.py, client.py),
driven the way the reference's tutorial drives the real thing (doc/02-echo, doc/03-broadcast).
CPU: over the oracle backend and over the engine's kernels on the SIMT emulator; GPU: the same
scenarios over the CUDA engine."""
import json

import pytest

import oracle_lib as O
from maelstrom_b200 import client as C
from maelstrom_b200 import errors
from maelstrom_b200.net import Net, NodeNotFound, to_wire


def oracle_backend(n, workload, **kw):
    return Net(O.Sim(n, workload=workload, **kw), O.body)


def engine_backend(n, workload, **kw):
    import maelstrom_b200 as mb
    name = {O.W_ECHO: "echo", O.W_BROADCAST: "broadcast"}[workload]
    return Net(mb.Sim(n, workload=name, **kw), mb.body)


def emul_backend(n, workload, **kw):
    # the engine's kernel sources on the CPU SIMT emulator (tests/native/emul)
    import emul_lib
    with emul_lib.use():
        return engine_backend(n, workload, **kw)


BACKENDS = [pytest.param(oracle_backend, id="oracle"),
            pytest.param(emul_backend, id="emul"),
            pytest.param(engine_backend, id="cuda", marks=pytest.mark.gpu)]


def init_node(net, node_id, node_ids):
    """db.clj:46-69: a fresh client sends init and expects init_ok within 10 s."""
    c = C.Client(net)
    body = c.rpc(node_id, {"type": "init", "node_id": node_id, "node_ids": node_ids}, 10_000)
    assert body["type"] == "init_ok"
    c.close()
    return body


@pytest.mark.parametrize("backend", BACKENDS)
def test_echo_tutorial(backend):
    net = backend(1, O.W_ECHO)
    body = init_node(net, "n0", ["n0"])
    # demo/go/node_test.go:102  {"src":"n3","body":{"in_reply_to":1,"type":"init_ok"}}
    golden = json.loads('{"src":"n3","body":{"in_reply_to":1,"type":"init_ok"}}')["body"]
    assert {k: body[k] for k in golden} == golden
    c = C.Client(net)
    assert c.node_id == "c1"                                          # c0 was the init client (client.clj:48)
    for i in range(12):                                               # doc/02-echo/index.md: 12 ops
        text = "Please echo %d" % (i * 7 % 128)                       # workload/echo.clj:72-75
        reply = c.rpc("n0", {"type": "echo", "echo": text})
        assert reply["type"] == "echo_ok" and reply["echo"] == text
        assert reply["in_reply_to"] == i + 1 and reply["msg_id"] == i + 2    # echo.rb:12-13
    st = net.sim.stats()
    assert st["all"] == {"send-count": 26, "recv-count": 26, "msg-count": 26}     # index.md:379-383
    assert st["servers"]["msg-count"] == 0


@pytest.mark.parametrize("backend", BACKENDS)
def test_broadcast_tutorial_five_nodes(backend):
    nodes = ["n%d" % i for i in range(5)]
    net = backend(5, O.W_BROADCAST, topology="grid", n_values=64)
    for n in nodes:
        init_node(net, n, nodes)
    clients = {n: C.Client(net) for n in nodes}
    for n, c in clients.items():                                      # workload/broadcast.clj:195-197
        assert c.rpc(n, {"type": "topology", "topology": {}})["type"] == "topology_ok"
    before = net.sim.stats()["servers"]["send-count"]
    for v in range(10):                                               # broadcast, then read (50/50 mix)
        n = nodes[v % 5]
        assert clients[n].rpc(n, {"type": "broadcast", "message": v})["type"] == "broadcast_ok"
    for n, c in clients.items():
        reply = c.rpc(n, {"type": "read"})
        assert reply["type"] == "read_ok" and reply["messages"] == list(range(10))
    # 6 server messages per broadcast on the 5-node grid with skip-sender (02-performance.md:71-76)
    assert net.sim.stats()["servers"]["send-count"] - before == 60
    # the envelope a node would have seen for the first client request (process.clj:162)
    w = to_wire({"id": 0, "src": "c0", "dest": "n0", "body": {"type": "init", "msg_id": 1}})
    assert json.loads(json.dumps(w))["body"]["type"] == "init"


@pytest.mark.parametrize("backend", BACKENDS)
def test_errors_and_timeouts(backend):
    net = backend(3, O.W_BROADCAST, topology="line", n_values=16)
    c = C.Client(net)
    with pytest.raises(AssertionError):                               # net.clj:174-175
        net.send({"src": c.node_id, "dest": "n99", "body": {"type": "read", "msg_id": 1}})
    with pytest.raises(NodeNotFound):                                 # net.clj:159-164
        net.recv("nobody", 10)
    # unsupported request -> error 10, definite (errors.edn; demo/go/node_test.go:51)
    with pytest.raises(C.RPCError) as e:
        c.rpc("n1", {"type": "add", "element": 3})
    assert e.value.code == 10 and e.value.name == "not-supported" and e.value.definite
    op = {"f": "add", "value": 3}
    assert C.with_errors(op, {"read"}, lambda: c.rpc("n1", {"type": "add", "element": 3}))["type"] == "fail"
    # a partition between the client and the node: the request is cut at dequeue (net.clj:234),
    # the client times out (client.clj:96-101)
    # (400 ms here instead of the 5000 ms default, client.clj:18-20: idle virtual ticks are rounds)
    net.drop(None, c.node_id, "n1")
    t0 = net.sim.now
    with pytest.raises(C.Timeout):
        c.rpc("n1", {"type": "read"}, 400)
    assert net.sim.now - t0 >= 400_000_000
    assert C.DEFAULT_TIMEOUT_MS == 5000
    assert C.with_errors({"f": "broadcast"}, {"read"}, lambda: c.rpc("n1", {"type": "read"}, 400))["type"] == "info"
    net.heal()
    # the stale reply rule: a late answer to an abandoned request is discarded (client.clj:106-107)
    assert c.rpc("n1", {"type": "read"})["type"] == "read_ok"
    c.close()
    with pytest.raises(AssertionError):
        net.send({"src": c.node_id, "dest": "n1", "body": {"type": "read", "msg_id": 9}})


def test_error_registry_matches_errors_edn():
    # resources/errors.edn:2-44: codes, names and the definite? flag (all but 0 and 13)
    assert sorted(errors.ERRORS) == [0, 1, 10, 11, 12, 13, 14, 20, 21, 22, 30]
    assert [c for c in errors.ERRORS if not errors.definite(c)] == [0, 13]
    assert errors.name(22) == "precondition-failed" and errors.name(30) == "txn-conflict"
    assert errors.name(999) == "unknown" and not errors.definite(999)


# ---------------------------------------------------------------- round 2: reads of the other workloads
def _backend_for(kind, n, workload_code, **kw):
    import maelstrom_b200 as mb
    name = {O.W_GSET: "g-set", O.W_RAFT: "lin-kv", O.W_ECHO: "echo", O.W_BROADCAST: "broadcast"}[workload_code]
    if kind == "oracle":
        return Net(O.Sim(n, workload=workload_code, **kw), O.body)
    if kind == "emul":
        import emul_lib
        with emul_lib.use():
            return Net(mb.Sim(n, workload=name, **kw), mb.body)
    return Net(mb.Sim(n, workload=name, **kw), mb.body)


KINDS = [pytest.param("oracle"), pytest.param("emul"), pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("kind", KINDS)
def test_gset_read_returns_value(kind):
    # demo/ruby/g_set.rb:13-15: read_ok {value: @set.to_a}; workload/g_set.clj:20-26
    net = _backend_for(kind, 3, O.W_GSET, n_values=64, gset_interval_ms=50)
    nodes = ["n0", "n1", "n2"]
    for n in nodes:
        init_node(net, n, nodes)
    c = C.Client(net)
    for v in (5, 9):
        assert c.rpc("n0", {"type": "add", "element": v})["type"] == "add_ok"
    r = c.rpc("n0", {"type": "read"})
    assert r["type"] == "read_ok" and r["value"] == [5, 9] and "messages" not in r
    net.sim.run(net.sim.now + 120_000_000)                            # two replication periods later
    r = c.rpc("n2", {"type": "read"})
    assert r["value"] == [5, 9]


@pytest.mark.parametrize("kind", KINDS)
def test_lin_kv_read_returns_value(kind):
    # demo/python/raft.py:158-192: read_ok {value}; workload/lin_kv.clj:12-38
    net = _backend_for(kind, 3, O.W_RAFT)
    nodes = ["n0", "n1", "n2"]
    for n in nodes:
        init_node(net, n, nodes)
    net.sim.run(5_000_000_000)                                        # an election has happened by now
    c = C.Client(net)
    assert c.rpc("n1", {"type": "write", "key": 1, "value": 42})["type"] == "write_ok"
    r = c.rpc("n2", {"type": "read", "key": 1})                       # proxied to the leader (raft.py:562-566)
    assert r["type"] == "read_ok" and r["value"] == 42 and "messages" not in r
    assert c.rpc("n0", {"type": "cas", "key": 1, "from": 42, "to": 43})["type"] == "cas_ok"
    assert c.rpc("n0", {"type": "read", "key": 1})["value"] == 43
    with pytest.raises(C.RPCError) as e:
        c.rpc("n0", {"type": "read", "key": 7})
    assert e.value.code == 20


@pytest.mark.parametrize("kind", KINDS)
def test_late_reply_to_a_closed_client_and_client_churn(kind, caplog):
    # Jepsen closes and reopens a client after every timeout (client.clj:55-59): the node's late
    # reply goes to a name that is gone, which must not stop the network, and 300 reopenings must
    # not exhaust the endpoint table (max_endpoints defaults to n_nodes + 256)
    import logging
    net = _backend_for(kind, 3, O.W_ECHO)
    net.log_send = net.log_recv = True
    with caplog.at_level(logging.INFO, logger="maelstrom.net"):
        c = C.Client(net)
        c.send({"dest": "n0", "body": {"type": "echo", "echo": "x"}})
        c.close()                                                     # before n0 has answered
        net.sim.run(3_000_000)
        assert net.sim.undeliverable() == 1
        for i in range(300):
            c = C.Client(net)
            assert c.rpc("n%d" % (i % 3), {"type": "echo", "echo": i})["echo"] == i
            c.close()
    assert any(":send" in r.getMessage() for r in caplog.records)      # net.clj:211
    assert any(":recv" in r.getMessage() for r in caplog.records)      # net.clj:241
