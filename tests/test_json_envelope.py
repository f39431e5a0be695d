"""ms_send_json / ms_recv_json: the protocol's JSON envelope through the C ABI (SURVEY.md 8a row H8:
process/parse-msg + keywordize-keys-1, process.clj:26-66; net/check-message, net.clj:27-37; the line
written to a node's STDIN, process.clj:162).  Golden lines: demo/go/node_test.go:51,67,102,132,178 and
doc/protocol.md:84-111."""
import ctypes as C
import json

import pytest

import oracle_lib as O

pytestmark = pytest.mark.usefixtures("engine_backend")


def sim(n=3, workload="echo", **kw):
    import maelstrom_b200 as mb
    return mb.Sim(n, workload=workload, **kw)


def send(g, line):
    return g.L.ms_send_json(g.h, line.encode())


def recv(g, ep, timeout_ns=1_000_000_000):
    buf = C.create_string_buffer(1 << 16)
    rc = g.L.ms_recv_json(g.h, ep, timeout_ns, buf, len(buf))
    assert rc >= 0, g.L.ms_last_error(g.h).decode()
    return json.loads(buf.value.decode()) if rc == 1 else None


def test_go_golden_lines_and_schema():
    g = sim()
    c0 = g.add_endpoint("c0")
    h1, h2 = g.add_endpoint("h1", 2), g.add_endpoint("h2", 2)         # host endpoints: what a --bin node would be
    # node_test.go:178  an RPC from one node to another with a key the network knows nothing about
    line = '{"src":"h1","dest":"h2","body":{"bar":"baz","msg_id":1,"type":"foo"}}'
    assert send(g, line) == 0                                          # first net id is 0 (net.clj:103)
    m = recv(g, h2)
    assert m == {"id": 0, "src": "h1", "dest": "h2", "body": {"bar": "baz", "msg_id": 1, "type": "foo"}}
    # node_test.go:51,67  error replies: code / in_reply_to / type survive, the text comes from errors.edn
    for code, golden in ((10, '{"body":{"code":10,"in_reply_to":1000,"text":"bad call","type":"error"}}'),
                         (13, '{"body":{"code":13,"in_reply_to":1000,"text":"bad call","type":"error"}}')):
        body = json.loads(golden)["body"]
        assert send(g, json.dumps({"src": "h2", "dest": "h1", "body": body})) > 0
        got = recv(g, h1)["body"]
        assert {k: got[k] for k in ("code", "in_reply_to", "type")} == {k: body[k] for k in ("code", "in_reply_to", "type")}
    # node_test.go:102,132  node output without "dest" is not a legal network message (net.clj:27-33)
    for bad in ('{"src":"n3","body":{"in_reply_to":1,"type":"init_ok"}}',
                '{"src":"n1","body":{"in_reply_to":2,"msg_id":2,"type":"echo_ok"}}',
                '{"src":"h1","dest":"h2","body":{"type":"x"},"extra":1}',           # disallowed key
                '{"src":7,"dest":"h2","body":{"type":"x"}}',                        # NodeId must be a string
                '{"src":"h1","dest":"h2","body":{"type":"x"},"id":"zero"}'):        # id must be an integer
        assert send(g, bad) == -2 and "Malformed network message" in g.L.ms_last_error(g.h).decode()
    assert send(g, "Error: oh no") == -2 and "not well-formed JSON" in g.L.ms_last_error(g.h).decode()
    assert send(g, '{"src":"h1","dest":"nobody","body":{"type":"x"}}') == -1         # net.clj:174-175
    assert "Invalid dest" in g.L.ms_last_error(g.h).decode()
    # the optional "id" is accepted and ignored: the net assigns its own (net.clj:197)
    assert send(g, '{"id":99,"src":"c0","dest":"n1","body":{"type":"echo","msg_id":1,"echo":"Please echo 35"}}') > 0
    reply = recv(g, c0)
    assert reply["src"] == "n1" and reply["dest"] == "c0" and reply["id"] != 99
    assert reply["body"] == {"type": "echo_ok", "in_reply_to": 1, "msg_id": 1, "echo": "Please echo 35"}   # doc/02-echo
    g.close()


def test_protocol_init_and_broadcast_read():
    # doc/protocol.md:84-111 init; doc/03-broadcast: broadcast / read with `messages`
    g = sim(5, workload="broadcast", topology="grid", n_values=64)
    c = g.add_endpoint("c1")
    init = {"src": "c1", "dest": "n3", "body": {"type": "init", "msg_id": 1, "node_id": "n3", "node_ids": ["n1", "n2", "n3"]}}
    assert send(g, json.dumps(init)) >= 0
    assert recv(g, c)["body"] == {"type": "init_ok", "in_reply_to": 1}
    for i, v in enumerate((7, 3, 11)):
        assert send(g, json.dumps({"src": "c1", "dest": "n%d" % i, "body": {"type": "broadcast", "message": v, "msg_id": 2 + i}})) > 0
        assert recv(g, c)["body"] == {"type": "broadcast_ok", "in_reply_to": 2 + i}
    g.run(g.now + 5_000_000)
    assert send(g, '{"src":"c1","dest":"n4","body":{"type":"read","msg_id":9}}') > 0
    assert recv(g, c)["body"] == {"type": "read_ok", "in_reply_to": 9, "messages": [3, 7, 11]}
    # a request the node has no handler for: error 10 with the registry's text (errors.edn)
    assert send(g, '{"src":"c1","dest":"n0","body":{"type":"frobnicate","msg_id":10,"x":[1,{"9":true}]}}') > 0
    assert recv(g, c)["body"] == {"type": "error", "in_reply_to": 10, "code": 10, "text": "not-supported"}
    g.close()


def test_services_and_txn_bodies():
    g = sim(2, workload="txn-list-append", max_endpoints=16)
    g.add_endpoint("lin-kv", 4)
    c = g.add_endpoint("c0")
    assert send(g, '{"src":"c0","dest":"lin-kv","body":{"type":"cas","key":3,"from":0,"to":5,"create_if_not_exists":true,"msg_id":1}}') >= 0
    assert recv(g, c)["body"] == {"type": "cas_ok", "in_reply_to": 1}
    assert send(g, '{"src":"c0","dest":"lin-kv","body":{"type":"read","key":3,"msg_id":2}}') > 0
    assert recv(g, c)["body"] == {"type": "read_ok", "in_reply_to": 2, "value": 5}
    assert send(g, '{"src":"c0","dest":"lin-kv","body":{"type":"read","key":4,"msg_id":3}}') > 0
    assert recv(g, c)["body"] == {"type": "error", "in_reply_to": 3, "code": 20, "text": "key-does-not-exist"}
    txn = {"src": "c0", "dest": "n1", "body": {"type": "txn", "msg_id": 4, "txn": [["append", 9, 1], ["r", 9, None]]}}
    assert send(g, json.dumps(txn)) > 0
    b = recv(g, c)["body"]
    assert b["type"] == "txn_ok" and b["in_reply_to"] == 4 and b["versions"][0] == 0 and b["versions"][1] >= 2
    g.close()
