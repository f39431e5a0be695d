"""Synchronous client over `net`, mirroring maelstrom.client (src/maelstrom/client.clj:41-172):
one outstanding request, stale replies discarded, 5000 ms default timeout (virtual time here),
`type: "error"` bodies raised as RPCError with the registry's definite? flag."""
from . import errors

DEFAULT_TIMEOUT_MS = 5000            # client.clj:18-20


class Timeout(Exception):            # {:type ::timeout :name :timeout :definite? false :code 0}  client.clj:96-101
    code = 0
    definite = False


class RPCError(Exception):           # {:type :rpc-error ...}  client.clj:125-138
    def __init__(self, body):
        self.code = body.get("code")
        self.name = errors.name(self.code)
        self.definite = errors.definite(self.code)
        self.body = body
        Exception.__init__(self, "%s (%s)" % (self.name, self.code))


class Client:
    def __init__(self, net):                                          # open!  client.clj:41-53
        net.next_client_id += 1
        self.net = net
        self.node_id = "c%d" % net.next_client_id
        net.add_node(self.node_id)
        self.next_msg_id = 0
        self.waiting_for = None

    def close(self):                                                  # close!  client.clj:55-59
        self.waiting_for = "closed"
        self.net.remove_node(self.node_id)

    def send(self, msg):                                              # send!  client.clj:66-79
        body = dict(msg["body"])
        if body.get("msg_id") is None:
            self.next_msg_id += 1
            body["msg_id"] = self.next_msg_id
        if self.waiting_for is not None:
            raise RuntimeError("Can't send more than one message at a time!")
        self.waiting_for = body["msg_id"]
        self.net.send({"src": self.node_id, "dest": msg["dest"], "body": body})

    def recv(self, timeout_ms=DEFAULT_TIMEOUT_MS):                    # recv!  client.clj:81-117
        target = self.waiting_for
        assert target is not None, "This client isn't waiting for any response!"
        deadline = self.net.sim.now + int(timeout_ms * 1_000_000)
        try:
            while True:
                remaining_ms = max(0, (deadline - self.net.sim.now) / 1e6)
                m = self.net.recv(self.node_id, remaining_ms)
                if m is None:
                    raise Timeout("Client read timeout")
                if m["body"].get("in_reply_to") != target:            # reply to a request we gave up on
                    continue
                return m
        finally:
            self.waiting_for = None

    def rpc(self, dest, body, timeout_ms=DEFAULT_TIMEOUT_MS):         # rpc!  client.clj:140-151
        self.send({"dest": dest, "body": body})
        m = self.recv(timeout_ms)
        if m["body"].get("type") == "error":                          # throw-errors!  client.clj:125-138
            raise RPCError(m["body"])
        return m["body"]


def with_errors(op, idempotent, fn):
    """with-errors (client.clj:153-172): run fn(); map timeouts and RPC errors to :fail / :info."""
    try:
        return fn()
    except Timeout:
        return dict(op, type="fail" if op.get("f") in idempotent else "info", error="net-timeout")
    except RPCError as e:
        kind = "fail" if (e.definite or op.get("f") in idempotent) else "info"
        return dict(op, type=kind, error=[e.name, e.body.get("text")])
