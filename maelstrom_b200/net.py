"""Host-side mirror of maelstrom.net's public functions (src/maelstrom/net.clj:79-247) over
the C ABI: the same names, argument meaning and error behaviour, with message maps
`{"src", "dest", "body"}` in and out (the JSON envelope of doc/protocol.md:36-45).

Works over anything with the engine.Sim method set (the CUDA engine in production; the
tests also drive the CPU oracle through it, which is what makes them read like tests of
the reference's own net).  Node ids are strings: "n<i>" are the simulated servers,
ids starting with "c" are clients (util.clj:7-10), anything else is a host endpoint.
"""
import logging

from . import errors
from .engine import F_MSG_ID, F_REPLY, KIND_CLIENT, KIND_HOST, KIND_SERVICE, SERVICES, TYPES, TYPE_NAMES


log = logging.getLogger("maelstrom.net")


class NodeNotFound(Exception):
    """{:type ::node-not-found :name :node-not-found :code 1 :definite? true} (net.clj:159-164)"""
    code = 1
    definite = True


class Net:
    def __init__(self, sim, body_factory, latency=None, log_send=False, log_recv=False):
        # (net latency log-send? log-recv?)  net.clj:79-103; the latency map is part of the sim's config
        self.sim = sim
        self._body = body_factory
        self.latency = latency
        self.log_send, self.log_recv = log_send, log_recv
        self.ids = {"n%d" % i: i for i in range(sim.n_nodes)}          # core.clj:231-238
        self.names = {i: n for n, i in self.ids.items()}
        self.next_client_id = -1                                      # net.clj:102
        self._blobs = {}                                              # payloads the device does not interpret
        self._next_blob = 1
        self._types = dict(TYPES)
        self._type_names = dict(TYPE_NAMES)
        # txn-list-append: database values travel as version ids (include/maelstrom_b200.h); the
        # contents are replayed here with apply-txn (demo/clojure/single_key_txn.clj:115-127)
        self._versions = {0: {}, 1: {}}
        self._txn_of = {}                                             # (client, msg_id) -> micro-ops

    # ---------------------------------------------------------------- endpoints
    def add_node(self, node_id):                                      # add-node!  net.clj:139-146
        assert isinstance(node_id, str), "Node id %r must be a string" % (node_id,)
        if node_id not in self.ids:
            kind = KIND_CLIENT if node_id.startswith("c") else KIND_HOST     # util.clj:7-10
            idx = self.sim.add_endpoint(node_id, kind)
            self.ids[node_id] = idx
            self.names[idx] = node_id
        return self

    def start_services(self, names=SERVICES):
        """service/start-services! with service/default-services (service.clj:265-296): the four
        Maelstrom-provided services join the network as device-resident endpoints."""
        for name in names:
            if name not in self.ids:
                idx = self.sim.add_endpoint(name, KIND_SERVICE)
                self.ids[name] = idx
                self.names[idx] = name
        return self

    def remove_node(self, node_id):                                   # remove-node!  net.clj:148-152
        idx = self.ids.pop(node_id, None)
        if idx is not None:
            self.sim.remove_endpoint(idx)
            self.names.pop(idx, None)
        return self

    # ---------------------------------------------------------------- bodies
    def _type_code(self, t):
        if t not in self._types:                                      # types the device has no handler for
            code = 1000 + len(self._types)
            self._types[t] = code
            self._type_names[code] = t
        return self._types[t]

    def _encode(self, body):
        t = body["type"]
        p0, p1 = 0, 0
        if t == "broadcast":
            p0 = int(body["message"])
        elif t in ("echo", "echo_ok"):
            p1 = self._next_blob
            self._next_blob += 1
            self._blobs[p1] = body.get("echo")
        elif t == "error":
            p0 = int(body.get("code", 13))
        elif t == "add":
            p0 = int(body["element"])
        create = False
        if t == "txn":                                                # workload/txn_list_append.clj
            ops = [list(op) for op in body["txn"]]
            handle = self._next_blob
            self._next_blob += 1
            self._blobs[handle] = ops
            if getattr(self.sim, "workload", 1) == 5:
                # hash-tree node (datomic_list_append.rb): it looks at the keys, so they travel in the payload --
                # up to four micro-ops of 16 bits: valid << 15 | append << 14 | key (include/maelstrom_b200.h)
                if len(ops) > 4 or any(not (0 <= int(op[1]) < 16384) for op in ops):
                    raise ValueError("txn-list-append-tree: at most 4 micro-ops on integer keys below 16384")
                packed = 0
                for i, op in enumerate(ops):
                    packed |= (0x8000 | (0x4000 if op[0] == "append" else 0) | int(op[1])) << (16 * i)
                return self._body(self._type_code(t), msg_id=body.get("msg_id"), p0=handle, p1=packed)
            return self._body(self._type_code(t), msg_id=body.get("msg_id"), p1=handle,
                              appends=any(op[0] == "append" for op in ops))
        if "key" in body and t in ("read", "write", "cas"):           # service requests (doc/services.md)
            p0 = int(body["key"])
            if t == "write":
                p1 = int(body["value"])
            elif t == "cas":
                p1 = int(body["from"]) | (int(body["to"]) << 32)
                create = bool(body.get("create_if_not_exists"))
            return self._body(self._type_code(t), msg_id=body.get("msg_id"), in_reply_to=body.get("in_reply_to"),
                              p0=p0, p1=p1, create=create)
        extra = {k: v for k, v in body.items()
                 if k not in ("type", "msg_id", "in_reply_to", "message", "echo", "code", "text", "element")}
        if extra and p1 == 0:
            p1 = self._next_blob
            self._next_blob += 1
            self._blobs[p1] = extra
        return self._body(self._type_code(t), msg_id=body.get("msg_id"), in_reply_to=body.get("in_reply_to"),
                          p0=p0, p1=p1)

    def _decode(self, m):
        code = int(m["type"])
        t = self._type_names.get(code, "type-%d" % code)
        body = {"type": t}
        if int(m["flags"]) & F_MSG_ID:
            body["msg_id"] = int(m["msg_id"])
        if int(m["flags"]) & F_REPLY:
            body["in_reply_to"] = int(m["in_reply_to"])
        if t == "broadcast":
            body["message"] = int(m["p0"])
        elif t in ("echo", "echo_ok"):
            body["echo"] = self._blobs.get(int(m["p1"]))
        elif t == "error":
            body["code"] = int(m["p0"])
            body["text"] = errors.name(int(m["p0"]))
        elif t == "read_ok" and self.names.get(int(m["src"])) in SERVICES:
            body["value"] = int(m["p1"])                                # service.clj:38-40
        elif t == "txn_ok":
            old, new = int(m["p1"]) & 0xFFFFFFFF, int(m["p1"]) >> 32
            ops = self._txn_of.pop((int(m["dest"]), int(m["in_reply_to"])), None)
            if ops is not None and old in self._versions:
                state, done = dict(self._versions[old]), []
                for f, k, v in ops:                                   # apply-txn
                    if f == "r":
                        done.append([f, k, list(state[k]) if k in state else None])
                    else:
                        state[k] = list(state.get(k, [])) + [v]
                        done.append([f, k, v])
                self._versions[new] = state
                body["txn"] = done
            body["versions"] = [old, new]
        elif t == "ts_ok":
            body["ts"] = int(m["p1"])                                   # service.clj:127-128
        elif t == "read_ok" and getattr(self.sim, "workload", 1) == 3:
            body["value"] = int(m["p1"])                                # lin-kv served by Raft: raft.py:171
        elif t == "read_ok":
            # the device message carries the set size; the members are read back from the node.
            # broadcast replies `messages` (workload/broadcast.clj:33-35), g-set `value` (g_set.rb:14)
            src = int(m["src"])
            members = sorted(int(v) for v in self.sim.node_set(src)) if src < self.sim.n_nodes else []
            key = "value" if getattr(self.sim, "workload", 1) == 2 else "messages"
            body[key] = members[:int(m["p0"])]
        elif int(m["p1"]) in self._blobs and isinstance(self._blobs[int(m["p1"])], dict):
            body.update(self._blobs[int(m["p1"])])
        return {"id": int(m["id"]), "src": self.names.get(int(m["src"]), str(int(m["src"]))),
                "dest": self.names.get(int(m["dest"]), str(int(m["dest"]))), "body": body}

    # ---------------------------------------------------------------- data plane
    def send(self, message):                                          # send!  net.clj:189-221
        src, dest = message.get("src"), message.get("dest")
        assert src, "No source for message %r" % (message,)           # message.clj:17-25
        assert dest, "No destination for message %r" % (message,)
        assert src in self.ids, "Invalid source for message %r" % (message,)      # net.clj:172-173
        assert dest in self.ids, "Invalid dest for message %r" % (message,)       # net.clj:174-175
        if message["body"].get("type") == "txn" and "msg_id" in message["body"]:
            self._txn_of[(self.ids[src], int(message["body"]["msg_id"]))] = [list(op) for op in message["body"]["txn"]]
        rc = self.sim.send(self.ids[src], self.ids[dest], self._encode(message["body"]))
        if rc < 0:
            raise NodeNotFound("No such node in network: %r" % (dest,))
        if self.log_send:                                             # (when log-send? (info :send ...))  net.clj:211
            log.info(":send %r", dict(message, id=rc))
        return self

    def recv(self, node_id, timeout_ms):                              # recv!  net.clj:223-247
        if node_id not in self.ids:
            raise NodeNotFound("No such node in network: %r" % (node_id,))
        m = self.sim.recv(self.ids[node_id], int(timeout_ms * 1_000_000))
        if m is None:
            return None
        out = self._decode(m)
        if self.log_recv:                                             # (when log-recv? (info :recv ...))  net.clj:241
            log.info(":recv %r", out)
        return out

    # ---------------------------------------------------------------- jepsen.net.proto/Net  (net.clj:105-122)
    def drop(self, test, src, dest):
        self.sim.drop(self.ids[src], self.ids[dest])

    def heal(self, test=None):
        self.sim.heal()

    def slow(self, test=None):
        self.sim.slow()

    def fast(self, test=None):
        self.sim.fast()

    def flaky(self, test=None):
        self.sim.flaky()


def to_wire(message):
    """The JSON envelope a node process would see / print (process.clj:26-66,162):
    `{"src","dest","body"}` (+ the net id on delivery)."""
    out = {"src": message["src"], "dest": message["dest"], "body": dict(message["body"])}
    if "id" in message:
        out["id"] = message["id"]
    return out
