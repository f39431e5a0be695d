(ns maelstrom.net
  "Drop-in for src/maelstrom/net.clj:79-247 over the B200 engine (include/maelstrom_b200.h through
  maelstrom.b200.Native, maelstrom_b200/csrc/ms_jni.c).  Same namespace, same public functions and
  arities -- net, jepsen-net, jepsen-os, add-node!, remove-node!, send!, recv! -- so that
  maelstrom.client, maelstrom.service, maelstrom.db and maelstrom.core keep calling what they call
  today.  The simulated servers are device-resident node programs (maelstrom.db calls start-nodes!
  instead of process/start-node!); clients and JVM-side services remain host endpoints.

  NOT compiled or run in the build image (no JVM there); the same call sequence is exercised from
  Python by maelstrom_b200/net.py + tests/test_net_mirror.py."
  (:require [clojure.tools.logging :refer [info]]
            [jepsen [net :as net] [os :as os]]
            [jepsen.net.proto :as net.proto]
            [maelstrom.util :as u]
            [maelstrom.net [journal :as j] [message :as msg]]
            [slingshot.slingshot :refer [throw+]])
  (:import (maelstrom.b200 Native Native$JournalSink)
           (java.nio ByteBuffer ByteOrder)))

;; ---------------------------------------------------------------- records (include/maelstrom_b200.h)
(def workloads {:echo 0 :broadcast 1 :g-set 2 :lin-kv 3 :txn-list-append 4 :txn-list-append-tree 5})
(def topologies {:grid 0 :line 1 :total 2 :tree 3 :tree2 3 :tree3 4 :tree4 5})
(def dists {:constant 0 :uniform 1 :exponential 2})
(def kinds {:client 1 :host 2 :sim-client 3 :service 4})
(def types
  {"init" 1 "init_ok" 2 "error" 3 "echo" 10 "echo_ok" 11 "topology" 20 "topology_ok" 21
   "broadcast" 22 "broadcast_ok" 23 "read" 24 "read_ok" 25 "add" 30 "add_ok" 31
   "replicate_one" 32 "replicate_full" 33 "write" 40 "write_ok" 41 "cas" 42 "cas_ok" 43
   "ts" 44 "ts_ok" 45 "request_vote" 50 "request_vote_res" 51 "append_entries" 52
   "append_entries_res" 53 "txn" 60 "txn_ok" 61})
(def type-names (into {} (map (fn [[k v]] [v k]) types)))
(def F-MSG-ID 1) (def F-REPLY 2) (def F-CREATE 4) (def F-APPENDS 8)

(defn- ^ByteBuffer direct [n]
  (doto (ByteBuffer/allocateDirect n) (.order ByteOrder/LITTLE_ENDIAN)))

(defn- ms-config
  "An ms_config (ABI 2, 136 bytes) as a direct buffer."
  [{:keys [node-count workload topology latency seed p-loss n-values]}]
  (doto (direct 136)
    (.putInt 0  (int node-count))
    (.putInt 4  (int (workloads workload)))
    (.putInt 8  (int (topologies (or topology :grid))))
    (.putInt 12 (int (dists (:dist latency :constant))))
    (.putInt 16 (int (:mean latency 0)))
    (.putInt 20 (unchecked-int (or seed 0x4D41454C)))
    (.putInt 24 (unchecked-int (bit-shift-right (long (or seed 0)) 32)))
    (.putDouble 32 (double (or p-loss 0.0)))
    (.putInt 40 (int (or n-values 65536)))
    (.putInt 64 2)))                                  ; journal_level: events + bodies

(defn- check! [net rc]
  (when (neg? rc)
    (if (= -1 rc)
      (throw+ {:type ::node-not-found :name :node-not-found :code 1 :definite? true}   ; net.clj:159-164
              nil (Native/lastError (:h @net)))
      (throw (ex-info (Native/lastError (:h @net)) {:code rc}))))
  rc)

;; ---------------------------------------------------------------- bodies <-> (type flags msg_id in_reply_to p0 p1)
(defn- blob! [net v]
  (let [k (swap! (:next-blob @net) inc)]
    (swap! (:blobs @net) assoc k v)
    k))

(defn- encode-body
  "doc/protocol.md:36-45: reserved keys type / msg_id / in_reply_to; 12 payload bytes.  What the
  device does not interpret stays in the host-side blob table, keyed by p1."
  [net body]
  (let [t     (name (:type body))
        flags (cond-> 0
                (:msg_id body) (bit-or F-MSG-ID)
                (:in_reply_to body) (bit-or F-REPLY)
                (:create_if_not_exists body) (bit-or F-CREATE)
                (and (= t "txn") (some #(= "append" (name (first %))) (:txn body))) (bit-or F-APPENDS))
        [p0 p1] (case t
                  "broadcast" [(:message body) 0]
                  "add"       [(:element body) 0]
                  "error"     [(:code body 13) 0]
                  ("echo" "echo_ok") [0 (blob! net (:echo body))]
                  "txn"       [0 (blob! net (:txn body))]
                  "read"      [(:key body 0) 0]
                  "write"     [(:key body) (:value body)]
                  "cas"       [(:key body) (bit-or (long (:from body)) (bit-shift-left (long (:to body)) 32))]
                  (let [extra (dissoc body :type :msg_id :in_reply_to)]
                    [0 (if (seq extra) (blob! net extra) 0)]))]
    [(get types t 1000) flags (or (:msg_id body) 0) (or (:in_reply_to body) 0) (or p0 0) (or p1 0)]))

(defn- decode-message
  "A 48-byte ms_msg -> maelstrom.net.message/Message."
  [net ^ByteBuffer b]
  (let [{:keys [names blobs workload h]} @net
        id (.getLong b 0) src (.getInt b 16) dest (.getInt b 20)
        msg-id (.getInt b 24) in-reply-to (.getInt b 28)
        type (bit-and (.getShort b 32) 0xFFFF) flags (bit-and (.getShort b 34) 0xFFFF)
        p0 (bit-and (.getInt b 36) 0xFFFFFFFF) p1 (.getLong b 40)
        t  (type-names type (str "type-" type))
        body (cond-> {:type t}
               (pos? (bit-and flags F-MSG-ID)) (assoc :msg_id msg-id)
               (pos? (bit-and flags F-REPLY))  (assoc :in_reply_to in-reply-to)
               (= t "broadcast") (assoc :message p0)
               (#{"echo" "echo_ok"} t) (assoc :echo (get @blobs p1))
               (= t "error") (assoc :code p0 :text (str "error " p0))
               (= t "ts_ok") (assoc :ts p1)
               (and (= t "read_ok") (= workload :lin-kv)) (assoc :value p1)
               (and (= t "read_ok") (u/service? (names src))) (assoc :value p1))]
    (msg/message id (names src) (names dest) body)))

;; ---------------------------------------------------------------- maelstrom.net's public functions
(defn net
  "net.clj:79-103.  `opts` carries what the reference passes on the command line and the engine
  needs at construction: :node-count, :workload, :topology (core.clj:36-47,160-199)."
  ([latency log-send? log-recv?] (net latency log-send? log-recv? {}))
  ([latency log-send? log-recv? opts]
   (let [h (Native/create (ms-config (assoc opts :latency latency)))]
     (when (zero? h) (throw (ex-info (Native/lastError 0) {})))
     (let [n (:node-count opts)]
       (atom {:h h :latency latency :log-send? log-send? :log-recv? log-recv? :workload (:workload opts)
              :ids (into {} (map (fn [i] [(str "n" i) i]) (range n)))           ; core.clj:231-238
              :names (into {} (map (fn [i] [i (str "n" i)]) (range n)))
              :next-client-id (atom -1) :blobs (atom {}) :next-blob (atom 0) :journal nil})))))

(defn jepsen-net
  "net.clj:105-122"
  [net]
  (reify net.proto/Net
    (drop!  [_ test src dest] (check! net (Native/netDrop (:h @net) ((:ids @net) src) ((:ids @net) dest))))
    (heal!  [_ test] (check! net (Native/netHeal (:h @net))))
    (slow!  [_ test] (check! net (Native/netSlow (:h @net))))
    (slow!  [_ test opts] (check! net (Native/netSlow (:h @net))))
    (fast!  [_ test] (check! net (Native/netFast (:h @net))))
    (flaky! [_ test] (check! net (Native/netFlaky (:h @net))))
    (shape! [_ test nodes behavior] nil)))

(defn jepsen-os
  "net.clj:124-137: the journal's lifecycle, on the primary node only."
  [net]
  (reify os/OS
    (setup! [_ test node]
      (when (= node (first (:nodes test)))
        (swap! net assoc :journal (j/journal test))))
    (teardown! [_ test node]
      (when (= node (first (:nodes test)))
        (when-let [jr (:journal @net)] (j/close! jr))
        (Native/destroy (:h @net))))))

(defn add-node!
  "net.clj:139-146"
  [net node-id]
  (assert (string? node-id) (str "Node id " (pr-str node-id) " must be a string"))
  (when-not ((:ids @net) node-id)
    (let [kind (cond (u/client? node-id) (kinds :client)                        ; util.clj:7-10
                     (#{"lin-kv" "seq-kv" "lww-kv" "lin-tso"} node-id)
                     (if (:jvm-services? @net) (kinds :host) (kinds :service))  ; service.clj:290-296
                     :else (kinds :host))
          idx  (check! net (Native/addEndpoint (:h @net) node-id kind))]
      (swap! net #(-> % (assoc-in [:ids node-id] idx) (assoc-in [:names idx] node-id)))))
  net)

(defn remove-node!
  "net.clj:148-152"
  [net node-id]
  (when-let [idx ((:ids @net) node-id)]
    (Native/removeEndpoint (:h @net) idx)
    (swap! net #(-> % (update :ids dissoc node-id) (update :names dissoc idx))))
  net)

(defn start-nodes!
  "Replaces process/start-node! (process.clj:168-215) for every server: the built-in node program
  of the workload runs on the device.  maelstrom.db/setup! still does the init RPC (db.clj:46-69)."
  [net]
  (check! net (Native/startNodes (:h @net) (workloads (:workload @net))))
  net)

(defn stop-nodes! [net] (check! net (Native/stopNodes (:h @net))) net)   ; process.clj:217-256

(defn send!
  "net.clj:189-221"
  [net message]
  (let [{:keys [h ids log-send?]} @net
        m (msg/validate (msg/message (:src message) (:dest message) (:body message)))]
    (assert (ids (:src m))  (str "Invalid source for message " (pr-str m)))     ; net.clj:172-173
    (assert (ids (:dest m)) (str "Invalid dest for message " (pr-str m)))       ; net.clj:174-175
    (let [[type flags msg-id in-reply-to p0 p1] (encode-body net (:body m))
          id (Native/send h (ids (:src m)) (ids (:dest m)) type flags msg-id in-reply-to p0 p1)]
      (check! net (int (min id 0)))
      (when log-send? (info :send (pr-str (assoc m :id id)))))
    net))

(defn recv!
  "net.clj:223-247: the message, or nil after timeout-ms of (virtual) time."
  [net node timeout-ms]
  (let [idx (or ((:ids @net) node)
                (throw+ {:type ::node-not-found :name :node-not-found :code 1 :definite? true}
                        nil (str "No such node in network: " (pr-str node))))
        buf (direct 48)]
    (when (= 1 (check! net (Native/recv (:h @net) idx (* 1000000 (long timeout-ms)) buf)))
      (let [m (decode-message net buf)]
        (when (:log-recv? @net) (info :recv (pr-str m)))
        m))))

;; ---------------------------------------------------------------- journal hand-over (net/journal.clj:205-239)
(defn drain-journal!
  "Moves everything journaled on the device into maelstrom.net.journal's Fressian stripes as
  Event{id time type message} records, so net.checker / net.viz read the files they always read.
  (ms_journal_open writes the same Fressian records from C++ without going through the JVM.)"
  [net]
  (let [{:keys [h journal names]} @net
        cap 65536 ev (direct (* 32 cap)) bd (direct (* 32 cap))]
    (loop []
      (let [n (Native/journalDrain h ev bd cap)]
        (check! net (int (min n 0)))
        (dotimes [i n]
          (let [o (* 32 i)
                eid (.getLong ev o) recv? (neg? eid)
                t (type-names (bit-and (.getShort bd (+ o 16)) 0xFFFF) "?")
                m (msg/message (.getLong ev (+ o 16)) (names (.getInt ev (+ o 24))) (names (.getInt ev (+ o 28)))
                               {:type t})]
            (j/log-event! journal (j/->Event (bit-and eid Long/MAX_VALUE) (.getLong ev (+ o 8))
                                             (if recv? :recv :send) m))))
        (when (= n cap) (recur))))
    net))
