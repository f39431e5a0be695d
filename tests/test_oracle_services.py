"""The oracle's services (oracle/oracle.cpp Service, restating src/maelstrom/service.clj) pinned
to the reference's own unit test, test/maelstrom/service_test.clj:6-53 (seq-kv-test), restated
case by case, plus the documented behaviour of lin-kv / lin-tso / lww-kv."""
import numpy as np

import oracle_lib as O

X = 7            # the key :x of the reference test


def rnd_stream(seed):
    rng = np.random.default_rng(seed)
    while True:
        yield int(rng.integers(0, 1 << 32))


def prep(buf=16):
    # service_test.clj:8-19: (sequential buf (persistent-kv)), then client "c0" writes x = 0 .. buf/2 - 1
    kv = O.Service("seq-kv", buf)
    r = rnd_stream(1)
    for i in range(buf // 2):
        out = kv.handle(1000, O.body("write", p0=X, p1=i), next(r))
        assert out.type == O.T["write_ok"]
    return kv


def test_seq_kv_fresh_client_reads_return_old_state():
    # service_test.clj:20-28: 64 fresh clients read x; more than one distinct value comes back
    kv, r = prep(), rnd_stream(2)
    reads = {int(kv.handle(2000 + i, O.body("read", p0=X), next(r)).p1) for i in range(64)}
    assert len(reads) > 1
    assert reads <= set(range(8))


def test_seq_kv_write_then_read_is_recent():
    # service_test.clj:30-38: a client that writes something unique is moved to the newest state
    for i in range(8):
        kv, r = prep(), rnd_stream(10 + i)
        client = 3000 + i
        kv.handle(client, O.body("write", p0=99, p1=i), next(r))
        out = kv.handle(client, O.body("read", p0=X), next(r))
        assert out.type == O.T["read_ok"] and out.p1 == 16 // 2 - 1


def test_seq_kv_reading_a_ton_converges():
    # service_test.clj:40-52: one client's reads are monotone and reach the newest value within 160 tries
    for i in range(8):
        kv, r = prep(), rnd_stream(20 + i)
        seen = []
        for trial in range(160):
            v = int(kv.handle(4000, O.body("read", p0=X), next(r)).p1)
            seen.append(v)
            if v == 7:
                break
        assert seen[-1] == 7
        assert seen == sorted(seen)            # "each client always observes a monotonic sequence" (service.clj:218-221)


def test_seq_kv_buffer_keeps_the_last_n_states_only():
    kv, r = O.Service("seq-kv", 4), rnd_stream(3)
    for i in range(20):
        kv.handle(1, O.body("write", p0=X, p1=i), next(r))
    reads = {int(kv.handle(100 + i, O.body("read", p0=X), next(r)).p1) for i in range(200)}
    assert reads == {16, 17, 18, 19}


def test_lin_kv_read_write_cas_and_error_codes():
    # service.clj:31-58; codes resources/errors.edn 20, 22
    kv = O.Service("lin-kv")
    assert (kv.handle(1, O.body("read", p0=5)).type, kv.handle(1, O.body("read", p0=5)).p0) == (O.T["error"], 20)
    assert kv.handle(1, O.body("write", p0=5, p1=3)).type == O.T["write_ok"]
    out = kv.handle(2, O.body("read", p0=5))
    assert out.type == O.T["read_ok"] and out.p1 == 3
    bad = kv.handle(2, O.body("cas", p0=5, p1=4 | (9 << 32)))
    assert bad.type == O.T["error"] and bad.p0 == 22
    assert kv.handle(2, O.body("cas", p0=5, p1=3 | (9 << 32))).type == O.T["cas_ok"]
    assert kv.handle(2, O.body("read", p0=5)).p1 == 9
    missing = kv.handle(2, O.body("cas", p0=6, p1=0 | (1 << 32)))
    assert missing.type == O.T["error"] and missing.p0 == 20
    assert kv.handle(2, O.body("cas", p0=6, p1=0 | (1 << 32), create=True)).type == O.T["cas_ok"]
    assert kv.handle(2, O.body("read", p0=6)).p1 == 1
    assert kv.handle(2, O.body("ts")) is None            # no clause in `case`: the service thread logs and moves on
    reply = kv.handle(2, O.body("read", msg_id=77, p0=6))
    assert reply.in_reply_to == 77 and reply.flags == O.F_REPLY       # service.clj:255-256


def test_lin_tso_counts_from_zero():
    tso = O.Service("lin-tso")
    assert [int(tso.handle(1, O.body("ts")).p1) for _ in range(5)] == [0, 1, 2, 3, 4]   # service.clj:123-141
    assert tso.handle(1, O.body("read", p0=0)) is None


def test_lww_kv_is_two_replicas_that_never_merge():
    # service.clj:229-236: the merged replica is dropped by the second replicas' binding, so a
    # write lands on one random replica and is visible only there
    kv = O.Service("lww-kv")
    assert kv.handle(1, O.body("write", p0=1, p1=5), rnd=0).type == O.T["write_ok"]          # replica 0
    assert kv.handle(1, O.body("read", p0=1), rnd=0).p1 == 5
    miss = kv.handle(1, O.body("read", p0=1), rnd=0x80000000)                                 # replica 1
    assert miss.type == O.T["error"] and miss.p0 == 20
    # LWWKV's cas has no create_if_not_exists branch (service.clj:80-95)
    out = kv.handle(1, O.body("cas", p0=2, p1=0 | (1 << 32), create=True), rnd=0)
    assert out.type == O.T["error"] and out.p0 == 20
