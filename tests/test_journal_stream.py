"""ms_run_streamed (include/maelstrom_b200.h): the journal handed over in batches through pinned host
memory while the simulation keeps running, in the 4-, 8-, 12- and 32-byte formats, expanded on the host by
ms_journal_decode (4-byte records: by an ms_jdecoder, which follows the stream).  Whatever the batch size and the size of the device's raw ring (back-pressure),
the stream must be the journal ms_journal_drain returns, i.e. the oracle's, event for event."""
import numpy as np
import pytest

import oracle_lib as O
from scenarios import make_pair, random_broadcast_ops

pytestmark = pytest.mark.usefixtures("engine_backend")


@pytest.mark.parametrize("fmt,buf_events,jcap,latency", [
    (8, 1 << 12, 20, 0), (12, 777, 20, 0), (32, 1 << 14, 20, 2), (8, 300, 11, 0), (8, 1 << 16, 20, 3),
    (4, 1 << 12, 20, 0), (4, 300, 11, 0), (4, 1 << 16, 20, 3), (4, 901, 20, 1)])
def test_stream_equals_oracle(fmt, buf_events, jcap, latency):
    n = 25
    g, o = make_pair(n, topology="grid", n_values=4096, ring_cap=512, max_window=256, journal_cap_log2=jcap,
                     journal_level=1, latency_dist="constant", latency_mean_ms=latency)
    cs_g = [g.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(3)]
    cs_o = [o.add_endpoint("c%d" % i, O.KIND_SIM_CLIENT) for i in range(3)]
    assert cs_g == cs_o
    ops, nv = random_broadcast_ops(n, cs_g, n_ticks=30, per_tick=10, seed=31)
    g.schedule(ops)
    o.schedule(ops)
    got, seen = [], []

    def sink(info, rounds, ev):
        assert info["format"] == fmt and len(ev) == info["n_events"] and info["n_events"] <= buf_events
        assert len(rounds) == info["n_rounds"] >= 1 and int(rounds[0]["ev_base"]) <= info["first_event"]
        seen.append((info["first_event"], info["n_events"]))
        got.append(ev.copy())

    horizon = (60 + 140 * latency) * 1_000_000
    n_ev, n_bytes = g.run_streamed(horizon // 2, sink, fmt=fmt, buf_events=buf_events, decode=True)
    g.run(horizon // 2 + 3_000_000)                    # a stretch of ordinary running + draining in between
    mid, _ = g.drain(bodies=False)
    got.append(mid)
    n2, b2 = g.run_streamed(horizon, sink, fmt=fmt, buf_events=buf_events, decode=True)
    o.run(horizon)
    ev_o, _ = o.journal()
    ev_g = np.concatenate(got)
    assert n_bytes == n_ev * fmt and n_ev + n2 + len(mid) == len(ev_o) == len(ev_g) > 5000
    for f in ("event_id", "time_ns", "msg_id", "src", "dest"):
        assert np.array_equal(ev_g[f], ev_o[f]), f
    # batches are contiguous and nothing is left to drain
    for (a, na), (b, _) in zip(seen, seen[1:]):
        assert a + na <= b
    assert len(g.drain(bodies=False)[0]) == 0
    assert g.stats() == o.stats() and g.now == o.now


def test_stream_format8_reports_what_it_cannot_hold():
    import maelstrom_b200 as mb
    # endpoint indices >= 65536 do not fit MS_JFMT_8
    g = mb.Sim(3, workload="echo", max_endpoints=70000, ring_cap=4, max_window=4, journal_level=1)
    cs = [g.add_endpoint("c%d" % i) for i in range(65540)]
    g.send(cs[-1], 0, mb.body("echo", msg_id=1, p1=1))
    with pytest.raises(mb.SimError) as e:
        g.run_streamed(2_000_000, lambda *a: None, fmt=8)
    assert "wider format" in str(e.value)
    g.close()


def test_stream_format4_decoder_needs_the_history():
    # a decoder that joins the stream late cannot name the senders of what was sent before it joined
    import maelstrom_b200 as mb
    from maelstrom_b200.engine import JournalDecoder
    g = mb.Sim(9, workload="broadcast", topology="grid", n_values=64, ring_cap=64, max_window=32, journal_level=1,
               latency_dist="constant", latency_mean_ms=3)                                # messages in flight across the cut
    c = g.add_endpoint("c0", O.KIND_SIM_CLIENT)
    ops, _ = random_broadcast_ops(9, [c], n_ticks=4, per_tick=3, seed=3)
    g.schedule(ops)
    g.run_streamed(2_000_000, lambda *a: None, fmt=4, decode=True)                     # the Sim's own decoder follows
    late = JournalDecoder(log2_window=10)
    with pytest.raises(mb.SimError) as e:
        g.run_streamed(30_000_000, lambda *a: None, fmt=4, decode=True, decoder=late)
    assert "ms_jdecoder" in str(e.value)
    g.close()
