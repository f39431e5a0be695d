// ms_kernels.cu -- hand-written sm_100a kernels of the discrete-event engine.
//
// One *round* of the simulation is three launches:
//   k_release   (only when a latency distribution can produce latency > 0):
//               scatters the timing-wheel slot that just became due into the
//               per-endpoint inbox rings;
//   k_snapshot  head <- limit, limit <- tail: freezes the window every endpoint
//               consumes this round, so messages sent in round r are first
//               visible in round r+1;
//   k_round     ONE fused kernel replacing process.clj's stdin/stdout pumps,
//               the node program, net/send! (net.clj:189-221) and net/recv!
//               (net.clj:223-247): per endpoint CTA
//                 stage window keys -> bitonic sort by message id (shared memory)
//                 -> partition check at dequeue -> node transition
//                 -> block scan of (recv, emit, new) counts
//                 -> decoupled look-back across CTAs (dense message / event ids)
//                 -> :recv events, emissions: Philox loss/latency, :send events,
//                    warp-aggregated scatter into the destination rings.
// No outbox exists: a message goes HBM ring -> registers -> HBM ring.
//
// HBM-bound integer work; tensor cores are deliberately idle.
#include <cuda_runtime.h>
#include <stdint.h>
#include "ms_device.cuh"

namespace msd {

#define FULL 0xFFFFFFFFu

// ------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// streaming 16-byte store: journal / ring records are written once and read by
// another SM (or the host) later, so keep them out of L1.
__device__ __forceinline__ void st_v4(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__device__ __forceinline__ void latch_error(DevState* st, uint32_t code, uint32_t arg) {
  if (atomicCAS(&st->error, 0u, code) == 0u) st->error_arg = arg;
}

// A round is skipped (by every kernel of the round alike) when the simulation has
// reached its stop time, an error is latched, or the journal ring is more than
// half full and the host has not drained it yet (back-pressure).
__device__ __forceinline__ bool round_skipped(const Params& p, const DevState* st) {
  if (st->now >= st->stop_ns || st->error) return true;
  if (p.jlevel && !p.jdiscard && st->next_event - st->journal_drained > ((p.jmask + 1) >> 1)) return true;
  return false;
}

// ------------------------------------------------------------------ record packing
struct Rec {  // ms_msg in registers
  uint64_t id;
  int64_t deadline;
  uint32_t src, dest, msg_id, in_reply_to;
  uint32_t tf;  // type | flags << 16
  uint32_t p0;
  uint64_t p1;
};

__device__ __forceinline__ void rec_store(uint4* slot, const Rec& r) {
  st_v4(slot + 0, make_uint4((uint32_t)r.id, (uint32_t)(r.id >> 32), (uint32_t)r.deadline,
                             (uint32_t)((uint64_t)r.deadline >> 32)));
  st_v4(slot + 1, make_uint4(r.src, r.dest, r.msg_id, r.in_reply_to));
  st_v4(slot + 2, make_uint4(r.tf, r.p0, (uint32_t)r.p1, (uint32_t)(r.p1 >> 32)));
}
__device__ __forceinline__ Rec rec_load(const uint4* slot) {
  const uint4 a = ld_v4(slot + 0), b = ld_v4(slot + 1), c = ld_v4(slot + 2);
  Rec r;
  r.id = (uint64_t)a.x | ((uint64_t)a.y << 32);
  r.deadline = (int64_t)((uint64_t)a.z | ((uint64_t)a.w << 32));
  r.src = b.x; r.dest = b.y; r.msg_id = b.z; r.in_reply_to = b.w;
  r.tf = c.x; r.p0 = c.y;
  r.p1 = (uint64_t)c.z | ((uint64_t)c.w << 32);
  return r;
}

__device__ __forceinline__ void journal_write(const Params& p, DevState* st, uint64_t ev_pos, bool recv,
                                              int64_t now, const Rec& r) {
  if (p.jlevel == 0) return;
  if (!p.jdiscard && ev_pos - ld_relaxed_u64(&st->journal_drained) > p.jmask) {
    latch_error(st, E_JOURNAL_OVERFLOW, (uint32_t)ev_pos);
    return;
  }
  const uint64_t eid = ev_pos | (recv ? MS_EVENT_RECV : 0ull);
  uint4* e = p.jev + (ev_pos & p.jmask) * 2;
  st_v4(e + 0, make_uint4((uint32_t)eid, (uint32_t)(eid >> 32), (uint32_t)now, (uint32_t)((uint64_t)now >> 32)));
  st_v4(e + 1, make_uint4((uint32_t)r.id, (uint32_t)(r.id >> 32), r.src, r.dest));
  if (p.jlevel >= 2) {
    uint4* b = p.jbody + (ev_pos & p.jmask) * 2;
    st_v4(b + 0, make_uint4((uint32_t)r.id, (uint32_t)(r.id >> 32), r.msg_id, r.in_reply_to));
    st_v4(b + 1, make_uint4(r.tf, r.p0, (uint32_t)r.p1, (uint32_t)(r.p1 >> 32)));
  }
}

// ------------------------------------------------------------------ block primitives
// Exclusive scan of a[0..n) in shared memory, total written to a[n] and returned.
__device__ uint64_t block_excl_scan(uint64_t* a, int n, uint64_t* wtmp /* >= 33 */) {
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (n + nt - 1) / nt;
  const int lo = min(tid * c, n), hi = min(lo + c, n);
  uint64_t sum = 0;
  for (int i = lo; i < hi; i++) sum += a[i];
  uint64_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t y = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) wtmp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int nw = nt >> 5;
    const uint64_t w = lane < nw ? wtmp[lane] : 0;
    uint64_t wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t y = __shfl_up_sync(FULL, wi, d);
      if (lane >= d) wi += y;
    }
    wtmp[lane] = wi - w;
    if (lane == 31) wtmp[32] = wi;
  }
  __syncthreads();
  uint64_t run = wtmp[warp] + incl - sum;
  for (int i = lo; i < hi; i++) {
    const uint64_t v = a[i];
    a[i] = run;
    run += v;
  }
  const uint64_t total = wtmp[32];
  if (tid == 0) a[n] = total;
  __syncthreads();
  return total;
}

// In-place ascending bitonic sort of keys[0..np) (np a power of two).
__device__ void block_bitonic_sort(uint64_t* keys, int np) {
  const int nt = blockDim.x, tid = threadIdx.x;
  for (int k = 2; k <= np; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (np >> 1); t += nt) {
        // t-th compare-exchange pair of this stage
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool asc = (i & k) == 0;
        const uint64_t a = keys[i], b = keys[l];
        if ((a > b) == asc) { keys[i] = b; keys[l] = a; }
      }
      __syncthreads();
    }
  }
}

// Decoupled look-back (Merrill & Garland single-pass scan): exclusive prefix of
// (events, emissions) over all lower tickets.  Executed by warp 0.
__device__ void lookback(const Params& p, DevState* st, uint32_t ticket, uint32_t tag, uint64_t agg_ev,
                         uint64_t agg_em, uint64_t* out /* smem [2] */) {
  const int lane = threadIdx.x & 31;
  Status* mine = p.status + ticket;
  if (ticket == 0) {
    if (lane == 0) {
      mine->inc_ev = agg_ev; mine->inc_em = agg_em;
      mine->agg_ev = agg_ev; mine->agg_em = agg_em;
      __threadfence();
      st_release_u32(&mine->flag_agg, tag);
      st_release_u32(&mine->flag_inc, tag);
      out[0] = 0; out[1] = 0;
    }
    return;
  }
  if (lane == 0) {
    mine->agg_ev = agg_ev; mine->agg_em = agg_em;
    __threadfence();
    st_release_u32(&mine->flag_agg, tag);
  }
  uint64_t ev = 0, em = 0;
  int64_t pos = (int64_t)ticket - 1;   // nearest predecessor inspected by lane 0
  bool failed = false;
  for (;;) {
    const int64_t idx = pos - lane;
    bool has_inc = false, has_agg = false;
    if (idx < 0) {
      has_inc = true;                 // virtual ticket -1: inclusive prefix 0
    } else {
      const Status* s = p.status + idx;
      uint32_t spins = 0;
      for (;;) {
        if (ld_acquire_u32(&s->flag_inc) == tag) { has_inc = true; break; }
        if (ld_acquire_u32(&s->flag_agg) == tag) break;
        if (++spins > (1u << 22)) { failed = true; break; }
        __nanosleep(64);
      }
    }
    if (__any_sync(FULL, failed)) { failed = true; break; }
    const uint32_t inc_mask = __ballot_sync(FULL, has_inc);
    const int first = inc_mask ? (__ffs(inc_mask) - 1) : 32;
    uint64_t cev = 0, cem = 0;
    if (lane < first) {               // predecessors closer than the first inclusive one
      const Status* s = p.status + idx;
      cev = ld_relaxed_u64(&s->agg_ev); cem = ld_relaxed_u64(&s->agg_em);
    } else if (lane == first && idx >= 0) {
      const Status* s = p.status + idx;
      cev = ld_relaxed_u64(&s->inc_ev); cem = ld_relaxed_u64(&s->inc_em);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      cev += __shfl_xor_sync(FULL, cev, d);
      cem += __shfl_xor_sync(FULL, cem, d);
    }
    ev += cev; em += cem;
    if (inc_mask) break;
    pos -= 32;
  }
  if (lane == 0) {
    if (failed) latch_error(st, E_LOOKBACK_TIMEOUT, ticket);
    mine->inc_ev = ev + agg_ev; mine->inc_em = em + agg_em;
    __threadfence();
    st_release_u32(&mine->flag_inc, tag);
    out[0] = ev; out[1] = em;
  }
}

// ------------------------------------------------------------------ emission (net/send!, net.clj:189-221)
struct EmitCtx {
  int64_t now;
  uint64_t id0;      // id of this CTA's emission 0
  uint64_t ev0;      // journal position of this CTA's emission 0
  uint32_t round_lo, round_hi;
  uint32_t emitter;  // Philox stream: endpoint index or kInjector
  uint32_t idx_bias; // added to local_idx for the Philox counter (injector slices)
  // per-thread counters, reduced at the end of the CTA
  uint32_t c_send_cl, c_send_sv, c_lost, c_zero;
};

// Must be called convergently by all 32 lanes of a warp.
__device__ __forceinline__ void emit_one(const Params& p, DevState* st, const NetParams& np, EmitCtx& cx,
                                         bool valid, Rec& r, uint32_t local_idx) {
  const int lane = threadIdx.x & 31;
  bool push = false;
  if (valid && (r.dest >= p.n_ep || p.kind[r.dest] == kRemoved)) {   // net.clj:172-175
    latch_error(st, E_INVALID_DEST, r.dest);
    valid = false;
  }
  if (valid) {
    r.id = cx.id0 + local_idx;                                       // net.clj:197
    uint32_t x[4];
    philox4x32_10(local_idx + cx.idx_bias, cx.emitter, cx.round_lo, cx.round_hi, p.seed_lo, p.seed_hi, x);
    const bool cl = kind_is_client(p.kind[r.src]) || kind_is_client(p.kind[r.dest]);   // util.clj:12-16
    const uint64_t lat = cl ? 0ull : latency_ms(np, x);              // net.clj:185-187
    r.deadline = cx.now + (int64_t)lat * kTickNs;                    // net.clj:202-205
    journal_write(p, st, cx.ev0 + local_idx, false, cx.now, r);      // net.clj:208 (before the loss roll)
    if (cl) cx.c_send_cl++; else cx.c_send_sv++;
    if ((uint64_t)x[0] < np.loss_thresh) {                           // net.clj:214-215
      cx.c_lost++;
    } else if (r.deadline <= cx.now) {
      push = true;
      cx.c_zero++;
    } else {
      // timing wheel: slot of the deadline tick
      const uint64_t tick = (uint64_t)(r.deadline / kTickNs);
      if (p.cal == nullptr || lat >= p.cal_slots) {
        latch_error(st, E_CALENDAR_OVERFLOW, (uint32_t)lat);
      } else {
        const uint32_t slot = (uint32_t)tick & (p.cal_slots - 1);
        const uint32_t k = atomicAdd(&p.cal_count[slot], 1u);
        if (k >= p.cal_cap) latch_error(st, E_CALENDAR_OVERFLOW, slot);
        else rec_store(p.cal + ((size_t)slot * p.cal_cap + k) * 3, r);
      }
    }
  }
  // warp-aggregated claim of ring slots: one atomic per distinct destination
  const uint32_t key = push ? r.dest : (0x80000000u | (uint32_t)lane);
  const uint32_t mask = __match_any_sync(FULL, key);
  const int leader = __ffs(mask) - 1;
  const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
  uint32_t base = 0;
  if (push && lane == leader) base = atomicAdd(&p.tail[r.dest], (uint32_t)__popc(mask));
  base = __shfl_sync(FULL, base, leader);
  if (push) {
    const uint32_t pos = base + rank;
    if ((uint32_t)(pos - p.head[r.dest]) >= p.ring_cap) {
      latch_error(st, E_RING_OVERFLOW, r.dest);
    } else {
      rec_store(p.ring + ((size_t)r.dest * p.ring_cap + (pos & p.ring_mask)) * 3, r);
    }
  }
}

__global__ void k_set_bit(uint32_t* words, size_t word, uint32_t bit) { atomicOr(words + word, 1u << bit); }

// ------------------------------------------------------------------ k_snapshot
// head <- limit, limit <- tail; also finds the largest window of the round so
// that exactly one size class of k_round runs it.
__global__ void k_snapshot(Params p) {
  DevState* st = p.st;
  if (round_skipped(p, st)) return;
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t wmax = 0;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n_ep; e += stride) {
    const uint32_t h = p.limit[e], l = p.tail[e];
    p.head[e] = h;
    p.limit[e] = l;
    if (p.kind[e] != kRemoved) wmax = max(wmax, l - h);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wmax = max(wmax, __shfl_xor_sync(FULL, wmax, d));
  if ((threadIdx.x & 31) == 0 && wmax) atomicMax(&st->round_max_window, wmax);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (st->cal_release) {
      p.cal_count[st->cal_release - 1] = 0;
      st->cal_release = 0;
    }
    st->slot_open = 1;
  }
}

// ------------------------------------------------------------------ k_release (timing wheel -> rings)
__global__ void k_release(Params p) {
  DevState* st = p.st;
  if (round_skipped(p, st) || st->cal_release == 0) return;
  const uint32_t slot = st->cal_release - 1;
  const uint32_t n = min(p.cal_count[slot], p.cal_cap);
  const uint32_t stride = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const uint32_t n_round = (n + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool valid = i < n;
    Rec r;
    r.dest = 0;
    if (valid) r = rec_load(p.cal + ((size_t)slot * p.cal_cap + i) * 3);
    const uint32_t key = valid ? r.dest : (0x80000000u | (uint32_t)lane);
    const uint32_t mask = __match_any_sync(FULL, key);
    const int leader = __ffs(mask) - 1;
    const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(&p.tail[r.dest], (uint32_t)__popc(mask));
    base = __shfl_sync(FULL, base, leader);
    if (valid) {
      const uint32_t pos = base + rank;
      // the previous window [head, limit) is fully consumed; only limit matters here
      if ((uint32_t)(pos - p.limit[r.dest]) >= p.ring_cap) latch_error(st, E_RING_OVERFLOW, r.dest);
      else rec_store(p.ring + ((size_t)r.dest * p.ring_cap + (pos & p.ring_mask)) * 3, r);
    }
  }
}

// ------------------------------------------------------------------ node programs
// Fields of a delivered message the node programs look at (2nd and 3rd vector
// of the 48-B record); re-read from the ring (L1 hit) in every phase.
struct MsgView {
  uint32_t src, msg_id, p0;
  uint32_t tf;   // type | flags << 16
};

__device__ __forceinline__ MsgView view_load(const uint4* rec) {
  const uint4 b = rec[1], c = rec[2];
  MsgView v;
  v.src = b.x; v.msg_id = b.z; v.tf = c.x; v.p0 = c.y;
  return v;
}

// vals[] bits
constexpr uint32_t V_FRESH = 1u << 31;  // broadcast value unseen before this round
constexpr uint32_t V_RECV = 1u << 30;   // passed the partition check
constexpr uint32_t V_MASK = (1u << 30) - 1u;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint32_t nbr_count(const Params& p, uint32_t e) {
  if (p.topology == MS_TOPO_TOTAL) return p.n_servers - 1;
  return p.nbr_off[e + 1] - p.nbr_off[e];
}
__device__ __forceinline__ uint32_t nbr_at(const Params& p, uint32_t e, uint32_t j) {
  if (p.topology == MS_TOPO_TOTAL) return j < e ? j : j + 1;   // broadcast.clj:82-89
  return p.nbr[p.nbr_off[e] + j];
}
// position of `src` in e's neighbor list, or 0xFFFFFFFF
__device__ __forceinline__ uint32_t nbr_pos(const Params& p, uint32_t e, uint32_t src) {
  if (p.topology == MS_TOPO_TOTAL) {
    if (src >= p.n_servers || src == e) return 0xFFFFFFFFu;
    return src < e ? src : src - 1;
  }
  const uint32_t lo = p.nbr_off[e], hi = p.nbr_off[e + 1];
  for (uint32_t j = lo; j < hi; j++) if (p.nbr[j] == src) return j - lo;
  return 0xFFFFFFFFu;
}

// number of emissions of one delivered message (count phase)
__device__ __forceinline__ uint32_t node_emit_count(const Params& p, uint32_t e, const MsgView& w, bool is_new) {
  const uint32_t type = w.tf & 0xFFFFu;
  const bool has_id = (w.tf >> 16) & MS_F_MSG_ID;
  const bool is_reply = (w.tf >> 16) & MS_F_REPLY;
  if (p.workload == MS_W_ECHO) {                       // demo/ruby/echo.rb:28-39
    return (type == MS_T_INIT || type == MS_T_ECHO) ? 1u : 0u;
  }
  // broadcast node (doc/03-broadcast/01-broadcast.md:527-544, 02-performance.md:61-67)
  if (is_reply) return 0;                              // node.rb:159-164
  switch (type) {
    case MS_T_INIT: case MS_T_TOPOLOGY: case MS_T_READ: return 1;
    case MS_T_BROADCAST: {
      uint32_t n = has_id ? 1u : 0u;
      if (is_new) {
        n += nbr_count(p, e);
        if (nbr_pos(p, e, w.src) != 0xFFFFFFFFu) n -= 1;   // skip whoever sent it to us
      }
      return n;
    }
    default: return has_id ? 1u : 0u;                  // error 10 not-supported (errors.edn)
  }
}

// k-th emission of a delivered message (emit phase)
__device__ __forceinline__ void node_emit(const Params& p, uint32_t e, const MsgView& w, uint32_t k,
                                          uint32_t nemit, uint32_t emit_idx, uint32_t msg_id_base,
                                          uint32_t set_before, uint32_t new_before, uint64_t p1, Rec& r) {
  const uint32_t type = w.tf & 0xFFFFu;
  r.src = e; r.dest = w.src; r.msg_id = 0; r.in_reply_to = w.msg_id;
  r.p0 = 0; r.p1 = 0;
  uint32_t otype = MS_T_ERROR, oflags = MS_F_REPLY;
  if (p.workload == MS_W_ECHO) {
    otype = (type == MS_T_INIT) ? MS_T_INIT_OK : MS_T_ECHO_OK;
    oflags |= MS_F_MSG_ID;
    r.msg_id = msg_id_base + 1 + emit_idx;             // echo.rb:12-13
    if (type == MS_T_ECHO) { r.p0 = w.p0; r.p1 = p1; }
  } else {
    switch (type) {
      case MS_T_INIT: otype = MS_T_INIT_OK; break;
      case MS_T_TOPOLOGY: otype = MS_T_TOPOLOGY_OK; break;
      case MS_T_READ: otype = MS_T_READ_OK; r.p0 = set_before + new_before; break;
      case MS_T_BROADCAST: {
        const bool has_id = (w.tf >> 16) & MS_F_MSG_ID;
        if (has_id && k == nemit - 1) { otype = MS_T_BROADCAST_OK; break; }
        // gossip to the k-th neighbor other than the sender, in topology order
        const uint32_t ps = nbr_pos(p, e, w.src);
        const uint32_t j = (ps != 0xFFFFFFFFu && k >= ps) ? k + 1 : k;
        r.dest = nbr_at(p, e, j);
        otype = MS_T_BROADCAST; oflags = 0; r.in_reply_to = 0; r.p0 = w.p0;
        break;
      }
      default: otype = MS_T_ERROR; r.p0 = 10; break;
    }
  }
  r.tf = otype | (oflags << 16);
}

// ------------------------------------------------------------------ k_round
// One CTA per ticket.  Dynamic shared memory, `cap` = window capacity of this
// size class:  keys u64[cap] | vals u32[cap] | aux u64[cap+1]   (20 B / message)
// The same kernel is launched once per size class every round; only the class
// whose (cap_lo, cap] interval contains the round's largest window executes.
__global__ void __launch_bounds__(512, 2) k_round(Params p, uint32_t cap_lo, uint32_t cap, uint32_t last_class) {
  DevState* st = p.st;
  if (round_skipped(p, st) || !st->slot_open) return;
  {
    const uint32_t rm = st->round_max_window;
    if (rm <= cap_lo && cap_lo != 0) return;
    if (rm > cap && !last_class) return;
  }

  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* vals = reinterpret_cast<uint32_t*>(keys + cap);
  uint64_t* aux = reinterpret_cast<uint64_t*>(vals + cap);     // cap+1 entries
  uint32_t* tab = reinterpret_cast<uint32_t*>(aux);            // dedupe table, 2*npad entries

  __shared__ uint32_t s_ticket;
  __shared__ uint64_t s_pref[2];
  __shared__ uint64_t s_wtmp[34];
  __shared__ NetParams s_np;
  __shared__ uint32_t s_mail_base;

  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  if (tid == 0) {
    s_ticket = atomicAdd(&st->ticket, 1u);
    s_np = *p.np;
  }
  __syncthreads();
  const uint32_t ticket = s_ticket;
  const NetParams np = s_np;
  const int64_t now = st->now;
  const uint64_t round = st->round;
  const uint32_t tag = (uint32_t)round + 1u;
  const uint32_t T = p.n_inj_tickets + p.n_ep;

  EmitCtx cx;
  cx.now = now;
  cx.round_lo = (uint32_t)round; cx.round_hi = (uint32_t)(round >> 32);
  cx.idx_bias = 0;
  cx.c_send_cl = cx.c_send_sv = cx.c_lost = cx.c_zero = 0;
  uint32_t c_recv_cl = 0, c_recv_sv = 0, c_part = 0, c_replies = 0;
  uint64_t n_ev_local = 0, n_em_local = 0;

  if (ticket < p.n_inj_tickets) {
    // ---------------------------------------------------------- injector slice
    // host sends staged by ms_send (call order), then scheduled ops whose time
    // has come (schedule order): DESIGN.md 2.3 step 1.
    const uint32_t n_host = st->inj_count;
    const uint64_t tick = (uint64_t)(now / kTickNs);
    const uint32_t cur = st->sched_cursor;
    uint32_t hi = (tick + 1 < p.n_tick_off) ? p.tick_off[tick + 1] : p.n_sched;
    if (hi < cur) hi = cur;
    const uint32_t K = n_host + (hi - cur);
    const uint32_t chunk = (K + p.n_inj_tickets - 1) / p.n_inj_tickets;
    const uint32_t lo = min(ticket * chunk, K), hi_s = min(lo + chunk, K);
    const uint32_t n_local = hi_s - lo;
    n_ev_local = n_local; n_em_local = n_local;
    if (tid < 32) lookback(p, st, ticket, tag, n_ev_local, n_em_local, s_pref);
    __syncthreads();
    cx.id0 = st->next_id + s_pref[1];
    cx.ev0 = st->next_event + s_pref[0];
    cx.emitter = kInjector;
    cx.idx_bias = lo;
    for (uint32_t base = 0; base < n_local; base += nt) {
      const uint32_t j = base + tid;
      const bool valid = j < n_local;
      Rec r;
      r.dest = 0; r.src = 0;
      if (valid) {
        const uint32_t g = lo + j;
        if (g < n_host) {
          const ms_msg m = p.inj_buf[g];
          r.src = m.src; r.dest = m.dest; r.msg_id = m.msg_id; r.in_reply_to = m.in_reply_to;
          r.tf = (uint32_t)m.type | ((uint32_t)m.flags << 16); r.p0 = m.p0; r.p1 = m.p1;
        } else {
          const ms_op op = p.sched[cur + (g - n_host)];
          r.src = op.src; r.dest = op.dest; r.msg_id = op.body.msg_id; r.in_reply_to = op.body.in_reply_to;
          r.tf = (uint32_t)op.body.type | ((uint32_t)op.body.flags << 16); r.p0 = op.body.p0; r.p1 = op.body.p1;
        }
      }
      emit_one(p, st, np, cx, valid, r, j);
    }
  } else {
    // ---------------------------------------------------------- endpoint CTA
    const uint32_t e = ticket - p.n_inj_tickets;
    const uint8_t kind = p.kind[e];
    const uint32_t head = p.head[e];
    uint32_t n = (kind == kRemoved) ? 0u : (p.limit[e] - head);
    if (n > cap || n > p.max_window) {
      if (tid == 0) latch_error(st, E_WINDOW_OVERFLOW, e);
      n = 0;
    }
    const uint4* myring = p.ring + (size_t)e * p.ring_cap * 3;
    int npad = 1;
    while (npad < (int)n) npad <<= 1;

    // P1: stage (id << 16 | slot) keys; slot = offset inside the window
    for (int i = tid; i < npad; i += nt) {
      uint64_t k = ~0ull;
      if (i < (int)n) {
        const uint64_t id = *reinterpret_cast<const uint64_t*>(myring + (size_t)((head + i) & p.ring_mask) * 3);
        if (id >> 48) latch_error(st, E_ID_RANGE, e);
        k = (id << 16) | (uint64_t)i;
      }
      keys[i] = k;
    }
    __syncthreads();
    // P2: order the due set by message id (all due deadlines equal `now`; the
    // reference's PriorityBlockingQueue leaves ties unspecified, net.clj:39-40,145)
    if (n > 1) block_bitonic_sort(keys, npad);

    // P3: partition check at dequeue (net.clj:234); broadcast: first-sight dedupe table
    const bool is_server = (kind == MS_KIND_SERVER);
    const bool bcast = is_server && p.workload == MS_W_BROADCAST;
    const int tsz = 2 * npad;   // dedupe table size (u32 entries)
    if (bcast) for (int i = tid; i < tsz; i += nt) tab[i] = 0xFFFFFFFFu;
    __syncthreads();
    const uint32_t* mybits = (is_server && p.bitmap) ? p.bitmap + (size_t)e * p.bm_words : nullptr;
    for (int i = tid; i < (int)n; i += nt) {
      const uint32_t slot = (uint32_t)(keys[i] & 0xFFFFu);
      const MsgView w = view_load(myring + (size_t)((head + slot) & p.ring_mask) * 3);
      bool cut = false;
      if (np.pair_active && p.pair_bits)
        cut = (p.pair_bits[(size_t)e * p.pair_words + (w.src >> 5)] >> (w.src & 31)) & 1u;
      if (!cut && np.comp_active) {
        // bulk partition: endpoints in different components are cut; 0xFFFFFFFF = not listed (never cut)
        const uint32_t cs = p.comp[w.src], ce = p.comp[e];
        cut = cs != ce && cs != 0xFFFFFFFFu && ce != 0xFFFFFFFFu;
      }
      uint32_t val = cut ? 0u : V_RECV;
      if (bcast && !cut && (w.tf & 0xFFFFu) == MS_T_BROADCAST && !((w.tf >> 16) & MS_F_REPLY)) {
        const uint32_t v = w.p0;
        if (v >= p.n_values || v > V_MASK) {
          latch_error(st, E_VALUE_RANGE, v);
        } else {
          val |= v;
          if (!((mybits[v >> 5] >> (v & 31)) & 1u)) {
            val |= V_FRESH;
            // publish vals[i] before i becomes visible in the table
            *reinterpret_cast<volatile uint32_t*>(&vals[i]) = val;
            __threadfence_block();
            // smallest sorted index among this round's copies of v wins the slot
            uint32_t h = hash32(v) & (tsz - 1);
            for (;;) {
              uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&tab[h]);
              if (cur == 0xFFFFFFFFu) {
                const uint32_t old = atomicCAS(&tab[h], 0xFFFFFFFFu, (uint32_t)i);
                if (old == 0xFFFFFFFFu) break;
                cur = old;
              }
              if ((*reinterpret_cast<volatile uint32_t*>(&vals[cur]) & V_MASK) == v) {
                atomicMin(&tab[h], (uint32_t)i);
                break;
              }
              h = (h + 1) & (tsz - 1);
            }
          }
        }
      }
      vals[i] = val;
    }
    __syncthreads();
    // P3b: resolve first-sight winners (reads the table), keep the result in a register-free
    // way: V_FRESH stays set only for winners
    if (bcast) {
      for (int i = tid; i < (int)n; i += nt) {
        const uint32_t val = vals[i];
        if (val & V_FRESH) {
          const uint32_t v = val & V_MASK;
          uint32_t h = hash32(v) & (tsz - 1);
          uint32_t win = tab[h];
          while (win != 0xFFFFFFFFu && (vals[win] & V_MASK) != v) {   // the entry exists: probing ends on it
            h = (h + 1) & (tsz - 1);
            win = tab[h];
          }
          if (win == (uint32_t)i) {
            atomicOr(p.bitmap + (size_t)e * p.bm_words + (v >> 5), 1u << (v & 31));
          } else {
            // a lower-id copy of v is processed first this round
            // (cleared after every thread has finished reading vals[])
            keys[i] |= 0x8000ull;   // slot < 2^15: bit 15 of the key marks "not new"
          }
        }
      }
      __syncthreads();
    }
    // packed counts: emit (bits 0-31) | recv (32-47) | new (48-63)   (aux aliases the table)
    for (int i = tid; i < (int)n; i += nt) {
      const uint32_t val = vals[i];
      uint64_t c = 0;
      if (val & V_RECV) {
        c = 1ull << 32;
        const bool is_new = (val & V_FRESH) && !(keys[i] & 0x8000ull);
        if (is_server) {
          const uint32_t slot = (uint32_t)(keys[i] & 0x7FFFu);
          const MsgView w = view_load(myring + (size_t)((head + slot) & p.ring_mask) * 3);
          c |= node_emit_count(p, e, w, is_new);
        }
        if (is_new) c |= 1ull << 48;
      }
      aux[i] = c;
    }
    __syncthreads();
    // P4
    const uint64_t tot = block_excl_scan(aux, (int)n, s_wtmp);
    const uint32_t n_emit = (uint32_t)tot;
    const uint32_t n_recv = (uint32_t)(tot >> 32) & 0xFFFFu;
    const uint32_t n_new = (uint32_t)(tot >> 48);
    n_ev_local = (uint64_t)n_recv + n_emit; n_em_local = n_emit;
    // P5
    if (tid < 32) lookback(p, st, ticket, tag, n_ev_local, n_em_local, s_pref);
    const bool mailed = (kind == MS_KIND_CLIENT || kind == MS_KIND_HOST);
    if (tid == 0 && mailed && n_recv) s_mail_base = atomicAdd(&st->mail_count, n_recv);
    __syncthreads();
    const uint64_t ev_base = st->next_event + s_pref[0];
    cx.id0 = st->next_id + s_pref[1];
    cx.ev0 = ev_base + n_recv;
    cx.emitter = e;
    const bool cl_ep = kind_is_client(kind);
    const uint32_t msg_id_base = (is_server && p.next_msg_id) ? p.next_msg_id[e] : 0;
    const uint32_t set_before = (is_server && p.set_count) ? p.set_count[e] : 0;

    // P6 + P7: :recv events (net.clj:244) and the emissions of each message
    for (uint32_t base = 0; base < n; base += nt) {
      const uint32_t i = base + tid;
      const bool live = i < n;
      uint32_t my_emit = 0, e_idx0 = 0, new_before = 0;
      MsgView w;
      w.src = 0; w.msg_id = 0; w.p0 = 0; w.tf = 0;
      uint64_t p1 = 0;
      if (live) {
        const uint64_t a0 = aux[i], a1 = aux[i + 1];
        e_idx0 = (uint32_t)a0;
        my_emit = (uint32_t)a1 - (uint32_t)a0;
        new_before = (uint32_t)(a0 >> 48);
        if (vals[i] & V_RECV) {
          const uint32_t k = (uint32_t)(a0 >> 32) & 0xFFFFu;
          const uint32_t slot = (uint32_t)(keys[i] & 0x7FFFu);
          const uint4* rp = myring + (size_t)((head + slot) & p.ring_mask) * 3;
          const uint4 va = rp[0], vb = rp[1], vc = rp[2];
          Rec m;
          m.id = (uint64_t)va.x | ((uint64_t)va.y << 32);
          m.deadline = (int64_t)((uint64_t)va.z | ((uint64_t)va.w << 32));
          m.src = vb.x; m.dest = vb.y; m.msg_id = vb.z; m.in_reply_to = vb.w;
          m.tf = vc.x; m.p0 = vc.y; m.p1 = (uint64_t)vc.z | ((uint64_t)vc.w << 32);
          w.src = m.src; w.msg_id = m.msg_id; w.p0 = m.p0; w.tf = m.tf;
          p1 = m.p1;
          journal_write(p, st, ev_base + k, true, now, m);
          const bool cl = cl_ep || kind_is_client(p.kind[w.src]);
          if (cl) c_recv_cl++; else c_recv_sv++;
          if (kind == MS_KIND_SIM_CLIENT && ((w.tf >> 16) & MS_F_REPLY)) c_replies++;
          if (mailed) {
            const uint32_t mpos = s_mail_base + k;
            if (mpos >= p.mail_cap) latch_error(st, E_MAIL_OVERFLOW, e);
            else rec_store(reinterpret_cast<uint4*>(p.mail) + (size_t)mpos * 3, m);
          }
        } else {
          c_part++;
        }
      }
      uint32_t maxk = my_emit;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) maxk = max(maxk, __shfl_xor_sync(FULL, maxk, d));
      for (uint32_t k = 0; k < maxk; k++) {
        const bool valid = k < my_emit;
        Rec r;
        r.dest = 0; r.src = e;
        if (valid) node_emit(p, e, w, k, my_emit, e_idx0 + k, msg_id_base, set_before, new_before, p1, r);
        emit_one(p, st, np, cx, valid, r, e_idx0 + k);
      }
    }
    if (tid == 0 && is_server) {
      if (p.workload == MS_W_ECHO && p.next_msg_id && n_emit) p.next_msg_id[e] = msg_id_base + n_emit;
      if (p.set_count && n_new) p.set_count[e] = set_before + n_new;
      if (n > st->max_window_seen) atomicMax(&st->max_window_seen, n);
    }
  }

  // ------------------------------------------------------------ CTA epilogue
  uint32_t cnt[8] = {cx.c_send_cl, cx.c_send_sv, c_recv_cl, c_recv_sv, cx.c_lost, cx.c_zero, c_part, c_replies};
#pragma unroll
  for (int q = 0; q < 8; q++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt[q] += __shfl_xor_sync(FULL, cnt[q], d);
  }
  if (lane == 0) {
    if (cnt[0]) { atomicAdd((unsigned long long*)&st->stats[0], (unsigned long long)cnt[0]);
                  atomicAdd((unsigned long long*)&st->stats[2], (unsigned long long)cnt[0]); }
    if (cnt[1]) { atomicAdd((unsigned long long*)&st->stats[0], (unsigned long long)cnt[1]);
                  atomicAdd((unsigned long long*)&st->stats[4], (unsigned long long)cnt[1]); }
    if (cnt[2]) { atomicAdd((unsigned long long*)&st->stats[1], (unsigned long long)cnt[2]);
                  atomicAdd((unsigned long long*)&st->stats[3], (unsigned long long)cnt[2]); }
    if (cnt[3]) { atomicAdd((unsigned long long*)&st->stats[1], (unsigned long long)cnt[3]);
                  atomicAdd((unsigned long long*)&st->stats[5], (unsigned long long)cnt[3]); }
    if (cnt[4]) atomicAdd((unsigned long long*)&st->lost, (unsigned long long)cnt[4]);
    if (cnt[5]) atomicAdd(&st->zero_pending, cnt[5]);
    if (cnt[6]) atomicAdd((unsigned long long*)&st->part_drops, (unsigned long long)cnt[6]);
    if (cnt[7]) atomicAdd((unsigned long long*)&st->client_replies, (unsigned long long)cnt[7]);
  }
  __syncthreads();
  if (tid == 0) {
    if (ticket == T - 1) {
      // inclusive prefix of the last ticket = totals of the round
      st->round_ev_total = s_pref[0] + n_ev_local;
      st->round_em_total = s_pref[1] + n_em_local;
    }
    __threadfence();
    const uint32_t d = atomicAdd(&st->done, 1u);
    if (d == T - 1) {
      // last CTA of the round: commit the round
      __threadfence();
      const uint64_t ev_total = *reinterpret_cast<volatile uint64_t*>(&st->round_ev_total);
      const uint64_t em_total = *reinterpret_cast<volatile uint64_t*>(&st->round_em_total);
      st->next_event += ev_total;
      st->next_id += em_total;
      if (p.jdiscard) st->journal_drained = st->next_event;
      const uint64_t tick = (uint64_t)(now / kTickNs);
      uint32_t hi = (tick + 1 < p.n_tick_off) ? p.tick_off[tick + 1] : p.n_sched;
      if (hi > st->sched_cursor) st->sched_cursor = hi;
      st->inj_count = 0;
      const uint32_t zp = *reinterpret_cast<volatile uint32_t*>(&st->zero_pending);
      if (zp == 0) {
        st->now = now + kTickNs;
        st->time_advanced = 1;
        if (p.cal) st->cal_release = (((uint32_t)(tick + 1)) & (p.cal_slots - 1)) + 1;
      } else {
        st->time_advanced = 0;
      }
      st->zero_pending = 0;
      st->round = round + 1;
      st->rounds_run += 1;
      st->ticket = 0;
      st->done = 0;
      st->round_max_window = 0;
      st->slot_open = 0;
    }
  }
}

}  // namespace msd

// ------------------------------------------------------------------ host-callable launchers
extern "C" {

cudaError_t msk_round_smem_attr(size_t bytes) {
  return cudaFuncSetAttribute(msd::k_round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

size_t msk_round_smem_bytes(uint32_t cap) { return (size_t)cap * 20 + 16; }

void msk_set_bit(uint32_t* words, size_t word, uint32_t bit, cudaStream_t s) {
  msd::k_set_bit<<<1, 1, 0, s>>>(words, word, bit);
}

// One round = [k_release] k_snapshot, then k_round once per window-size class
// (caps[] ascending, threads[] per class); exactly one class executes.
void msk_launch_round(const msd::Params* p, int n_classes, const uint32_t* caps, const int* threads,
                      int with_release, cudaStream_t s, cudaEvent_t before_round, cudaEvent_t after_round) {
  const uint32_t n_ep = p->n_ep;
  if (with_release) msd::k_release<<<296, 256, 0, s>>>(*p);
  const int sb = 256;
  int sg = (int)((n_ep + sb - 1) / sb);
  if (sg > 296) sg = 296;
  if (sg < 1) sg = 1;
  msd::k_snapshot<<<sg, sb, 0, s>>>(*p);
  if (before_round) cudaEventRecord(before_round, s);
  for (int c = 0; c < n_classes; c++) {
    const uint32_t lo = c ? caps[c - 1] : 0u;
    msd::k_round<<<p->n_inj_tickets + n_ep, threads[c], msk_round_smem_bytes(caps[c]), s>>>(
        *p, lo, caps[c], c == n_classes - 1 ? 1u : 0u);
  }
  if (after_round) cudaEventRecord(after_round, s);
}

}  // extern "C"
