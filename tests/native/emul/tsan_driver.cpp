// tsan_driver.cpp (TEST INFRASTRUCTURE) -- drives every kernel family through the C ABI on the
// CPU SIMT emulator built with ThreadSanitizer (tools/emul_tsan.sh).  No oracle here: the point is
// TSan's verdict on unsynchronised accesses between threads of a CTA (see simt.cpp, MS_TSAN).
#include <stdio.h>
#include <string.h>

#include <vector>

#include "maelstrom_b200.h"

static ms_body body(uint16_t type, uint32_t msg_id, uint32_t p0 = 0, uint64_t p1 = 0, uint16_t extra_flags = 0) {
  ms_body b;
  memset(&b, 0, sizeof b);
  b.type = type; b.flags = MS_F_MSG_ID | extra_flags; b.msg_id = msg_id; b.p0 = p0; b.p1 = p1;
  return b;
}

static ms_sim* make(uint32_t n, uint32_t workload, uint32_t latency_ms) {
  ms_config c;
  memset(&c, 0, sizeof c);
  c.n_nodes = n; c.workload = workload; c.topology = MS_TOPO_GRID;
  c.latency_dist = MS_DIST_CONSTANT; c.latency_mean_ms = latency_ms;
  c.seed_lo = 0x4D41454C; c.n_values = 4096; c.gset_interval_ms = 7;
  c.max_endpoints = n + 16; c.ring_cap = 8192; c.max_window = 4096; c.journal_cap_log2 = 18; c.journal_level = 2;
  c.calendar_slots = 16; c.calendar_cap = 1 << 17;
  ms_sim* s = ms_create(&c);
  if (!s) { fprintf(stderr, "ms_create failed: %s\n", ms_last_error(nullptr)); }
  return s;
}

static int run(ms_sim* s, int64_t until_ns) {
  std::vector<ms_event> ev(1 << 16);
  for (;;) {
    const int rc = ms_run(s, until_ns);
    if (rc < 0) { fprintf(stderr, "ms_run: %s\n", ms_last_error(s)); return rc; }
    size_t n = 0;
    do { ms_journal_drain(s, ev.data(), nullptr, ev.size(), &n); } while (n == ev.size());
    if (rc == 0) return 0;
  }
}

int main() {
  int bad = 0;
  {  // broadcast: big windows (sender blocks, first-sight table, neighbor claims), then latency 1 (wheel)
    for (uint32_t lat = 0; lat < 2; lat++) {
      ms_sim* s = make(36, MS_W_BROADCAST, lat);
      const int c = ms_add_endpoint(s, "c0", MS_KIND_SIM_CLIENT);
      std::vector<ms_op> ops;
      for (uint32_t k = 0; k < 3600; k++) {                 // windows of > 512 messages at the hot nodes: classes 0-2
        ms_op op;
        memset(&op, 0, sizeof op);
        op.time_ns = (int64_t)(k / 1200) * 1000000; op.src = (uint32_t)c; op.dest = (k % 5 == 0) ? 14u : (k * 7) % 36;
        op.body = body(MS_T_BROADCAST, k + 1, k % 3300);
        ops.push_back(op);
      }
      ms_schedule_ops(s, ops.data(), ops.size());
      bad |= run(s, 30000000);
      ms_destroy(s);
    }
  }
  {  // g-set: snapshots, merges, reads in the same windows
    ms_sim* s = make(9, MS_W_GSET, 2);                       // latency 2 ms: replicate_full through the timing wheel
    const int c = ms_add_endpoint(s, "c0", MS_KIND_SIM_CLIENT);
    ms_body b;
    for (uint32_t i = 0; i < 9; i++) { b = body(MS_T_INIT, 100 + i); ms_send(s, (uint32_t)c, i, &b); }
    std::vector<ms_op> ops;
    for (uint32_t k = 0; k < 300; k++) {
      ms_op op;
      memset(&op, 0, sizeof op);
      op.time_ns = (int64_t)(k / 10) * 1000000; op.src = (uint32_t)c; op.dest = k % 9;
      op.body = (k % 4 == 3) ? body(MS_T_READ, k + 1) : body(MS_T_ADD, k + 1, k);
      ops.push_back(op);
    }
    ms_schedule_ops(s, ops.data(), ops.size());
    bad |= run(s, 60000000);
    ms_destroy(s);
  }
  {  // services beside an echo cluster
    ms_sim* s = make(2, MS_W_ECHO, 1);
    const char* names[4] = {"lin-kv", "seq-kv", "lww-kv", "lin-tso"};
    int sv[4];
    for (int k = 0; k < 4; k++) sv[k] = ms_add_endpoint(s, names[k], MS_KIND_SERVICE);
    const int c = ms_add_endpoint(s, "c0", MS_KIND_SIM_CLIENT);
    std::vector<ms_op> ops;
    for (uint32_t k = 0; k < 400; k++) {
      ms_op op;
      memset(&op, 0, sizeof op);
      op.time_ns = (int64_t)(k / 40) * 1000000; op.src = (uint32_t)c; op.dest = (uint32_t)sv[k % 4];
      const uint16_t t = (k % 4 == 3) ? MS_T_TS : (uint16_t)((k / 4) % 3 == 0 ? MS_T_READ : ((k / 4) % 3 == 1 ? MS_T_WRITE : MS_T_CAS));
      op.body = body(t, k + 1, k % 5, (uint64_t)(k % 3) | ((uint64_t)(k % 4) << 32), (k % 8 == 6) ? MS_F_CREATE : 0);
      ops.push_back(op);
    }
    ms_schedule_ops(s, ops.data(), ops.size());
    bad |= run(s, 20000000);
    ms_destroy(s);
  }
  {  // txn-list-append over lin-kv
    ms_sim* s = make(3, MS_W_TXN, 1);
    ms_add_endpoint(s, "lin-kv", MS_KIND_SERVICE);
    const int c = ms_add_endpoint(s, "c0", MS_KIND_SIM_CLIENT);
    std::vector<ms_op> ops;
    for (uint32_t k = 0; k < 120; k++) {
      ms_op op;
      memset(&op, 0, sizeof op);
      op.time_ns = (int64_t)(k / 6) * 1000000; op.src = (uint32_t)c; op.dest = k % 3;
      op.body = body(MS_T_TXN, k + 1, 0, 1000 + k, (k % 3) ? MS_F_APPENDS : 0);
      ops.push_back(op);
    }
    ms_schedule_ops(s, ops.data(), ops.size());
    bad |= run(s, 40000000);
    ms_destroy(s);
  }
  {  // Raft: election + replication (2.6 s of virtual time is enough for most seeds' first election)
    ms_sim* s = make(3, MS_W_RAFT, 0);
    const int c = ms_add_endpoint(s, "c0", MS_KIND_SIM_CLIENT);
    ms_body b;
    for (uint32_t i = 0; i < 3; i++) { b = body(MS_T_INIT, 100 + i); ms_send(s, (uint32_t)c, i, &b); }
    std::vector<ms_op> ops;
    for (uint32_t k = 0; k < 40; k++) {
      ms_op op;
      memset(&op, 0, sizeof op);
      op.time_ns = 4100000000ll + (int64_t)(k / 2) * 1000000; op.src = (uint32_t)c; op.dest = k % 3;
      op.body = body((k % 2) ? MS_T_WRITE : MS_T_READ, k + 1, k % 3, k);
      ops.push_back(op);
    }
    ms_schedule_ops(s, ops.data(), ops.size());
    bad |= run(s, 4200000000ll);
    uint64_t st[8];
    for (uint32_t i = 0; i < 3; i++) { ms_raft_state(s, i, st); printf("raft node %u state %llu term %llu log %llu\n", i, (unsigned long long)st[0], (unsigned long long)st[1], (unsigned long long)st[6]); }
    ms_destroy(s);
  }
  printf("driver done, errors=%d\n", bad);
  return bad ? 1 : 0;
}
