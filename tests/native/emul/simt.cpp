// simt.cpp (TEST INFRASTRUCTURE) -- fiber scheduler and runtime-API stubs behind
// tests/native/emul/cuda_runtime.h.  See that header for the execution model.
#include <cuda_runtime.h>
#include <stdio.h>
#include <sys/mman.h>

#include <chrono>
#include <thread>
#include <vector>

// MS_TSAN (tools/emul_tsan.sh): kernels + engine are compiled with -fsanitize=thread, this file is
// not.  Every emulated thread is a TSan fiber; the scheduler's context switches carry no
// synchronisation, __syncthreads() and warp collectives are release/acquire pairs, a kernel launch
// is ordered after the host code before it and before the host code after it, and CTAs of a launch
// are chained (their __shared__ statics are the same host memory).  What TSan then reports is two
// threads of ONE CTA touching the same shared / global location with no barrier in between.
#ifdef MS_TSAN
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
void __tsan_acquire(void* addr);
void __tsan_release(void* addr);
}
#define TSAN(x) x
#else
#define TSAN(x)
#endif

namespace simt {

thread_local Tls tls;

// ------------------------------------------------------------------ context switch (x86-64 SysV)
extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size simt_switch,.-simt_switch
)");

constexpr size_t kStack = 128 << 10;
constexpr unsigned kMaxThreads = 1024;

struct Warp {
  unsigned live = 0;       // lanes that have not exited
  unsigned count = 0;      // lanes waiting in the current collective
  unsigned gen = 0;
  int op = 0, arg0 = 0;
  bool arrived[32];
  uint64_t in[32];
  int arg[32];
  uint64_t out[2][32];
};

struct Sched {
  char* stacks = nullptr;            // kMaxThreads * kStack
  void* sp[kMaxThreads];
  bool done[kMaxThreads];
  void* main_sp = nullptr;
  unsigned nt = 0, live = 0, cur = 0;
  unsigned bar_count = 0, bar_gen = 0;
  int bar_or[2] = {0, 0};
  Warp warps[kMaxThreads / 32];
  const std::function<void()>* body = nullptr;
  unsigned char* smem = nullptr;
  uint64_t progress = 0;
  void* fib[kMaxThreads];            // MS_TSAN: TSan fiber of each emulated thread
  void* main_fib = nullptr;
  char sync_bar, sync_launch_begin, sync_launch_end, sync_cta_chain;   // MS_TSAN: addresses to release / acquire on
  char sync_warp[kMaxThreads / 32];
  unsigned bar_calls[kMaxThreads];   // MS_TSAN self-test
};
static thread_local Sched* g_s = nullptr;

static Sched* sched() {
  if (!g_s) {
    g_s = new Sched();
    g_s->stacks = (char*)mmap(nullptr, (size_t)kMaxThreads * kStack, PROT_READ | PROT_WRITE,
                              MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_s->stacks == (char*)MAP_FAILED) { fprintf(stderr, "simt: mmap failed\n"); abort(); }
    if (posix_memalign((void**)&g_s->smem, 128, 256 << 10)) abort();
  }
  return g_s;
}

[[noreturn]] static void die(const char* what) {
  fprintf(stderr, "simt emulator: %s (block %u thread %u)\n", what, tls.bid.x, tls.tid.x);
  abort();
}

static void yield_to_scheduler() {
  Sched* s = g_s;
  const unsigned me = s->cur;
  TSAN(__tsan_switch_to_fiber(s->main_fib, 1);)
  simt_switch(&s->sp[me], s->main_sp);
  tls.tid.x = me;   // resumed
}

static void release_barrier(Sched* s) {
  s->bar_count = 0;
  s->bar_or[(s->bar_gen + 1) & 1] = 0;
  s->bar_gen++;
  s->progress++;
}

static void finish_collective(Warp& w);

static void fiber_main() {
  Sched* s = g_s;
  TSAN(__tsan_acquire(&s->sync_launch_begin); __tsan_acquire(&s->sync_cta_chain);)
  (*s->body)();
  TSAN(__tsan_release(&s->sync_launch_end); __tsan_release(&s->sync_cta_chain);)
  const unsigned me = s->cur;
  s->done[me] = true;
  s->live--;
  s->progress++;
  Warp& w = s->warps[me >> 5];
  w.live--;
  // an exiting thread may complete what the others are waiting for
  if (s->bar_count && s->bar_count == s->live) release_barrier(s);
  if (w.count && w.count == w.live) finish_collective(w);
  void* dummy;
  TSAN(__tsan_switch_to_fiber(s->main_fib, 1);)
  simt_switch(&dummy, s->main_sp);
  die("resumed a finished thread");
}

static int barrier(int pred, bool plain);
void sync_threads() { (void)barrier(0, true); }
int sync_threads_or(int pred) { return barrier(pred, false); }

static int barrier(int pred, bool plain) {
  Sched* s = g_s;
#ifdef MS_TSAN
  // self-test of the race check: MS_TSAN_SKIP_BARRIER=k turns every k-th barrier of every CTA into
  // a no-op (the same one for all its threads), which TSan must then report
  static const unsigned skip = getenv("MS_TSAN_SKIP_BARRIER") ? (unsigned)atoi(getenv("MS_TSAN_SKIP_BARRIER")) : 0;
  if (skip && plain && (++s->bar_calls[s->cur] % skip) == 0) return 0;
#endif
  const unsigned gen = s->bar_gen;
  if (pred) s->bar_or[gen & 1] = 1;
  s->bar_count++;
  s->progress++;
  TSAN(__tsan_release(&s->sync_bar);)
  if (s->bar_count == s->live) release_barrier(s);
  while (s->bar_gen == gen) yield_to_scheduler();
  TSAN(__tsan_acquire(&s->sync_bar);)
  return s->bar_or[gen & 1];
}

static void finish_collective(Warp& w) {
  uint64_t* out = w.out[w.gen & 1];
  uint64_t any = 0, ballot = 0;
  for (int l = 0; l < 32; l++)
    if (w.arrived[l] && w.in[l]) { any = 1; ballot |= 1ull << l; }
  for (int l = 0; l < 32; l++) {
    if (!w.arrived[l]) continue;
    switch (w.op) {
      case OP_SHFL_UP: out[l] = (l >= w.arg0 && w.arrived[l - w.arg0]) ? w.in[l - w.arg0] : w.in[l]; break;
      case OP_SHFL_XOR: out[l] = w.arrived[l ^ w.arg0] ? w.in[l ^ w.arg0] : w.in[l]; break;
      case OP_SHFL_IDX: { const int src = w.arg[l] & 31; out[l] = w.arrived[src] ? w.in[src] : w.in[l]; break; }
      case OP_MATCH_ANY: {
        uint64_t m = 0;
        for (int q = 0; q < 32; q++) if (w.arrived[q] && w.in[q] == w.in[l]) m |= 1ull << q;
        out[l] = m;
        break;
      }
      case OP_ANY: out[l] = any; break;
      case OP_BALLOT: out[l] = ballot; break;
      default: die("unknown warp collective");
    }
  }
  for (int l = 0; l < 32; l++) w.arrived[l] = false;
  w.count = 0;
  w.op = 0;
  w.gen++;
  g_s->progress++;
}

uint64_t warp_op(int op, unsigned mask, uint64_t v, int arg) {
  Sched* s = g_s;
  const unsigned me = s->cur, lane = me & 31;
  Warp& w = s->warps[me >> 5];
  if (!((mask >> lane) & 1u)) die("warp collective called by a lane outside its mask");
  if (w.count == 0) {
    w.op = op;
    w.arg0 = arg;
  } else if (w.op != op || (op != OP_SHFL_IDX && w.arg0 != arg)) {
    die("divergent warp collective: lanes of one warp reached different __*_sync calls");
  }
  const unsigned gen = w.gen;
  w.arrived[lane] = true;
  w.in[lane] = v;
  w.arg[lane] = arg;
  w.count++;
  s->progress++;
  TSAN(__tsan_release(&s->sync_warp[me >> 5]);)
  if (w.count == w.live) finish_collective(w);
  while (w.gen == gen) yield_to_scheduler();
  TSAN(__tsan_acquire(&s->sync_warp[me >> 5]);)
  return w.out[gen & 1][lane];
}

void relax() {
  // a spin-wait (k_barrier polling a peer's flag): let the other host threads (shards) run AND
  // the other fibers of this CTA, which on a GPU would be running concurrently
  std::this_thread::yield();
  if (g_s && g_s->body) {
    g_s->progress++;
    yield_to_scheduler();
  }
}

unsigned char* dyn_smem() { return sched()->smem; }

static void run_cta(Sched* s, unsigned nt) {
  s->nt = nt;
  s->live = nt;
  s->bar_count = 0;
  s->bar_or[0] = s->bar_or[1] = 0;
  const unsigned nw = (nt + 31) / 32;
  for (unsigned w = 0; w < nw; w++) {
    Warp& W = s->warps[w];
    W.live = (w + 1) * 32 <= nt ? 32 : nt - w * 32;
    W.count = 0;
    W.op = 0;
    for (int l = 0; l < 32; l++) W.arrived[l] = false;
  }
  for (unsigned t = 0; t < nt; t++) {
    s->done[t] = false;
    s->bar_calls[t] = 0;
    uintptr_t top = (uintptr_t)(s->stacks + (size_t)(t + 1) * kStack) & ~(uintptr_t)15;
    void** p = (void**)top;
    p[-1] = nullptr;                  // return address slot of fiber_main (never used)
    p[-2] = (void*)&fiber_main;       // popped by simt_switch's ret
    for (int r = 3; r <= 8; r++) p[-r] = nullptr;   // rbp rbx r12-r15
    s->sp[t] = (void*)(p - 8);
    TSAN(s->fib[t] = __tsan_create_fiber(0);)
  }
  TSAN(s->main_fib = __tsan_get_current_fiber();)
  // MS_EMUL_ORDER: 0 = threads resume in index order (default), 1 = reverse, 2 = a different
  // pseudo-random order every sweep.  Results must not depend on it: a missing barrier usually does.
  static const int order_mode = getenv("MS_EMUL_ORDER") ? atoi(getenv("MS_EMUL_ORDER")) : 0;
  static thread_local uint32_t lcg = 12345u;
  while (s->live) {
    const uint64_t before = s->progress;
    uint32_t mul = 1, add = 0;
    if (order_mode == 2) {            // t -> (t * odd + add) mod 2^k is a permutation of [0, 2^k)
      lcg = lcg * 1664525u + 1013904223u;
      mul = (lcg >> 8) | 1u;
      add = lcg >> 20;
    }
    unsigned span = 1;
    while (span < nt) span <<= 1;
    for (unsigned u = 0; u < span; u++) {
      unsigned t = u;
      if (order_mode == 1) t = span - 1 - u;
      else if (order_mode == 2) t = (u * mul + add) & (span - 1);
      if (t >= nt) continue;
      if (s->done[t]) continue;
      s->cur = t;
      tls.tid.x = t;
      TSAN(__tsan_switch_to_fiber(s->fib[t], 1);)
      simt_switch(&s->main_sp, s->sp[t]);
    }
    if (s->live && s->progress == before) die("deadlock: no thread of the CTA can make progress");
  }
  TSAN(for (unsigned t = 0; t < nt; t++) __tsan_destroy_fiber(s->fib[t]);)
}

// MS_EMUL_PROFILE=1: wall time per (grid, block) launch shape, printed at exit
struct ProfRow { unsigned grid, block; uint64_t n; double sec; };
static ProfRow g_prof[64];
static int g_nprof = 0;
static void prof_dump() {
  for (int i = 0; i < g_nprof; i++)
    fprintf(stderr, "simt profile: grid %5u block %4u launches %8llu total %.3f s (%.1f us each)\n", g_prof[i].grid,
            g_prof[i].block, (unsigned long long)g_prof[i].n, g_prof[i].sec, 1e6 * g_prof[i].sec / (double)g_prof[i].n);
}
static void launch_impl(unsigned grid, unsigned block, size_t dyn, const std::function<void()>& body);
void launch(unsigned grid, unsigned block, size_t dyn, const std::function<void()>& body) {
  static const bool profile = getenv("MS_EMUL_PROFILE") != nullptr;
  if (!profile) { launch_impl(grid, block, dyn, body); return; }
  const auto t0 = std::chrono::steady_clock::now();
  launch_impl(grid, block, dyn, body);
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  int i = 0;
  while (i < g_nprof && !(g_prof[i].grid == grid && g_prof[i].block == block)) i++;
  if (i == g_nprof) {
    if (g_nprof == 64) return;
    if (g_nprof == 0) atexit(prof_dump);
    g_prof[g_nprof++] = ProfRow{grid, block, 0, 0.0};
  }
  g_prof[i].n++;
  g_prof[i].sec += dt;
}
static void launch_impl(unsigned grid, unsigned block, size_t dyn, const std::function<void()>& body) {
  Sched* s = sched();
  if (s->body) die("nested kernel launch");
  if (block == 0 || block > kMaxThreads) die("bad block size");
  if (dyn > (256u << 10)) die("dynamic shared memory too large");
  s->body = &body;
  tls.bdim = dim3{block, 1, 1};
  tls.gdim = dim3{grid, 1, 1};
  tls.tid = uint3{0, 0, 0};
  static const int order_mode = getenv("MS_EMUL_ORDER") ? atoi(getenv("MS_EMUL_ORDER")) : 0;
  TSAN(__tsan_release(&s->sync_launch_begin);)
  for (unsigned bb = 0; bb < grid; bb++) {
    const unsigned b = order_mode ? grid - 1 - bb : bb;     // CTAs must not depend on launch order either
    tls.bid = uint3{b, 0, 0};
    memset(s->smem, 0xCD, dyn);
    run_cta(s, block);
  }
  TSAN(__tsan_acquire(&s->sync_launch_end);)
  s->body = nullptr;
}

}  // namespace simt

long long clock64() {
  return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

// -------------------------------------------------------------------- runtime API stubs
struct simt_event { std::chrono::steady_clock::time_point t; };
struct simt_stream { int unused; };

const char* cudaGetErrorString(cudaError_t e) {
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorNotSupported: return "operation not supported (emulation)";
    default: return "invalid value";
  }
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 8; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof(*p));
  strcpy(p->name, "simt-emulator");
  const char* e = getenv("MS_EMUL_SMS");
  p->multiProcessorCount = e ? atoi(e) : 1;
  p->sharedMemPerBlockOptin = 227 << 10;
  p->totalGlobalMem = (size_t)8 << 30;
  p->major = 10;
  return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t bytes) {
  void* q = nullptr;
  if (posix_memalign(&q, 256, bytes ? bytes : 1)) return cudaErrorMemoryAllocation;
  memset(q, 0xCD, bytes);   // device memory is not zeroed for you
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new simt_stream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* st) { *st = cudaStreamCaptureStatusNone; return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simt_event(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new simt_event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  memcpy(p, h.reserved, sizeof(*p));
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
