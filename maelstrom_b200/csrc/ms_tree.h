// ms_tree.h -- the persistent hash tree of demo/ruby/datomic_list_append.rb (class Tree / Leaf / Branch, :47-320) as
// arithmetic on tree SHAPES: which keys a leaf holds, which ranges and child pointers a branch holds.  The values
// (the appended lists) never influence which messages a node sends, so they are not kept here; a caller replays
// apply_txn over the chain of committed roots (as for the single-key variant).
//
// One header for the CUDA engine (csrc/ms_raft.cuh, thread 0 of a node's CTA) and for the CPU oracle
// (oracle/oracle.cpp): the tree arithmetic is a pure function of (root, micro-ops, what the node may look at), and
// restating a pure function twice adds nothing; what the two sides restate separately, and what the parity tests
// compare, is the message-level behaviour around it (RPC sequencing, queueing, retries).  tests/test_txn_tree.py pins
// the pieces that have known answers: Zlib.crc32 of the key's decimal string, range splitting, pointer numbering.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(MS_EMUL)
#define MST_HD __host__ __device__
#else
#define MST_HD
#endif

namespace mst {

constexpr uint32_t kRing = 128;        // Tree::RING_SIZE, :52
constexpr uint32_t kBranch = 8;        // Tree::BRANCH_FACTOR, :55
constexpr uint32_t kMaxKeys = 20;      // keys a leaf record can hold here (the reference: unbounded); more = capacity error
constexpr uint32_t kMaxDepth = 16;     // branches above a leaf (the reference: unbounded)
constexpr uint32_t kMaxOps = 4;        // micro-ops per txn carried in the 64-bit payload
constexpr uint32_t kPtrEmpty = 1;      // the literal pointer "empty" (Tree.empty, :65-67); 0 = no pointer

// Immutable tree node, 64 B.  type 1 = leaf {range, keys}, 2 = branch {range, (upper bound, child pointer) x 8}.
struct Rec {
  uint8_t type, lo, hi, n;             // range [lo, hi) of hash values, hi <= 128; n = keys of a leaf
  union {
    uint16_t keys[kMaxKeys];
    struct { uint8_t upper[kBranch]; uint32_t child[kBranch]; } b;
  };
  uint8_t pad[64 - 4 - 2 * kMaxKeys];
};
static_assert(sizeof(Rec) == 64, "tree records are 64 bytes");

// micro-op k of a txn payload: 16 bits = valid << 15 | append << 14 | key (14 bits)
MST_HD inline bool op_valid(uint64_t ops, uint32_t k) { return ((ops >> (16 * k)) & 0x8000u) != 0; }
MST_HD inline bool op_append(uint64_t ops, uint32_t k) { return ((ops >> (16 * k)) & 0x4000u) != 0; }
MST_HD inline uint32_t op_key(uint64_t ops, uint32_t k) { return (uint32_t)(ops >> (16 * k)) & 0x3FFFu; }

// Tree.hash (:59-61): Zlib.crc32(k.to_s) % RING_SIZE, k a non-negative integer
MST_HD inline uint32_t key_hash(uint32_t k) {
  char s[12];
  int n = 0;
  do { s[n++] = (char)('0' + k % 10); k /= 10; } while (k);
  uint32_t crc = 0xFFFFFFFFu;
  for (int i = n - 1; i >= 0; i--) {
    crc ^= (uint8_t)s[i];
    for (int b = 0; b < 8; b++) crc = (crc >> 1) ^ (0xEDB88320u & (0u - (crc & 1u)));
  }
  return (crc ^ 0xFFFFFFFFu) % kRing;
}

// Pointers are "#{node_id}-#{p}" (new_ptr, :355-358) with p = 1, 2, ... per node: here 2 + node * per_node + (p - 1).
MST_HD inline uint32_t ptr_of(uint32_t node, uint32_t per_node, uint32_t p) { return 2u + node * per_node + (p - 1u); }
MST_HD inline bool ptr_is_new(uint32_t ptr, uint32_t node, uint32_t per_node, uint32_t start_counter) {
  // made by `node` after its counter stood at start_counter, i.e. inside the transaction being evaluated
  if (ptr < 2u) return false;
  const uint32_t x = ptr - 2u;
  return x / per_node == node && x % per_node >= start_counter;
}

enum Status { kDone = 0, kNeedLoad = 1, kCapacity = 2 };

// What a node may look at: Store::rec(ptr) -> Rec* (contents by pointer; immutable once made),
// Store::cached(ptr) -> has Tree.load put it into this node's @@cache (:83-101)?
//
// apply_txn (:395-417) on shapes.  Evaluates the micro-ops against the tree `root1`; nodes made on the way get the
// pointers counter + 1, counter + 2, ... in the order the reference's recursion calls new_ptr (deepest first; a split
// numbers its eight leaves, then the branch above them).  Stops at the first node it may not look at yet
// (Branch#branch_index / Tree.load would block on a read of lww-kv there): the caller reads it, puts it into the cache,
// resets the counter to start_counter and evaluates again -- the result does not depend on where it was interrupted.
template <class Store>
MST_HD inline Status apply_txn(Store& S, uint32_t node, uint32_t per_node, uint32_t root1, uint64_t ops,
                               uint32_t start_counter, uint32_t& counter, uint32_t& root2, uint32_t& load_ptr) {
  auto accessible = [&](uint32_t ptr) { return ptr_is_new(ptr, node, per_node, start_counter) || S.cached(ptr); };
  uint32_t t = root1;
  if (!accessible(t)) { load_ptr = t; return kNeedLoad; }              // current_tree: Tree.load of the root (:361-368)
  for (uint32_t k = 0; k < kMaxOps; k++) {
    if (!op_valid(ops, k)) continue;
    const uint32_t key = op_key(ops, k), h = key_hash(key);
    uint32_t path_ptr[kMaxDepth];
    uint8_t path_i[kMaxDepth];
    uint32_t depth = 0, cur = t;
    for (;;) {                                                        // Branch#[] / Branch#assoc: branch_index, :229-247
      if (!accessible(cur)) { load_ptr = cur; return kNeedLoad; }
      const Rec* r = S.rec(cur);
      if (r->type != 2) break;
      uint32_t i = 0;
      while (i + 1 < kBranch && !(h < r->b.upper[i])) i++;
      if (depth == kMaxDepth) return kCapacity;
      path_ptr[depth] = cur; path_i[depth] = (uint8_t)i; depth++;
      cur = r->b.child[i];
    }
    if (!op_append(ops, k)) continue;                                 // "r": t[k], the tree stays (:401-403)
    // Leaf#assoc (:163-199)
    const Rec leaf = *S.rec(cur);
    bool has = false;
    for (uint32_t j = 0; j < leaf.n; j++) has = has || leaf.keys[j] == key;
    uint32_t child;
    if (has || leaf.n < kBranch) {
      if (counter >= per_node) return kCapacity;
      Rec nl = leaf;
      if (!has) { if (nl.n >= kMaxKeys) return kCapacity; nl.keys[nl.n++] = (uint16_t)key; }
      child = ptr_of(node, per_node, ++counter);
      *S.rec(child) = nl;
    } else {
      // replace the leaf by a branch over eight leaves that split its range (:170-197)
      if (leaf.n + 1u > kMaxKeys || counter + kBranch + 1u > per_node) return kCapacity;
      uint16_t all[kMaxKeys + 1];
      for (uint32_t j = 0; j < leaf.n; j++) all[j] = leaf.keys[j];
      all[leaf.n] = (uint16_t)key;
      const uint32_t lower = leaf.lo, upper = leaf.hi, bs = (upper - lower) / kBranch;
      Rec br{};
      br.type = 2; br.lo = leaf.lo; br.hi = leaf.hi; br.n = (uint8_t)kBranch;
      for (uint32_t i = 0; i < kBranch; i++) {
        const uint32_t b_lo = lower + i * bs, b_hi = (i == kBranch - 1) ? upper : b_lo + bs;
        Rec lf{};
        lf.type = 1; lf.lo = (uint8_t)b_lo; lf.hi = (uint8_t)b_hi; lf.n = 0;
        for (uint32_t j = 0; j <= leaf.n; j++) {
          const uint32_t hj = key_hash(all[j]);
          if (b_lo <= hj && hj < b_hi) lf.keys[lf.n++] = all[j];
        }
        const uint32_t lp = ptr_of(node, per_node, ++counter);
        *S.rec(lp) = lf;
        br.b.upper[i] = (uint8_t)b_hi;
        br.b.child[i] = lp;
      }
      child = ptr_of(node, per_node, ++counter);
      *S.rec(child) = br;
    }
    // Branch#assoc on the way back up (:249-259): a copy of every branch on the path, pointing at the new child
    for (uint32_t d = depth; d-- > 0;) {
      if (counter >= per_node) return kCapacity;
      Rec nb = *S.rec(path_ptr[d]);
      nb.b.child[path_i[d]] = child;
      child = ptr_of(node, per_node, ++counter);
      *S.rec(child) = nb;
    }
    t = child;
  }
  root2 = t;
  return kDone;
}

constexpr uint32_t kMaxWrites = 4 * (kMaxDepth + kBranch + 1);   // nodes one txn can make

// Branch#save! / Leaf#save! (:202-213, :282-311): the unsaved nodes reachable from `root` through nodes made in this
// transaction, children before their parent, in child order -- the order in which the writes to lww-kv go out.
// Fills out[0..n); returns false when the walk is deeper than the stack or out is full.
template <class Store>
MST_HD inline bool save_order(Store& S, uint32_t node, uint32_t per_node, uint32_t start_counter, uint32_t root,
                              uint32_t* out, uint32_t& n) {
  uint32_t st_ptr[kMaxDepth + 2];
  uint8_t st_i[kMaxDepth + 2];
  uint32_t sp = 0;
  n = 0;
  if (!ptr_is_new(root, node, per_node, start_counter)) return true;   // saved already: nothing to do
  st_ptr[sp] = root; st_i[sp] = 0; sp++;
  while (sp) {
    const uint32_t ptr = st_ptr[sp - 1];
    const Rec* r = S.rec(ptr);
    bool pushed = false;
    if (r->type == 2) {
      while (st_i[sp - 1] < kBranch) {
        const uint32_t c = r->b.child[st_i[sp - 1]++];
        if (ptr_is_new(c, node, per_node, start_counter)) {
          if (sp == kMaxDepth + 2) return false;
          st_ptr[sp] = c; st_i[sp] = 0; sp++;
          pushed = true;
          break;
        }
      }
    }
    if (pushed) continue;
    if (n == kMaxWrites) return false;
    out[n++] = ptr;
    sp--;
  }
  return true;
}

}  // namespace mst
