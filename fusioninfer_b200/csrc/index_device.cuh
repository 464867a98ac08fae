// index_device.cuh — device-side lookup of the hash index (see index_kernels.cu).
//
// Two levels: the open-addressed TABLE (keys[slot], node_of[slot]; buckets of 4 keys = one 32-byte
// sector, linear probing over buckets) maps a block hash to a NODE; nodes are numbered in insertion
// order and own the data: klog[node] = the key, rows[node] = the membership bitset over the LOCAL endpoints,
// cnt[node] = its popcount, rmask[node] = the ranks whose row of this key is non-empty (key present ⇔
// rmask != 0; a single rank: rmask = (cnt > 0)).  Lookups return node ids.  Because a prompt's block hashes are inserted in chain order
// (upstream PreRequest: indexer.Add(hashes, pod)), the nodes of a cached prefix are consecutive: the
// match kernel verifies "node of block i+1 == node of block i + 1" with one coalesced read of klog
// instead of hashing into the table for every block (index_kernels.cu, match_kernels.cu).
#pragma once
#include "kernels.cuh"

namespace fi {

constexpr uint32_t NODE_INVALID = 0xFFFFFFFFu;  // node_of[] of a slot whose node is not published yet

struct BucketRegs {
  uint4 q[BUCKET_KEYS / 2];  // the bucket's BUCKET_KEYS keys
  uint4 nodes;               // and their nodes (node_of[b*4 .. b*4+3]) — loaded in parallel, not after the scan
};
static_assert(BUCKET_KEYS == 4, "BucketRegs::nodes holds exactly four node ids");

__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// issue the loads of bucket b (no dependence on their result)
__device__ __forceinline__ BucketRegs bucket_load(const IndexView& ix, uint64_t b) {
  const uint4* p = reinterpret_cast<const uint4*>(ix.keys + b * BUCKET_KEYS);
  BucketRegs r;
#pragma unroll
  for (int i = 0; i < BUCKET_KEYS / 2; ++i) r.q[i] = __ldg(p + i);
  r.nodes = __ldg(reinterpret_cast<const uint4*>(ix.node_of + b * BUCKET_KEYS));
  return r;
}

// 0..BUCKET_KEYS-1: position of h; BUCKET_KEYS: the bucket has an EMPTY key (definite miss);
// BUCKET_KEYS+1: full without a match, keep probing
__device__ __forceinline__ int bucket_scan(const BucketRegs& r, uint64_t h) {
  int pos = BUCKET_KEYS + 1;
  bool empty = false;
#pragma unroll
  for (int i = BUCKET_KEYS / 2 - 1; i >= 0; --i) {
    const uint64_t k0 = u64_of(r.q[i].x, r.q[i].y), k1 = u64_of(r.q[i].z, r.q[i].w);
    empty |= (k0 == KEY_EMPTY) | (k1 == KEY_EMPTY);
    if (k1 == h) pos = 2 * i + 1;
    if (k0 == h) pos = 2 * i;
  }
  if (pos <= BUCKET_KEYS - 1) return pos;
  return empty ? BUCKET_KEYS : BUCKET_KEYS + 1;
}

__device__ __forceinline__ uint32_t bucket_node(const BucketRegs& r, int j) {
  return j == 0 ? r.nodes.x : j == 1 ? r.nodes.y : j == 2 ? r.nodes.z : r.nodes.w;
}

// out-of-line continuation of a lookup whose home bucket was full without a match
static __device__ __noinline__ uint32_t index_resolve_overflow(const IndexView ix, uint64_t h) {
  uint64_t b = h & ix.bmask;
  for (uint64_t it = 0; it < ix.bmask; ++it) {
    b = (b + 1) & ix.bmask;
    const BucketRegs r = bucket_load(ix, b);
    const int j = bucket_scan(r, h);
    if (j < BUCKET_KEYS) return bucket_node(r, j);
    if (j == BUCKET_KEYS) return SLOT_MISS;
  }
  return SLOT_MISS;
}

// finish a lookup whose home bucket was loaded into `first`: node of h, or SLOT_MISS
__device__ __forceinline__ uint32_t index_resolve(const IndexView& ix, uint64_t h, const BucketRegs& first) {
  const int j = bucket_scan(first, h);
  if (j < BUCKET_KEYS) return bucket_node(first, j);
  if (j == BUCKET_KEYS) return SLOT_MISS;
  return index_resolve_overflow(ix, h);  // rare: home bucket full
}

__device__ __forceinline__ bool key_is_special(uint64_t h) { return h == KEY_EMPTY || h == KEY_TOMB; }

// node holding key h, whatever its row (regular keys: present ⇔ row non-empty).  The hashes 0 and ~0
// (the table's EMPTY / TOMB markers) own the fixed nodes C and C+1.
__device__ __forceinline__ uint32_t index_find_key(const IndexView& ix, uint64_t h) {
  if (key_is_special(h)) return (uint32_t)(ix.C + (h == KEY_TOMB ? 1 : 0));
  return index_resolve(ix, h, bucket_load(ix, h & ix.bmask));
}

// node of h if at least one endpoint of the pool holds it (rmask != 0 ⇔ the key is in the table), else SLOT_MISS
__device__ __forceinline__ uint32_t index_find(const IndexView& ix, uint64_t h) {
  if (key_is_special(h)) {
    const uint64_t s = ix.C + (h == KEY_TOMB ? 1 : 0);
    return ix.rmask[s] ? (uint32_t)s : SLOT_MISS;
  }
  return index_resolve(ix, h, bucket_load(ix, h & ix.bmask));
}

// Keys-only variant for the fallback where every lane of a chunk probes the table (a prefix whose nodes are
// scattered): the node is fetched with a second, dependent read on a hit — a miss costs one DRAM transaction
// instead of two.
__device__ __forceinline__ BucketRegs bucket_load_keys(const IndexView& ix, uint64_t b) {
  const uint4* p = reinterpret_cast<const uint4*>(ix.keys + b * BUCKET_KEYS);
  BucketRegs r;
#pragma unroll
  for (int i = 0; i < BUCKET_KEYS / 2; ++i) r.q[i] = __ldg(p + i);
  r.nodes = make_uint4(0, 0, 0, 0);
  return r;
}
__device__ __forceinline__ uint32_t index_find_lazy(const IndexView& ix, uint64_t h) {
  if (key_is_special(h)) {
    const uint64_t s = ix.C + (h == KEY_TOMB ? 1 : 0);
    return ix.rmask[s] ? (uint32_t)s : SLOT_MISS;
  }
  uint64_t b = h & ix.bmask;
  for (uint64_t it = 0; it <= ix.bmask; ++it) {
    const BucketRegs r = bucket_load_keys(ix, b);
    const int j = bucket_scan(r, h);
    if (j < BUCKET_KEYS) return __ldg(ix.node_of + b * BUCKET_KEYS + j);
    if (j == BUCKET_KEYS) return SLOT_MISS;
    b = (b + 1) & ix.bmask;
  }
  return SLOT_MISS;
}

// Speculation: is `cand` the node of the regular key h?  (klog of a free or retired node is 0, and a
// regular key is never 0.)
__device__ __forceinline__ bool node_holds(const IndexView& ix, uint32_t cand, uint64_t h) {
  return cand < ix.C && __ldg(ix.klog + cand) == h;
}

}  // namespace fi
