// index_kernels.cu — GPU-resident open-addressed hash index of (endpoint, block-hash)
// membership, sm_100a.
//
// Logical content = upstream's prefix indexer (hashToPods),
// physically a table keyed by block hash whose value is a bitset row over the local
// endpoints: one 128 B row read answers "which of 1024 endpoints hold this block".
// Keys live in buckets of 4 (one 32 B sector); linear probing over buckets.
// Inserts claim an EMPTY key with atomicCAS and give it the next NODE: nodes are
// numbered in insertion order and own the key's row (index_device.cuh), so a chain
// inserted in order occupies consecutive rows.  Membership bits flip with
// atomicOr/atomicAnd, cnt tracks the row popcount.
//
// Key presence is rmask[node] != 0: bit g says "rank g's row of this key is non-empty".
// With one rank that is just cnt > 0.  With an endpoint-range sharded pool every rank's
// table is a directory of the WHOLE pool's keys (rows only for its own endpoints): the
// owner of an endpoint applies the SET / CLEAR exactly (the row bit makes it idempotent)
// and logs the transitions of its row, empty -> non-empty (APPEAR) and back (VANISH); the
// other ranks replay those into their rmask (index_remote_*).  A key nobody holds any more
// becomes a tombstone and its node is retired (neither is reused until a rebuild compacts
// the live nodes in order), which keeps lookups exact without reading the row — and lets
// every rank find upstream's stopping point, the first block NO pod holds, on its own.
#include "index_device.cuh"
#include "kernels.cuh"

namespace fi {

namespace {

// find the table slot of h or claim an EMPTY one (*claimed = true: the caller allocates its node).
// SLOT_MISS on a full table.
__device__ uint32_t table_find_or_claim(const IndexView& ix, IndexCounters* ctr, uint64_t h, bool* claimed) {
  *claimed = false;
  uint64_t b = h & ix.bmask;
  for (uint64_t it = 0; it <= ix.bmask; ++it) {
    unsigned long long* kb = reinterpret_cast<unsigned long long*>(ix.keys + b * BUCKET_KEYS);
#pragma unroll 1
    for (int j = 0; j < BUCKET_KEYS; ++j) {
      unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(kb + j);
      if (k == h) return (uint32_t)(b * BUCKET_KEYS + j);
      if (k == KEY_EMPTY) {
        unsigned long long old = atomicCAS(kb + j, (unsigned long long)KEY_EMPTY, (unsigned long long)h);
        if (old == KEY_EMPTY) {
          *claimed = true;
          return (uint32_t)(b * BUCKET_KEYS + j);
        }
        if (old == h) return (uint32_t)(b * BUCKET_KEYS + j);
        // someone else claimed it for another key: keep scanning
      }
    }
    b = (b + 1) & ix.bmask;
  }
  atomicExch(&ctr->overflow, 1ull);
  return SLOT_MISS;
}

// Ordered reservation of one CTA pass: the flagged threads get CONSECUTIVE positions of *counter in
// thread (= op) order — one atomicAdd per CTA.  Used for the node numbers of new keys (the keys of a chain
// that arrives as consecutive ops end up next to each other in klog / rows) and for the APPEAR log (so that
// the ranks replaying it allocate consecutive nodes too).  Every thread of the CTA must call it.
__device__ unsigned long long cta_reserve(bool flag, unsigned long long* counter) {
  __shared__ uint32_t s_warp[8];
  __shared__ unsigned long long s_base;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xFFFFFFFFu, flag);
  const uint32_t rank_in_warp = __popc(m & ((1u << lane) - 1u));
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int w = 0; w < 8; ++w) {
      const uint32_t c = s_warp[w];
      s_warp[w] = tot;
      tot += c;
    }
    s_base = tot ? atomicAdd(counter, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  const unsigned long long pos = flag ? s_base + s_warp[warp] + rank_in_warp : ~0ull;
  __syncthreads();  // s_warp / s_base are reused by the next reservation
  return pos;
}

// node of a slot some other thread claimed: wait until that thread has published it
__device__ __forceinline__ uint32_t wait_node(const IndexView& ix, uint32_t slot) {
  const volatile uint32_t* p = ix.node_of + slot;
  uint32_t n;
  while ((n = *p) == NODE_INVALID) __nanosleep(20);
  return n;
}

// REMOTE = false: this rank's own SET ops (row bit, popcount, APPEAR log).
// REMOTE = true:  replay of rank `rank`'s APPEAR log: directory entry only (rmask bit of that rank).
template <bool REMOTE>
__global__ void __launch_bounds__(256) index_set_kernel(IndexView ix, IndexCounters* ctr, const fi_index_op* __restrict__ ops,
                                                        const uint64_t* __restrict__ hashes, uint64_t n, uint32_t ep_begin,
                                                        uint32_t ep_count, uint32_t rank, GossipLog log) {
  const uint32_t rbit = 1u << rank;
  const uint32_t zero_node = (uint32_t)(ix.C + 2);
  // every thread of a CTA runs the same number of passes (block-wide barriers inside)
  for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x; base < n; base += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = base + threadIdx.x;
    uint64_t hsh = 0;
    bool active = false;
    uint32_t e = 0;
    if (i < n) {
      if (REMOTE) {
        hsh = hashes[i];
        active = true;
      } else {
        const fi_index_op op = ops[i];
        hsh = op.hash;
        e = op.endpoint - ep_begin;
        active = op.op == FI_OP_SET && e < ep_count;
      }
    }
    // 1. table slot (claim an empty one for a new key)
    bool claimed = false;
    uint32_t slot = SLOT_MISS, node = NODE_INVALID;
    if (active) {
      if (key_is_special(hsh)) node = (uint32_t)(ix.C + (hsh == KEY_TOMB ? 1 : 0));
      else slot = table_find_or_claim(ix, ctr, hsh, &claimed);
    }
    // 2. new keys get consecutive nodes in op order and publish them
    const unsigned long long mine = cta_reserve(claimed, &ctr->used);
    if (claimed) {
      if (mine < ix.C) {
        node = (uint32_t)mine;
        ix.klog[node] = hsh;
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(ix.node_of + slot) = node;
      } else {
        // out of nodes (reported through ctr->overflow): retire the slot again so that "key in the table ⇔
        // some rank holds it" keeps holding, and release the threads waiting for its node
        atomicExch(&ctr->overflow, 1ull);
        ix.keys[slot] = KEY_TOMB;
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(ix.node_of + slot) = zero_node;
      }
    }
    // 3. keys that were already there (possibly claimed by another CTA a moment ago): their node.  No
    // thread waits before it has published its own nodes, so the waits cannot form a cycle.
    if (active && !claimed && slot != SLOT_MISS) node = wait_node(ix, slot);
    bool appear = false;
    if (active && node != NODE_INVALID && node != zero_node) {
      if (REMOTE) {
        atomicOr(ix.rmask + node, rbit);
      } else {
        const uint32_t bit = 1u << (e & 31);
        const uint32_t old = atomicOr(ix.rows + ((uint64_t)node << ix.logW) + (e >> 5), bit);
        if (!(old & bit) && atomicAdd(ix.cnt + node, 1u) == 0u) {  // this rank's row: empty -> non-empty
          atomicOr(ix.rmask + node, rbit);
          appear = true;
        }
      }
    }
    if (!REMOTE && log.n_appear) {  // warp-uniform: sharded pools only
      const unsigned long long pos = cta_reserve(appear, log.n_appear);
      if (appear && pos < log.cap) log.appear[pos] = hsh;
    }
  }
}

// table slot of a regular key (not its node): the clear path retires the slot
__device__ uint32_t table_find_slot(const IndexView& ix, uint64_t h) {
  uint64_t b = h & ix.bmask;
  for (uint64_t it = 0; it <= ix.bmask; ++it) {
    const BucketRegs r = bucket_load(ix, b);
    const int j = bucket_scan(r, h);
    if (j < BUCKET_KEYS) return (uint32_t)(b * BUCKET_KEYS + j);
    if (j == BUCKET_KEYS) return SLOT_MISS;
    b = (b + 1) & ix.bmask;
  }
  return SLOT_MISS;
}

// rank `rbit`'s row of the key at (slot, node) has emptied: drop its directory bit; if no rank holds the key
// any more retire the key and its node
__device__ __forceinline__ void rank_vanished(const IndexView& ix, IndexCounters* ctr, uint32_t slot, uint32_t node,
                                              uint32_t rbit) {
  const uint32_t oldm = atomicAnd(ix.rmask + node, ~rbit);
  if ((oldm & rbit) && (oldm & ~rbit) == 0u && slot != SLOT_MISS) {
    ix.keys[slot] = KEY_TOMB;
    ix.klog[node] = 0;
    atomicAdd(&ctr->tombstones, 1ull);
  }
}

template <bool REMOTE>
__global__ void __launch_bounds__(256) index_clear_kernel(IndexView ix, IndexCounters* ctr, const fi_index_op* __restrict__ ops,
                                                          const uint64_t* __restrict__ hashes, uint64_t n, uint32_t ep_begin,
                                                          uint32_t ep_count, uint32_t rank, GossipLog log,
                                                          const unsigned long long* __restrict__ n_dev) {
  const uint32_t rbit = 1u << rank;
  if (n_dev) {  // op count produced on the device (lru_evict_kernel): n is only the buffer's capacity
    const unsigned long long nd = *n_dev;
    n = nd < n ? nd : n;
  }
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t hsh;
    uint32_t e = 0;
    if (REMOTE) {
      hsh = hashes[i];
    } else {
      const fi_index_op op = ops[i];
      e = op.endpoint - ep_begin;
      if (op.op != FI_OP_CLEAR || e >= ep_count) continue;
      hsh = op.hash;
    }
    uint32_t slot = SLOT_MISS, node;
    if (key_is_special(hsh)) {
      node = (uint32_t)(ix.C + (hsh == KEY_TOMB ? 1 : 0));
    } else {
      slot = table_find_slot(ix, hsh);
      if (slot == SLOT_MISS) continue;
      node = ix.node_of[slot];
      if (node >= ix.C) continue;  // a slot parked by a node overflow
    }
    if (REMOTE) {
      rank_vanished(ix, ctr, slot, node, rbit);
      continue;
    }
    const uint32_t bit = 1u << (e & 31);
    const uint32_t old = atomicAnd(ix.rows + ((uint64_t)node << ix.logW) + (e >> 5), ~bit);
    if ((old & bit) && atomicSub(ix.cnt + node, 1u) == 1u) {  // this rank's row emptied
      rank_vanished(ix, ctr, slot, node, rbit);
      if (log.n_vanish) {
        const unsigned long long pos = atomicAdd(log.n_vanish, 1ull);
        if (pos < log.cap) log.vanish[pos] = hsh;
      }
    }
  }
}

// Re-insert every live node of `from` into the (fresh) index `to`, in node order, so that runs of
// consecutive nodes stay consecutive (each CTA pass compacts 256 consecutive old nodes into one range).
__global__ void __launch_bounds__(256) index_rebuild_kernel(IndexView from, IndexView to, IndexCounters* ctr) {
  for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x; base < from.C; base += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t s = base + threadIdx.x;
    uint64_t h = 0;
    if (s < from.C) h = from.klog[s];
    bool claimed = false;
    uint32_t slot = SLOT_MISS;
    if (h != 0) slot = table_find_or_claim(to, ctr, h, &claimed);  // keys are unique: always a fresh claim
    const unsigned long long d = cta_reserve(claimed, &ctr->used);
    if (claimed && d < to.C) {
      to.klog[d] = h;
      to.node_of[slot] = (uint32_t)d;
      to.cnt[d] = from.cnt[s];
      to.rmask[d] = from.rmask[s];
      const uint32_t* src = from.rows + (s << from.logW);
      uint32_t* dst = to.rows + (d << to.logW);
      for (uint32_t w = 0; w < from.W; ++w) dst[w] = src[w];
    }
  }
  // the two special nodes keep their place
  if (blockIdx.x == 0 && threadIdx.x < 2) {
    const uint64_t s = from.C + threadIdx.x;
    if (from.rmask[s]) {
      to.cnt[s] = from.cnt[s];
      to.rmask[s] = from.rmask[s];
      for (uint32_t w = 0; w < from.W; ++w) to.rows[(s << to.logW) + w] = from.rows[(s << from.logW) + w];
    }
  }
}

// membership query (tests / diagnostics): out[i] = 1 iff (endpoint, hash) is present
__global__ void __launch_bounds__(256) index_contains_kernel(IndexView ix, const fi_index_op* __restrict__ q, uint64_t n,
                                                             uint32_t ep_begin, uint32_t ep_count,
                                                             uint8_t* __restrict__ out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t e = q[i].endpoint - ep_begin;
    uint8_t r = 0;
    if (e < ep_count) {
      const uint32_t node = index_find(ix, q[i].hash);
      if (node != SLOT_MISS) r = (ix.rows[((uint64_t)node << ix.logW) + (e >> 5)] >> (e & 31)) & 1u;
    }
    out[i] = r;
  }
}

inline unsigned grid_for(uint64_t n) {
  uint64_t g = (n + 255) / 256;
  if (g > 148ull * 16) g = 148ull * 16;
  if (g == 0) g = 1;
  return (unsigned)g;
}

}  // namespace

cudaError_t launch_index_set(IndexView ix, IndexCounters* ctr, const fi_index_op* ops, uint64_t n, uint32_t ep_begin,
                             uint32_t ep_count, uint32_t rank, GossipLog log, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  index_set_kernel<false><<<grid_for(n), 256, 0, s>>>(ix, ctr, ops, nullptr, n, ep_begin, ep_count, rank, log);
  return cudaGetLastError();
}

cudaError_t launch_index_clear(IndexView ix, IndexCounters* ctr, const fi_index_op* ops, uint64_t n,
                               uint32_t ep_begin, uint32_t ep_count, uint32_t rank, GossipLog log, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  index_clear_kernel<false><<<grid_for(n), 256, 0, s>>>(ix, ctr, ops, nullptr, n, ep_begin, ep_count, rank, log, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_index_clear_counted(IndexView ix, IndexCounters* ctr, const fi_index_op* ops, uint64_t cap,
                                       const unsigned long long* n_dev, uint32_t ep_begin, uint32_t ep_count, uint32_t rank,
                                       GossipLog log, cudaStream_t s) {
  if (cap == 0) return cudaSuccess;
  index_clear_kernel<false><<<grid_for(cap), 256, 0, s>>>(ix, ctr, ops, nullptr, cap, ep_begin, ep_count, rank, log, n_dev);
  return cudaGetLastError();
}

cudaError_t launch_index_remote_appear(IndexView ix, IndexCounters* ctr, const uint64_t* hashes, uint64_t n, uint32_t rank,
                                       cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  index_set_kernel<true><<<grid_for(n), 256, 0, s>>>(ix, ctr, nullptr, hashes, n, 0, 0, rank, GossipLog{});
  return cudaGetLastError();
}

cudaError_t launch_index_remote_vanish(IndexView ix, IndexCounters* ctr, const uint64_t* hashes, uint64_t n, uint32_t rank,
                                       cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  index_clear_kernel<true><<<grid_for(n), 256, 0, s>>>(ix, ctr, nullptr, hashes, n, 0, 0, rank, GossipLog{}, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_index_rebuild(IndexView from, IndexView to, IndexCounters* ctr, cudaStream_t s) {
  index_rebuild_kernel<<<grid_for(from.C), 256, 0, s>>>(from, to, ctr);
  return cudaGetLastError();
}

cudaError_t launch_index_contains(IndexView ix, const fi_index_op* q, uint64_t n, uint32_t ep_begin,
                                  uint32_t ep_count, uint8_t* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  index_contains_kernel<<<grid_for(n), 256, 0, s>>>(ix, q, n, ep_begin, ep_count, out);
  return cudaGetLastError();
}

}  // namespace fi
