// index_kernels.cu — GPU-resident open-addressed hash index of (endpoint, block-hash)
// membership, sm_100a.
//
// Logical content = upstream's prefix indexer (hashToPods),
// physically a table keyed by block hash whose value is a bitset row over the local
// endpoints: one 128 B row read answers "which of 1024 endpoints hold this block".
// Keys live in buckets of 4 (one 32 B sector); linear probing over buckets.
// Inserts claim an EMPTY key with atomicCAS, membership bits flip with
// atomicOr/atomicAnd, cnt tracks the row popcount so that "key present ⇔ row
// non-empty" holds: a key whose row empties becomes a tombstone (never reused
// until a rebuild), which keeps lookups exact without reading the row.
#include "index_device.cuh"
#include "kernels.cuh"

namespace fi {

namespace {

// find the slot of h or claim an EMPTY one.  SLOT_MISS on a full table.
__device__ uint32_t index_find_or_claim(const IndexView& ix, IndexCounters* ctr, uint64_t h) {
  if (h == KEY_EMPTY) return (uint32_t)ix.C;
  if (h == KEY_TOMB) return (uint32_t)(ix.C + 1);
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
          atomicAdd(&ctr->used, 1ull);
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

__global__ void __launch_bounds__(256) index_set_kernel(IndexView ix, IndexCounters* ctr,
                                                        const fi_index_op* __restrict__ ops, uint64_t n,
                                                        uint32_t ep_begin, uint32_t ep_count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const fi_index_op op = ops[i];
    const uint32_t e = op.endpoint - ep_begin;
    if (op.op != FI_OP_SET || e >= ep_count) continue;
    const uint32_t slot = index_find_or_claim(ix, ctr, op.hash);
    if (slot == SLOT_MISS) continue;
    const uint32_t bit = 1u << (e & 31);
    const uint32_t old = atomicOr(ix.rows + ((uint64_t)slot << ix.logW) + (e >> 5), bit);
    if (!(old & bit)) atomicAdd(ix.cnt + slot, 1u);
  }
}

__global__ void __launch_bounds__(256) index_clear_kernel(IndexView ix, IndexCounters* ctr,
                                                          const fi_index_op* __restrict__ ops, uint64_t n,
                                                          uint32_t ep_begin, uint32_t ep_count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const fi_index_op op = ops[i];
    const uint32_t e = op.endpoint - ep_begin;
    if (op.op != FI_OP_CLEAR || e >= ep_count) continue;
    const uint32_t slot = index_find_key(ix, op.hash);
    if (slot == SLOT_MISS) continue;
    const uint32_t bit = 1u << (e & 31);
    const uint32_t old = atomicAnd(ix.rows + ((uint64_t)slot << ix.logW) + (e >> 5), ~bit);
    if (old & bit) {
      const uint32_t c = atomicSub(ix.cnt + slot, 1u);
      if (c == 1u && slot < ix.C) {  // row emptied: retire the key
        ix.keys[slot] = KEY_TOMB;
        atomicAdd(&ctr->tombstones, 1ull);
      }
    }
  }
}

// re-insert every live key of `from` into the (zeroed) table `to`, moving its row
__global__ void __launch_bounds__(256) index_rebuild_kernel(IndexView from, IndexView to, IndexCounters* ctr) {
  const uint64_t total = from.C + 2;
  for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < total;
       s += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t h;
    if (s >= from.C) {
      if (from.cnt[s] == 0) continue;
      h = (s == from.C) ? KEY_EMPTY : KEY_TOMB;
    } else {
      h = from.keys[s];
      if (h == KEY_EMPTY || h == KEY_TOMB) continue;
    }
    const uint32_t d = index_find_or_claim(to, ctr, h);
    if (d == SLOT_MISS) continue;
    to.cnt[d] = from.cnt[s];
    const uint32_t* src = from.rows + (s << from.logW);
    uint32_t* dst = to.rows + ((uint64_t)d << to.logW);
    for (uint32_t w = 0; w < from.W; ++w) dst[w] = src[w];
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
      const uint32_t slot = index_find(ix, q[i].hash);
      if (slot != SLOT_MISS) r = (ix.rows[((uint64_t)slot << ix.logW) + (e >> 5)] >> (e & 31)) & 1u;
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
                             uint32_t ep_count, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  index_set_kernel<<<grid_for(n), 256, 0, s>>>(ix, ctr, ops, n, ep_begin, ep_count);
  return cudaGetLastError();
}

cudaError_t launch_index_clear(IndexView ix, IndexCounters* ctr, const fi_index_op* ops, uint64_t n,
                               uint32_t ep_begin, uint32_t ep_count, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  index_clear_kernel<<<grid_for(n), 256, 0, s>>>(ix, ctr, ops, n, ep_begin, ep_count);
  return cudaGetLastError();
}

cudaError_t launch_index_rebuild(IndexView from, IndexView to, IndexCounters* ctr, cudaStream_t s) {
  index_rebuild_kernel<<<grid_for(from.C + 2), 256, 0, s>>>(from, to, ctr);
  return cudaGetLastError();
}

cudaError_t launch_index_contains(IndexView ix, const fi_index_op* q, uint64_t n, uint32_t ep_begin,
                                  uint32_t ep_count, uint8_t* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  index_contains_kernel<<<grid_for(n), 256, 0, s>>>(ix, q, n, ep_begin, ep_count, out);
  return cudaGetLastError();
}

}  // namespace fi
