// index_device.cuh — device-side lookup of the hash index (see index_kernels.cu).
#pragma once
#include "kernels.cuh"

namespace fi {

struct BucketRegs {
  uint4 a, b;  // 4 keys = one 32-byte sector
};

__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// issue the loads of the home bucket of h (no dependence on their result)
__device__ __forceinline__ BucketRegs bucket_load(const IndexView& ix, uint64_t b) {
  const uint4* p = reinterpret_cast<const uint4*>(ix.keys + b * BUCKET_KEYS);
  BucketRegs r;
  r.a = __ldg(p);
  r.b = __ldg(p + 1);
  return r;
}

// 0..3: position of h in the bucket; 4: bucket has an EMPTY key (definite miss); 5: full, keep probing
__device__ __forceinline__ int bucket_scan(const BucketRegs& r, uint64_t h) {
  const uint64_t k0 = u64_of(r.a.x, r.a.y), k1 = u64_of(r.a.z, r.a.w);
  const uint64_t k2 = u64_of(r.b.x, r.b.y), k3 = u64_of(r.b.z, r.b.w);
  if (k0 == h) return 0;
  if (k1 == h) return 1;
  if (k2 == h) return 2;
  if (k3 == h) return 3;
  if (k0 == KEY_EMPTY || k1 == KEY_EMPTY || k2 == KEY_EMPTY || k3 == KEY_EMPTY) return 4;
  return 5;
}

// finish a lookup whose home bucket was loaded into `first`
__device__ __forceinline__ uint32_t index_resolve(const IndexView& ix, uint64_t h, const BucketRegs& first) {
  uint64_t b = h & ix.bmask;
  int j = bucket_scan(first, h);
  if (j < 4) return (uint32_t)(b * BUCKET_KEYS + j);
  if (j == 4) return SLOT_MISS;
  for (uint64_t it = 0; it < ix.bmask; ++it) {  // rare: home bucket full
    b = (b + 1) & ix.bmask;
    const BucketRegs r = bucket_load(ix, b);
    j = bucket_scan(r, h);
    if (j < 4) return (uint32_t)(b * BUCKET_KEYS + j);
    if (j == 4) return SLOT_MISS;
  }
  return SLOT_MISS;
}

__device__ __forceinline__ bool key_is_special(uint64_t h) { return h == KEY_EMPTY || h == KEY_TOMB; }

// slot holding key h, whatever its row (regular keys: present ⇒ row non-empty)
__device__ __forceinline__ uint32_t index_find_key(const IndexView& ix, uint64_t h) {
  if (key_is_special(h)) return (uint32_t)(ix.C + (h == KEY_TOMB ? 1 : 0));
  return index_resolve(ix, h, bucket_load(ix, h & ix.bmask));
}

// slot of h if at least one local endpoint holds it, else SLOT_MISS
__device__ __forceinline__ uint32_t index_find(const IndexView& ix, uint64_t h) {
  if (key_is_special(h)) {
    const uint64_t s = ix.C + (h == KEY_TOMB ? 1 : 0);
    return ix.cnt[s] ? (uint32_t)s : SLOT_MISS;
  }
  return index_resolve(ix, h, bucket_load(ix, h & ix.bmask));
}

}  // namespace fi
