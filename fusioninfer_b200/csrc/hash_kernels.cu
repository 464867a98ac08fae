// hash_kernels.cu — batched token-block hashing for sm_100a (integer, HBM-stream bound).
//
// Computes the chained block keys of SURVEY.md Appendix A.1 (upstream
// prefix.hashPrompt; block size / cap from /root/reference/pkg/router/
// strategy.go:57-58,147-148):  h_i = XXH64(0, block_i ‖ LE64(h_{i-1})).
//
//   hash_blocks     one thread per block: the two (block_bytes/32) stripes, merge
//                   and length add — everything that does not depend on h_{i-1}.
//                   128-bit loads, 64 B per thread, a warp covers 2 KiB contiguous.
//   chain_finalize  one thread per request walks the serial tail+avalanche link.
//   hash_generic    block sizes that are not a multiple of 32 (e.g. the reference's
//                   blockSize: 5): fully serial per request, byte loads.
#include "kernels.cuh"
#include "xxh64.cuh"

namespace fi {

namespace {

__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

template <int STRIPES>
__global__ void __launch_bounds__(256) hash_blocks_kernel(const uint8_t* __restrict__ prompts,
                                                          const uint64_t* __restrict__ offsets, uint32_t M,
                                                          uint32_t MP, uint64_t* __restrict__ pre,
                                                          uint32_t* __restrict__ nblocks) {
  constexpr uint32_t B = STRIPES * 32;
  const uint32_t r = blockIdx.x;
  const uint64_t off = offsets[r];
  const uint64_t len = offsets[r + 1] - off;
  const uint64_t nb64 = len / B;
  const uint32_t n = nb64 > M ? M : (uint32_t)nb64;
  if (threadIdx.x == 0) nblocks[r] = n;
  const uint8_t* base = prompts + off;
  uint64_t* out = pre + (uint64_t)r * MP;
  if ((reinterpret_cast<uintptr_t>(base) & 15) == 0) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint4* p = reinterpret_cast<const uint4*>(base + (uint64_t)i * B);
      uint4 q[2 * STRIPES];
#pragma unroll
      for (int s = 0; s < 2 * STRIPES; ++s) q[s] = __ldg(p + s);
      XAcc a = xacc_init();
#pragma unroll
      for (int s = 0; s < STRIPES; ++s)
        xacc_stripe(a, pack64(q[2 * s].x, q[2 * s].y), pack64(q[2 * s].z, q[2 * s].w),
                    pack64(q[2 * s + 1].x, q[2 * s + 1].y), pack64(q[2 * s + 1].z, q[2 * s + 1].w));
      out[i] = xacc_finish(a, (uint64_t)B + 8);
    }
  } else {
    // arbitrary byte alignment: aligned 64-bit windows + funnel shift
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uintptr_t addr = reinterpret_cast<uintptr_t>(base + (uint64_t)i * B);
      const uint64_t* wp = reinterpret_cast<const uint64_t*>(addr & ~(uintptr_t)7);
      const uint32_t sh = (uint32_t)(addr & 7) * 8;
      uint64_t w[4 * STRIPES + 1];
#pragma unroll
      for (int k = 0; k < 4 * STRIPES; ++k) w[k] = __ldg(wp + k);
      w[4 * STRIPES] = sh ? __ldg(wp + 4 * STRIPES) : 0;
      if (sh) {
#pragma unroll
        for (int k = 0; k < 4 * STRIPES; ++k) w[k] = (w[k] >> sh) | (w[k + 1] << (64 - sh));
      }
      XAcc a = xacc_init();
#pragma unroll
      for (int s = 0; s < STRIPES; ++s) xacc_stripe(a, w[4 * s], w[4 * s + 1], w[4 * s + 2], w[4 * s + 3]);
      out[i] = xacc_finish(a, (uint64_t)B + 8);
    }
  }
}

// any block size that is a multiple of 32 (runtime stripe count)
__global__ void __launch_bounds__(256) hash_blocks_any_kernel(const uint8_t* __restrict__ prompts,
                                                              const uint64_t* __restrict__ offsets, uint32_t B,
                                                              uint32_t M, uint32_t MP, uint64_t* __restrict__ pre,
                                                              uint32_t* __restrict__ nblocks) {
  const uint32_t r = blockIdx.x;
  const uint64_t off = offsets[r];
  const uint64_t len = offsets[r + 1] - off;
  const uint64_t nb64 = len / B;
  const uint32_t n = nb64 > M ? M : (uint32_t)nb64;
  if (threadIdx.x == 0) nblocks[r] = n;
  const uint8_t* base = prompts + off;
  uint64_t* out = pre + (uint64_t)r * MP;
  const uint32_t stripes = B / 32;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(base + (uint64_t)i * B);
    const uint64_t* wp = reinterpret_cast<const uint64_t*>(addr & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(addr & 7) * 8;
    XAcc a = xacc_init();
    uint64_t cur = __ldg(wp);
    for (uint32_t s = 0; s < stripes; ++s) {
      uint64_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (sh) {
          uint64_t nxt = __ldg(wp + 4 * s + k + 1);
          w[k] = (cur >> sh) | (nxt << (64 - sh));
          cur = nxt;
        } else {
          w[k] = __ldg(wp + 4 * s + k);
        }
      }
      xacc_stripe(a, w[0], w[1], w[2], w[3]);
    }
    out[i] = xacc_finish(a, (uint64_t)B + 8);
  }
}

// One thread per request: h_i = chain_step(pre_i, h_{i-1}); groups of 8 with the
// next group's pre-states prefetched.  Entries [n, MP) are zeroed.
__global__ void __launch_bounds__(32) chain_finalize_kernel(const uint64_t* __restrict__ pre,
                                                            const uint32_t* __restrict__ nblocks,
                                                            const uint64_t* __restrict__ h0, uint32_t R, uint32_t MP,
                                                            uint64_t* __restrict__ chain) {
  const uint32_t r = blockIdx.x * 32 + threadIdx.x;
  if (r >= R) return;
  const uint32_t n = nblocks[r];
  uint64_t h = h0[r];
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(pre + (uint64_t)r * MP);
  ulonglong2* c = reinterpret_cast<ulonglong2*>(chain + (uint64_t)r * MP);
  const uint32_t MP2 = MP / 2;                 // 16-byte units per row
  const uint32_t ng = (n + 7) / 8;             // groups of 8 hashes = 4 units
  // Prefetch distance 2 groups: a group's 8 serial links take ~900 cycles, so loads
  // issued two groups ahead have landed whatever point of the loop body the
  // compiler schedules them at (with distance 1 it sinks them to the loop end and
  // the L2 latency of 32 uncoalesced lines per warp is exposed every group).
  ulonglong2 cur[4], nxt[4], nn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    cur[k] = make_ulonglong2(0, 0);
    nxt[k] = make_ulonglong2(0, 0);
    nn[k] = make_ulonglong2(0, 0);
    if (ng > 0 && (uint32_t)k < MP2) cur[k] = p[k];
    if (ng > 1 && 4 + (uint32_t)k < MP2) nxt[k] = p[4 + k];
  }
  for (uint32_t g = 0; g < ng; ++g) {
    if (g + 2 < ng) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if ((g + 2) * 4 + k < MP2) nn[k] = p[(g + 2) * 4 + k];
    }
    const uint32_t i0 = g * 8;
    if (i0 + 8 <= n) {
      // full group: nothing but the serial links on the dependency chain
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ulonglong2 o;
        h = chain_step(cur[k].x, h);
        o.x = h;
        h = chain_step(cur[k].y, h);
        o.y = h;
        c[g * 4 + k] = o;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ulonglong2 o;
        uint64_t t = chain_step(cur[k].x, h);
        const bool v0 = i0 + 2 * k < n;
        h = v0 ? t : h;
        o.x = v0 ? t : 0;
        t = chain_step(cur[k].y, h);
        const bool v1 = i0 + 2 * k + 1 < n;
        h = v1 ? t : h;
        o.y = v1 ? t : 0;
        if (g * 4 + k < MP2) c[g * 4 + k] = o;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cur[k] = nxt[k];
      nxt[k] = nn[k];
    }
  }
  for (uint32_t u = ng * 4; u < MP2; ++u) c[u] = make_ulonglong2(0, 0);
}

// Fully serial path for block sizes that are not a multiple of 32.
__global__ void __launch_bounds__(32) hash_generic_kernel(const uint8_t* __restrict__ prompts,
                                                          const uint64_t* __restrict__ offsets,
                                                          const uint64_t* __restrict__ h0, uint32_t R, uint32_t B,
                                                          uint32_t M, uint32_t MP, uint64_t* __restrict__ chain,
                                                          uint32_t* __restrict__ nblocks) {
  const uint32_t r = blockIdx.x * 32 + threadIdx.x;
  if (r >= R) return;
  const uint64_t off = offsets[r];
  const uint64_t len = offsets[r + 1] - off;
  const uint64_t nb64 = len / B;
  const uint32_t n = nb64 > M ? M : (uint32_t)nb64;
  nblocks[r] = n;
  uint64_t* out = chain + (uint64_t)r * MP;
  uint64_t h = h0[r];
  const uint8_t* base = prompts + off;
  for (uint32_t i = 0; i < n; ++i) {
    ChainMsg m{base + (uint64_t)i * B, B, h, true};
    h = xxh64_msg(m);
    out[i] = h;
  }
  for (uint32_t i = n; i < MP; ++i) out[i] = 0;
}

// -----------------------------------------------------------------------------------------
// Overlapped variant (block sizes 32/64/128): producers hash in CHUNK-MAJOR order — chunk j
// (32 blocks) of every request before chunk j+1 of any — and publish one counter per
// (32-request group, chunk); the walker kernel, launched right after on a second stream,
// walks a chunk's serial links as soon as that counter is complete.  The ~45 us serial
// chain walk then hides behind the HBM-bound hashing instead of following it.
// Deadlock-free by launch order: producers never wait; if the two kernels are ever
// serialised (profilers do that) the producers simply finish first.
// -----------------------------------------------------------------------------------------
constexpr int kProdWarps = 8;  // producer CTA: 8 requests x one 32-block chunk

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int STRIPES>
__global__ void __launch_bounds__(kProdWarps * 32) hash_chunks_kernel(const uint8_t* __restrict__ prompts,
                                                                      const uint64_t* __restrict__ offsets, uint32_t R,
                                                                      uint32_t M, uint32_t MP, uint32_t nch,
                                                                      uint64_t* __restrict__ pre,
                                                                      uint32_t* __restrict__ nblocks,
                                                                      uint32_t* __restrict__ ready) {
  constexpr uint32_t B = STRIPES * 32;
  const uint32_t n_rg = (R + kProdWarps - 1) / kProdWarps;
  const uint32_t j = blockIdx.x / n_rg;   // chunk (slow index: chunk-major order)
  const uint32_t rg = blockIdx.x % n_rg;  // group of kProdWarps requests
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = rg * kProdWarps + warp;
  if (r < R) {
    const uint64_t off = offsets[r];
    const uint64_t len = offsets[r + 1] - off;
    const uint64_t nb64 = len / B;
    const uint32_t n = nb64 > M ? M : (uint32_t)nb64;
    if (j == 0 && lane == 0) nblocks[r] = n;
    const uint32_t i = j * 32 + lane;
    if (i < n) {
      const uint8_t* blk = prompts + off + (uint64_t)i * B;
      XAcc a = xacc_init();
      if ((reinterpret_cast<uintptr_t>(blk) & 15) == 0) {
        const uint4* p = reinterpret_cast<const uint4*>(blk);
        uint4 q[2 * STRIPES];
#pragma unroll
        for (int s = 0; s < 2 * STRIPES; ++s) q[s] = __ldg(p + s);
#pragma unroll
        for (int s = 0; s < STRIPES; ++s)
          xacc_stripe(a, pack64(q[2 * s].x, q[2 * s].y), pack64(q[2 * s].z, q[2 * s].w),
                      pack64(q[2 * s + 1].x, q[2 * s + 1].y), pack64(q[2 * s + 1].z, q[2 * s + 1].w));
      } else {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(blk);
        const uint64_t* wp = reinterpret_cast<const uint64_t*>(addr & ~(uintptr_t)7);
        const uint32_t sh = (uint32_t)(addr & 7) * 8;
        uint64_t w[4 * STRIPES + 1];
#pragma unroll
        for (int k = 0; k < 4 * STRIPES; ++k) w[k] = __ldg(wp + k);
        w[4 * STRIPES] = sh ? __ldg(wp + 4 * STRIPES) : 0;
        if (sh) {
#pragma unroll
          for (int k = 0; k < 4 * STRIPES; ++k) w[k] = (w[k] >> sh) | (w[k + 1] << (64 - sh));
        }
#pragma unroll
        for (int s = 0; s < STRIPES; ++s) xacc_stripe(a, w[4 * s], w[4 * s + 1], w[4 * s + 2], w[4 * s + 3]);
      }
      __stcg(pre + (uint64_t)r * MP + i, xacc_finish(a, (uint64_t)B + 8));
    }
  }
  // publish: the CTA barrier orders every thread's stores before thread 0, whose gpu-scope
  // release (cumulative) orders them before the counter increment for the acquiring walker
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* flag = ready + (uint64_t)(rg * kProdWarps / 32) * nch + j;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(flag), "r"(1u) : "memory");
  }
}

// One thread per request, one warp per 32-request group; waits for each chunk's counter.
__global__ void __launch_bounds__(32) chain_walk_kernel(const uint64_t* __restrict__ pre,
                                                        const uint64_t* __restrict__ offsets,
                                                        const uint64_t* __restrict__ h0, uint32_t R, uint32_t B,
                                                        uint32_t M, uint32_t MP, uint32_t nch,
                                                        const uint32_t* __restrict__ ready, uint64_t* __restrict__ chain) {
  const uint32_t grp = blockIdx.x;
  const uint32_t r = grp * 32 + threadIdx.x;
  const bool valid = r < R;
  uint32_t n = 0;
  uint64_t h = 0;
  if (valid) {
    const uint64_t len = offsets[r + 1] - offsets[r];
    const uint64_t nb64 = len / B;
    n = nb64 > M ? M : (uint32_t)nb64;
    h = h0[r];
  }
  // producers signalling this group: ceil(requests in the group / kProdWarps)
  const uint32_t in_group = min(32u, R - grp * 32);
  const uint32_t target = (in_group + kProdWarps - 1) / kProdWarps;
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(pre + (uint64_t)(valid ? r : 0) * MP);
  ulonglong2* c = reinterpret_cast<ulonglong2*>(chain + (uint64_t)(valid ? r : 0) * MP);
  const uint32_t MP2 = MP / 2;
  for (uint32_t j = 0; j < nch; ++j) {
    if (!__any_sync(0xFFFFFFFFu, j * 32 < n)) break;
    if (threadIdx.x == 0) {
      const uint32_t* flag = ready + (uint64_t)grp * nch + j;
      while (ld_acquire_u32(flag) < target) __nanosleep(100);
    }
    __syncwarp();
    if (j * 32 >= n) continue;
    const uint32_t u0 = j * 16;  // first 16-byte unit of the chunk
    // groups of 8 links (4 units); the next group's pre-states are loaded (L2, not L1: other
    // SMs wrote them during this kernel) while this group's serial links run
    ulonglong2 cur[4], nxt[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cur[k] = make_ulonglong2(0, 0);
      nxt[k] = make_ulonglong2(0, 0);
      if (u0 + k < MP2) cur[k] = __ldcg(p + u0 + k);
    }
#pragma unroll 1
    for (uint32_t g = 0; g < 4; ++g) {
      const uint32_t i0 = j * 32 + g * 8;
      if (i0 >= n) break;
      const uint32_t ub = u0 + g * 4;
      if (g + 1 < 4 && i0 + 8 < n) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (ub + 4 + k < MP2) nxt[k] = __ldcg(p + ub + 4 + k);
      }
      if (i0 + 8 <= n) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ulonglong2 o;
          h = chain_step(cur[k].x, h);
          o.x = h;
          h = chain_step(cur[k].y, h);
          o.y = h;
          c[ub + k] = o;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ulonglong2 o;
          uint64_t t = chain_step(cur[k].x, h);
          const bool v0 = i0 + 2 * k < n;
          h = v0 ? t : h;
          o.x = v0 ? t : 0;
          t = chain_step(cur[k].y, h);
          const bool v1 = i0 + 2 * k + 1 < n;
          h = v1 ? t : h;
          o.y = v1 ? t : 0;
          if (ub + k < MP2) c[ub + k] = o;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
    }
  }
  if (valid)
    for (uint32_t u = ((n + 7) / 8) * 4; u < MP2; ++u) c[u] = make_ulonglong2(0, 0);
}

}  // namespace

cudaError_t launch_hash_blocks(const uint8_t* prompts, const uint64_t* offsets, uint32_t R, uint32_t B, uint32_t M,
                               uint32_t MP, uint64_t* pre, uint32_t* nblocks, cudaStream_t s) {
  if (R == 0) return cudaSuccess;
  uint32_t threads = (M + 31) / 32 * 32;
  if (threads > 256) threads = 256;
  if (B == 64)
    hash_blocks_kernel<2><<<R, threads, 0, s>>>(prompts, offsets, M, MP, pre, nblocks);
  else if (B == 32)
    hash_blocks_kernel<1><<<R, threads, 0, s>>>(prompts, offsets, M, MP, pre, nblocks);
  else if (B == 128)
    hash_blocks_kernel<4><<<R, threads, 0, s>>>(prompts, offsets, M, MP, pre, nblocks);
  else
    hash_blocks_any_kernel<<<R, threads, 0, s>>>(prompts, offsets, B, M, MP, pre, nblocks);
  return cudaGetLastError();
}

cudaError_t launch_chain_finalize(const uint64_t* pre, const uint32_t* nblocks, const uint64_t* h0, uint32_t R,
                                  uint32_t MP, uint64_t* chain, cudaStream_t s) {
  if (R == 0) return cudaSuccess;
  chain_finalize_kernel<<<(R + 31) / 32, 32, 0, s>>>(pre, nblocks, h0, R, MP, chain);
  return cudaGetLastError();
}

cudaError_t launch_hash_generic(const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0, uint32_t R,
                                uint32_t B, uint32_t M, uint32_t MP, uint64_t* chain, uint32_t* nblocks,
                                cudaStream_t s) {
  if (R == 0) return cudaSuccess;
  hash_generic_kernel<<<(R + 31) / 32, 32, 0, s>>>(prompts, offsets, h0, R, B, M, MP, chain, nblocks);
  return cudaGetLastError();
}


// ---- overlapped hashing (see hash_chunks_kernel) -----------------------------------------
bool hash_overlap_supported(uint32_t B) { return B == 32 || B == 64 || B == 128; }

uint32_t hash_overlap_flag_words(uint32_t R, uint32_t M) { return ((R + 31) / 32) * ((M + 31) / 32); }

cudaError_t launch_hash_chunks(const uint8_t* prompts, const uint64_t* offsets, uint32_t R, uint32_t B, uint32_t M,
                               uint32_t MP, uint64_t* pre, uint32_t* nblocks, uint32_t* ready, cudaStream_t s) {
  if (R == 0) return cudaSuccess;
  const uint32_t nch = (M + 31) / 32;
  const uint32_t grid = nch * ((R + kProdWarps - 1) / kProdWarps);
  if (B == 64)
    hash_chunks_kernel<2><<<grid, kProdWarps * 32, 0, s>>>(prompts, offsets, R, M, MP, nch, pre, nblocks, ready);
  else if (B == 32)
    hash_chunks_kernel<1><<<grid, kProdWarps * 32, 0, s>>>(prompts, offsets, R, M, MP, nch, pre, nblocks, ready);
  else if (B == 128)
    hash_chunks_kernel<4><<<grid, kProdWarps * 32, 0, s>>>(prompts, offsets, R, M, MP, nch, pre, nblocks, ready);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_chain_walk(const uint64_t* pre, const uint64_t* offsets, const uint64_t* h0, uint32_t R, uint32_t B,
                              uint32_t M, uint32_t MP, const uint32_t* ready, uint64_t* chain, cudaStream_t s) {
  if (R == 0) return cudaSuccess;
  const uint32_t nch = (M + 31) / 32;
  chain_walk_kernel<<<(R + 31) / 32, 32, 0, s>>>(pre, offsets, h0, R, B, M, MP, nch, ready, chain);
  return cudaGetLastError();
}

}  // namespace fi
