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
#include "xxh64_sm100.cuh"

namespace fi {

namespace {

__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// Prompt bytes are read exactly once: stream them through L2 with an evict-first policy so they
// do not push out the index rows/keys of the popular prefixes, which the match kernel re-reads
// every batch (its latency is L2-hit-rate bound; the hot set is ~80 MB of the 126 MB L2).
__device__ __forceinline__ uint64_t make_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ld_stream_v4(const uint4* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;\n"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}

// Pre-state layout: tiled request-minor within groups of 32 requests, in 16-byte units
// (two consecutive blocks):  unit u of request r lives at  ((r/32)*MP2 + u)*32 + r%32.
// The chain walker (one lane per request) then reads 512 contiguous bytes per warp load — 4 L1TEX
// wavefronts instead of the 32 of a row-major layout, which saturated the wavefront rate
// (measured: 17.6 of the kernel's 45 us were those loads).
__device__ __forceinline__ uint64_t pre_index(uint32_t r, uint32_t i, uint32_t MP2) {
  return ((((uint64_t)(r >> 5) * MP2 + (i >> 1)) * 32 + (r & 31)) << 1) + (i & 1);
}

// one 32-byte stripe per load: every lane fetches whole 32-byte sectors exactly once (with 16-byte loads the
// two halves of a sector were requested by two instructions, and with L1::no_allocate both went to L2: ncu
// counted 2x the prompt bytes between L2 and L1)
struct Stripe {
  uint32_t w[8];
};
__device__ __forceinline__ Stripe ld_stream_v8(const void* p, uint64_t pol) {
  Stripe v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;\n"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void xacc2_stripe(XAcc2& a, const Stripe& q) {
  a.v1 = xround2(a.v1, U2{q.w[0], q.w[1]});
  a.v2 = xround2(a.v2, U2{q.w[2], q.w[3]});
  a.v3 = xround2(a.v3, U2{q.w[4], q.w[5]});
  a.v4 = xround2(a.v4, U2{q.w[6], q.w[7]});
}

template <int STRIPES>
__global__ void __launch_bounds__(256, STRIPES <= 2 ? 8 : 5) hash_blocks_kernel(const uint8_t* __restrict__ prompts,
                                                          const uint64_t* __restrict__ offsets, uint32_t R, uint32_t M,
                                                          uint32_t MP, uint64_t* __restrict__ pre,
                                                          uint32_t* __restrict__ nblocks, uint32_t* __restrict__ zero_word) {
  constexpr uint32_t B = STRIPES * 32;
  const uint32_t MP2 = MP / 2;
  const uint64_t pol = make_evict_first_policy();
  if (zero_word && blockIdx.x == 0 && threadIdx.x == 0) *zero_word = 0;
  // one request per CTA pass; the grid is R CTAs, or capped when the kernel has to share the SMs with the
  // previous batch's match_pick (pipelined API)
  for (uint32_t r = blockIdx.x; r < R; r += gridDim.x) {
    const uint64_t off = offsets[r];
    const uint64_t len = offsets[r + 1] - off;
    const uint64_t nb64 = len / B;
    const uint32_t n = nb64 > M ? M : (uint32_t)nb64;
    if (threadIdx.x == 0) nblocks[r] = n;
    const uint8_t* base = prompts + off;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(base) & 31);
    if (mis == 0) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint8_t* p = base + (uint64_t)i * B;
        Stripe q[STRIPES];
#pragma unroll
        for (int s = 0; s < STRIPES; ++s) q[s] = ld_stream_v8(p + 32 * s, pol);
        XAcc2 a = xacc2_init();
#pragma unroll
        for (int s = 0; s < STRIPES; ++s) xacc2_stripe(a, q[s]);
        pre[pre_index(r, i, MP2)] = xacc2_finish(a, (uint64_t)B + 8);
      }
    } else if ((mis & 15) == 0) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(base + (uint64_t)i * B);
        uint4 q[2 * STRIPES];
#pragma unroll
        for (int s = 0; s < 2 * STRIPES; ++s) q[s] = ld_stream_v4(p + s, pol);
        XAcc2 a = xacc2_init();
#pragma unroll
        for (int s = 0; s < STRIPES; ++s)
          xacc2_stripe(a, Stripe{{q[2 * s].x, q[2 * s].y, q[2 * s].z, q[2 * s].w, q[2 * s + 1].x, q[2 * s + 1].y,
                                  q[2 * s + 1].z, q[2 * s + 1].w}});
        pre[pre_index(r, i, MP2)] = xacc2_finish(a, (uint64_t)B + 8);
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
        XAcc2 a = xacc2_init();
#pragma unroll
        for (int s = 0; s < STRIPES; ++s)
          xacc2_stripe(a, Stripe{{(uint32_t)w[4 * s], (uint32_t)(w[4 * s] >> 32), (uint32_t)w[4 * s + 1],
                                  (uint32_t)(w[4 * s + 1] >> 32), (uint32_t)w[4 * s + 2], (uint32_t)(w[4 * s + 2] >> 32),
                                  (uint32_t)w[4 * s + 3], (uint32_t)(w[4 * s + 3] >> 32)}});
        pre[pre_index(r, i, MP2)] = xacc2_finish(a, (uint64_t)B + 8);
      }
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
  const uint32_t MP2 = MP / 2;
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
    pre[pre_index(r, i, MP2)] = xacc_finish(a, (uint64_t)B + 8);
  }
}

// One lane per request, one warp per group of 32 requests: h_i = chain_step(pre_i, h_{i-1}), in
// groups of 8 links.  The walk is a pure dependency chain — 40 integer instructions per link, 5 dependent
// 64-bit multiplies, ~110 cycles by ptxas' own stall counts — run by ONE warp per scheduler, in order: every
// other instruction in the loop and every scoreboard wait adds straight to the batch's critical path (ncu,
// round 1: 59 instructions and 222 cycles per link).  What the loop is built around:
//   * pre-states are prefetched kAhead groups ahead with cp.async into a shared-memory ring (register
//     prefetching does not work: ptxas puts every ring load on one counting scoreboard, so waiting for the
//     oldest also waits for the newest; cp.async commit/wait groups have the needed "all but the N newest"
//     semantics), and the ring is read into registers one group EARLY (volatile ld.shared at the top of the
//     iteration), so neither the wait nor the shared-memory latency sits between two links;
//   * each lane stores its 8 hashes straight to its chain row as four 16-byte stores (row-major [r][i]: what
//     the match kernel stages and chains_out returns) — but one group LATE, at the top of the next iteration,
//     from registers nothing else writes for a whole group.  Stored right after the links (first version of
//     this kernel: 197 cycles per link) the four scattered STG.128 shared one set of data registers, and each
//     had to wait for the previous one's operand read behind 32 L1 wavefronts;
//   * the loop is unrolled by two groups with the register roles swapped, so no buffer is ever copied;
//   * the buffers are padded to whole groups (MP % 8 == 0): no per-unit predicates.
// Entries [n, MP) of every row are zeroed.
// kRing ring slots, prefetch distance kRing - 1 groups.  Two shapes: RING = 4 with the whole register file (one warp
// per scheduler on 128 SMs: nothing else hides the pre-state loads), and RING = 3 capped at 64 registers / 24 KB so
// that ALL 128 CTAs fit on the 16 SMs of the pipelined path's walker partition (eight warps per scheduler hide each
// other's latency there).

__device__ __forceinline__ void cp_async16_cg(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void st_row16(ulonglong2* p, const ulonglong2& v) {
  asm volatile("st.global.v2.u64 [%0], {%1, %2};\n" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}
__device__ __forceinline__ ulonglong2 lds16(const ulonglong2* p) {
  ulonglong2 v;
  const unsigned sa = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];\n" : "=l"(v.x), "=l"(v.y) : "r"(sa) : "memory");
  return v;
}

// Four warps (128 requests) per CTA, one per SM sub-partition.
constexpr int kChainWarps = 4;

template <int kRing, int MINB>
__global__ void __launch_bounds__(kChainWarps * 32, MINB) chain_finalize_kernel(const uint64_t* __restrict__ pre,
                                                                          const uint32_t* __restrict__ nblocks,
                                                                          const uint64_t* __restrict__ h0, uint32_t R,
                                                                          uint32_t MP, uint64_t* __restrict__ chain) {
  constexpr int kAhead = kRing - 1;
  __shared__ __align__(16) ulonglong2 s_ring[kChainWarps][kRing][4][32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t grp = blockIdx.x * kChainWarps + warp;
  if (grp * 32 >= R) return;  // whole warp out of range (no block-wide barriers below)
  ulonglong2(*ring)[4][32] = s_ring[warp];
  const uint32_t r = grp * 32 + lane;
  const bool valid = r < R;
  const uint32_t n = valid ? nblocks[r] : 0;
  uint64_t h = valid ? h0[r] : 0;
  const uint32_t MP2 = MP / 2;  // 16-byte units per row; MP % 8 == 0: whole groups of 4 units
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(pre) + ((uint64_t)grp * MP2) * 32 + lane;  // unit u at p[u*32]
  ulonglong2* out = reinterpret_cast<ulonglong2*>(chain) + (uint64_t)(valid ? r : 0) * MP2;
  uint32_t ng_warp = (n + 7) / 8;        // groups that need arithmetic: warp-uniform maximum ...
  uint32_t ng_full = valid ? n / 8 : 0;  // ... and the groups in which every lane has 8 blocks: minimum
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ng_warp = max(ng_warp, __shfl_xor_sync(0xFFFFFFFFu, ng_warp, d));
    ng_full = min(ng_full, __shfl_xor_sync(0xFFFFFFFFu, ng_full, d));
  }

  auto issue = [&](uint32_t g) {  // one commit group per call, empty or not: keeps the count uniform
    if (g < ng_warp) {
      ulonglong2* dst = &ring[g % kRing][0][lane];
      const ulonglong2* src = p + (uint64_t)g * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) cp_async16_cg(dst + k * 32, src + k * 32);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kAhead; ++s) issue((uint32_t)s);

  ulonglong2 inA[4], inB[4], outA[4], outB[4];
  cp_async_wait<kAhead - 1>();  // group 0 has landed (a lane reads only its own copies: no barrier)
#pragma unroll
  for (int k = 0; k < 4; ++k) inA[k] = lds16(&ring[0][k][lane]);

  // one group: refill the ring, store the PREVIOUS group's hashes (if any), fetch the NEXT group's pre-states
  // into `nxt`, then the 8 links of `cur` into `res`
  auto group = [&](uint32_t g, const ulonglong2 (&cur)[4], ulonglong2 (&nxt)[4], const ulonglong2 (&prev)[4],
                   ulonglong2 (&res)[4], bool store_prev) {
    issue(g + kAhead);  // refills the slot whose registers were taken one group ago
    if (store_prev && valid) {
      ulonglong2* o = out + (g - 1) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) st_row16(o + k, prev[k]);
    }
    cp_async_wait<kAhead - 1>();  // group g + 1 has landed; its registers are needed only by the next group
#pragma unroll
    for (int k = 0; k < 4; ++k) nxt[k] = lds16(&ring[(g + 1) % kRing][k][lane]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      h = chain_step(cur[k].x, h);
      res[k].x = h;
      h = chain_step(cur[k].y, h);
      res[k].y = h;
    }
  };

  uint32_t g = 0;
  bool pending = false;  // outA / outB of group g - 1 not stored yet (which one: parity of g)
  // ---- groups in which every lane of the warp has all 8 blocks, two per iteration (A/B roles swap)
  if (ng_full >= 1) {
    group(0, inA, inB, outB, outA, false);
    g = 1;
    pending = true;
#pragma unroll 1
    for (; g + 1 < ng_full; g += 2) {
      group(g, inB, inA, outA, outB, true);
      group(g + 1, inA, inB, outB, outA, true);
    }
    if (g < ng_full) {  // one more (odd position): B in, A out
      group(g, inB, inA, outA, outB, true);
      ++g;
      // bring the roles back to "next input in inA, last output in outA"
#pragma unroll
      for (int k = 0; k < 4; ++k) outA[k] = outB[k];
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) inA[k] = inB[k];
    }
  }
  if (pending && valid) {
    ulonglong2* o = out + (g - 1) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) st_row16(o + k, outA[k]);
  }
  // ---- ragged groups: some lane's chain ends inside (next input is in inA)
#pragma unroll 1
  for (; g < ng_warp; ++g) {
    issue(g + kAhead);
    cp_async_wait<kAhead - 1>();
#pragma unroll
    for (int k = 0; k < 4; ++k) inB[k] = lds16(&ring[(g + 1) % kRing][k][lane]);
    ulonglong2* o = out + g * 4;
    const uint32_t i0 = g * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ulonglong2 v;
      uint64_t t = chain_step(inA[k].x, h);
      const bool v0 = i0 + 2 * k < n;
      h = v0 ? t : h;
      v.x = v0 ? t : 0;
      t = chain_step(inA[k].y, h);
      const bool v1 = i0 + 2 * k + 1 < n;
      h = v1 ? t : h;
      v.y = v1 ? t : 0;
      if (valid) st_row16(o + k, v);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) inA[k] = inB[k];
  }
  cp_async_wait<0>();
  // ---- the rest of every row is zero
  if (valid) {
    const ulonglong2 z = make_ulonglong2(0, 0);
    for (uint32_t u = g * 4; u < MP2; ++u) st_row16(out + u, z);
  }
}

// Fully serial path for block sizes that are not a multiple of 32.
__global__ void __launch_bounds__(128) hash_generic_kernel(const uint8_t* __restrict__ prompts,
                                                          const uint64_t* __restrict__ offsets,
                                                          const uint64_t* __restrict__ h0, uint32_t R, uint32_t B,
                                                          uint32_t M, uint32_t MP, uint64_t* __restrict__ chain,
                                                          uint32_t* __restrict__ nblocks) {
  const uint32_t r = blockIdx.x * 128 + threadIdx.x;  // 4 warps per CTA: one per SM sub-partition
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

}  // namespace

cudaError_t launch_hash_blocks(const uint8_t* prompts, const uint64_t* offsets, uint32_t R, uint32_t B, uint32_t M,
                               uint32_t MP, uint64_t* pre, uint32_t* nblocks, uint32_t grid_cap, cudaStream_t s,
                               uint32_t* zero_word) {
  if (R == 0) return cudaSuccess;
  uint32_t threads = (M + 31) / 32 * 32;
  if (threads > 256) threads = 256;
  const uint32_t grid = (grid_cap && grid_cap < R) ? grid_cap : R;
  if (B == 64)
    hash_blocks_kernel<2><<<grid, threads, 0, s>>>(prompts, offsets, R, M, MP, pre, nblocks, zero_word);
  else if (B == 32)
    hash_blocks_kernel<1><<<grid, threads, 0, s>>>(prompts, offsets, R, M, MP, pre, nblocks, zero_word);
  else if (B == 128)
    hash_blocks_kernel<4><<<grid, threads, 0, s>>>(prompts, offsets, R, M, MP, pre, nblocks, zero_word);
  else {
    hash_blocks_any_kernel<<<R, threads, 0, s>>>(prompts, offsets, B, M, MP, pre, nblocks);
    if (zero_word) cudaMemsetAsync(zero_word, 0, sizeof(uint32_t), s);
  }
  return cudaGetLastError();
}

cudaError_t launch_chain_finalize(const uint64_t* pre, const uint32_t* nblocks, const uint64_t* h0, uint32_t R,
                                  uint32_t MP, uint64_t* chain, bool compact, cudaStream_t s) {
  if (R == 0) return cudaSuccess;
  const uint32_t groups = (R + 31) / 32;
  const uint32_t grid = (groups + kChainWarps - 1) / kChainWarps;
  if (compact) chain_finalize_kernel<3, 8><<<grid, kChainWarps * 32, 0, s>>>(pre, nblocks, h0, R, MP, chain);
  else chain_finalize_kernel<4, 1><<<grid, kChainWarps * 32, 0, s>>>(pre, nblocks, h0, R, MP, chain);
  return cudaGetLastError();
}

cudaError_t launch_hash_generic(const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0, uint32_t R,
                                uint32_t B, uint32_t M, uint32_t MP, uint64_t* chain, uint32_t* nblocks,
                                cudaStream_t s) {
  if (R == 0) return cudaSuccess;
  hash_generic_kernel<<<(R + 127) / 128, 128, 0, s>>>(prompts, offsets, h0, R, B, M, MP, chain, nblocks);
  return cudaGetLastError();
}


}  // namespace fi
