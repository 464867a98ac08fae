// match_kernels.cu — batched longest-prefix-match + weighted score + warp-shuffle
// argmax over all endpoints, sm_100a.  HBM/L2-latency bound random row reads; no
// tensor cores (there is no dense contraction on this path).
//
// One warp per request (persistent, strided over requests):
//   1. stage the request's block-hash chain in shared memory (cp.async);
//   2. 32 blocks at a time: every lane probes one block hash in the key table
//      (one 32 B sector), the warp finds the first miss with a ballot
//      (upstream Plugin.matchLongestPrefix stops at the first block no pod holds —
//      SURVEY.md Appendix A.3), the next 32 probes are issued, then the rows of
//      the present blocks are read 16 at a time, fully coalesced (a row is the
//      bitset over the local endpoints; a lane owns one 32-endpoint word);
//   3. per-endpoint match counts accumulate in bit-planes (bitslice.cuh);
//   4. only endpoints with a non-zero count are scored individually, in fp64 with
//      explicit round-to-nearest mul/add in profile order (SURVEY.md Appendix A.4,
//      weights of /root/reference/pkg/router/strategy.go:66,157,163); all others
//      share the per-batch "zero-match best" precomputed by prepare_endpoints —
//      valid because every scorer weight is >= 0, so a total is monotone in the
//      match count;
//   5. warp-shuffle argmax, ties to the lowest endpoint index (Appendix A.5), then
//      the pd-profile-handler threshold rule (Appendix A.6,
//      /root/reference/pkg/router/strategy.go:129-133).
//
// Rows narrower than 32 words (fewer than 1024 local endpoints, e.g. an
// endpoint-range shard of a multi-GPU pool) are read G = 32/L rows per load
// instruction by G lane groups whose counters are merged at the end.
#include <climits>

#include "bitslice.cuh"
#include "index_device.cuh"
#include "kernels.cuh"

namespace fi {

namespace {

constexpr int kWarps = 8;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct Best {
  double score;
  uint32_t e;  // local endpoint or FI_NO_ENDPOINT
  uint32_t m;
};

__device__ __forceinline__ bool better(double s, uint32_t e, const Best& b) {
  return s > b.score || (s == b.score && e < b.e);
}

// SURVEY.md Appendix A.4 — identical operation order to oracle/epp_oracle.cpp:total_score
__device__ __forceinline__ double total_score(const ProfileDev& pr, const double* __restrict__ sc_p, uint32_t Epad,
                                              uint32_t e, uint32_t m, uint32_t n) {
  double total = 0.0;
#pragma unroll
  for (int s = 0; s < (int)FI_EPP_MAX_SCORERS; ++s) {
    if (s < (int)pr.n_scorers) {
      double v;
      if (pr.kind[s] == FI_SCORER_PREFIX)
        v = n ? __ddiv_rn((double)m, (double)n) : 0.0;
      else
        v = sc_p[(uint64_t)s * Epad + e];
      total = __dadd_rn(total, __dmul_rn(v, pr.weight[s]));
    }
  }
  return total;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// PD rule shared by the single-GPU kernel and the multi-GPU merge kernel
__device__ __forceinline__ bool pd_prefill_runs(uint32_t dec_endpoint, uint32_t dec_match, uint32_t n, uint64_t len,
                                                double threshold) {
  double hit = (dec_endpoint != FI_NO_ENDPOINT && n) ? __ddiv_rn((double)dec_match, (double)n) : 0.0;
  double miss_bytes = __dmul_rn(__dsub_rn(1.0, hit), (double)len);
  return miss_bytes >= threshold;
}

template <int L, int WPL, bool LPM, bool GMASK>
__global__ void __launch_bounds__(kWarps * 32) match_pick_kernel(const MatchParams p) {
  static_assert(L == 32 || WPL == 1, "multi-word lanes only for full-width rows");
  constexpr int G = 32 / L;                              // rows per load instruction
  constexpr int BATCH = (L >= 16 ? 16 : L) / WPL > 0 ? (L >= 16 ? 16 : L) / WPL : 1;  // load instrs in flight
  extern __shared__ __align__(16) uint64_t s_mem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t = lane % L;  // word within the row
  const int g = lane / L;  // row group
  uint64_t* s_chain = s_mem + (size_t)warp * p.MP;
  const IndexView ix = p.ix;
  const uint32_t P = p.st.n_profiles;

  for (uint32_t r = blockIdx.x * kWarps + warp; r < p.R; r += gridDim.x * kWarps) {
    const uint32_t n = p.nblocks[r];
    // ---- 1. stage the chain ---------------------------------------------------
    {
      const uint64_t* crow = p.chain + (uint64_t)r * p.MP;
      for (uint32_t u = lane; 2 * u < n; u += 32) cp_async16(s_chain + 2 * u, crow + 2 * u);
      cp_async_wait_all();
      __syncwarp();
    }
    // ---- global first miss from the ranks' presence masks (sharded upstream mode)
    uint32_t kg = n;
    if (GMASK) {
      uint32_t orv = 0;
      if ((uint32_t)lane < p.mask_words)
        for (uint32_t rk = 0; rk < p.gmask_ranks; ++rk)
          orv |= p.gmask[((uint64_t)rk * p.R + r) * p.mask_words + lane];
      uint32_t inv = ~orv;
      uint32_t pos = ((uint32_t)lane < p.mask_words && inv) ? lane * 32 + (__ffs(inv) - 1) : 0xFFFFFFFFu;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) pos = min(pos, __shfl_xor_sync(FULL, pos, d));
      kg = min(n, pos);
    }

    BitCounter cnt[WPL];
    uint32_t alive[WPL];
#pragma unroll
    for (int x = 0; x < WPL; ++x) {
      bc_clear(cnt[x]);
      alive[x] = 0xFFFFFFFFu;
    }
    uint32_t matched_rows = 0;
    bool real_miss = false;

    // ---- 2./3. probe + row reads, 32 blocks per chunk ---------------------------
    const uint32_t nchunks = (kg + 31) / 32;
    uint64_t h = 0;
    bool valid = (uint32_t)lane < kg;
    BucketRegs br;
    br.a = make_uint4(0, 0, 0, 0);
    br.b = br.a;
    if (valid) {
      h = s_chain[lane];
      if (!key_is_special(h)) br = bucket_load(ix, h & ix.bmask);
    }
    for (uint32_t c = 0; c < nchunks; ++c) {
      uint32_t slot = SLOT_MISS;
      if (valid) slot = key_is_special(h) ? index_find(ix, h) : index_resolve(ix, h, br);
      uint32_t rows_here;
      bool stop = false;
      if (GMASK) {
        rows_here = min(32u, kg - c * 32);
      } else {
        const unsigned mm = __ballot_sync(FULL, slot == SLOT_MISS);
        rows_here = mm ? (uint32_t)(__ffs(mm) - 1) : 32u;
        if (mm) {
          stop = true;
          real_miss = (c * 32 + rows_here) < n;
        }
      }
      // issue the next chunk's probes before touching this chunk's rows
      uint64_t hn = 0;
      bool validn = false;
      BucketRegs brn = br;
      if (!stop && c + 1 < nchunks) {
        const uint32_t idx = (c + 1) * 32 + lane;
        validn = idx < kg;
        if (validn) {
          hn = s_chain[idx];
          if (!key_is_special(hn)) brn = bucket_load(ix, hn & ix.bmask);
        }
      }
      // rows of this chunk
#pragma unroll 1
      for (int q0 = 0; q0 < L; q0 += BATCH) {
        if ((uint32_t)(q0 * G) >= rows_here) break;
        uint32_t w[WPL][BATCH];
#pragma unroll
        for (int qi = 0; qi < BATCH; ++qi) {
          const int j = (q0 + qi) * G + g;  // row of the chunk this lane helps read
          const uint32_t s = __shfl_sync(FULL, slot, j & 31);
          const bool ok = (uint32_t)j < rows_here && s != SLOT_MISS;
          const uint32_t* rp = ix.rows + ((uint64_t)s << ix.logW) + t;
#pragma unroll
          for (int x = 0; x < WPL; ++x) w[x][qi] = ok ? __ldg(rp + 32 * x) : 0u;
        }
        if (LPM) {
#pragma unroll
          for (int qi = 0; qi < BATCH; ++qi) {
#pragma unroll
            for (int x = 0; x < WPL; ++x) {
              uint32_t v = w[x][qi];
              if (G > 1) {  // prefix-AND over the G rows of this instruction
#pragma unroll
                for (int d = 1; d < G; d <<= 1) {
                  const uint32_t o = __shfl_up_sync(FULL, v, d * L);
                  if (g >= d) v &= o;
                }
                v &= alive[x];
                alive[x] = __shfl_sync(FULL, v, (G - 1) * L + t);
              } else {
                v &= alive[x];
                alive[x] = v;
              }
              w[x][qi] = v;
            }
          }
        }
#pragma unroll
        for (int x = 0; x < WPL; ++x) bc_add<BATCH>(cnt[x], w[x]);
      }
      matched_rows += rows_here;
      if (stop) break;
      if (LPM) {  // every local endpoint already dropped out: nothing more can match
        bool any = false;
#pragma unroll
        for (int x = 0; x < WPL; ++x) any |= alive[x] != 0;
        if (!__ballot_sync(FULL, any)) break;
      }
      h = hn;
      valid = validn;
      br = brn;
    }

    // ---- merge the lane groups' counters ---------------------------------------
    if (G > 1) {
#pragma unroll
      for (int d = L; d < 32; d <<= 1) {
        BitCounter o;
#pragma unroll
        for (int pl = 0; pl < NPLANES; ++pl) o.c[pl] = __shfl_xor_sync(FULL, cnt[0].c[pl], d);
        bc_merge(cnt[0], o);
      }
    }

    // ---- 4./5. score candidates, argmax, PD rule ---------------------------------
    uint32_t dec_e = FI_NO_ENDPOINT, dec_m = 0;
#pragma unroll
    for (int pi = 0; pi < (int)FI_EPP_MAX_PROFILES; ++pi) {
      if (pi < (int)P) {
        const ProfileDev& pr = p.st.prof[pi];
        const double* sc_p = p.st.sc + (uint64_t)pi * FI_EPP_MAX_SCORERS * p.st.Epad;
        Best b;
        b.score = -1.0;
        b.e = FI_NO_ENDPOINT;
        b.m = 0;
        if (g == 0) {
#pragma unroll
          for (int x = 0; x < WPL; ++x) {
            const uint32_t wi = t + 32 * x;
            uint32_t cand = bc_nonzero(cnt[x]) & p.st.elig[(uint64_t)pi * ix.W + wi];
            while (cand) {
              const uint32_t bit = __ffs(cand) - 1;
              cand &= cand - 1;
              const uint32_t e = wi * 32 + bit;
              const uint32_t m = bc_get(cnt[x], bit);
              const double s = total_score(pr, sc_p, p.st.Epad, e, m, n);
              if (better(s, e, b)) {
                b.score = s;
                b.e = e;
                b.m = m;
              }
            }
          }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const double os = __shfl_xor_sync(FULL, b.score, d);
          const uint32_t oe = __shfl_xor_sync(FULL, b.e, d);
          const uint32_t om = __shfl_xor_sync(FULL, b.m, d);
          if (better(os, oe, b)) {
            b.score = os;
            b.e = oe;
            b.m = om;
          }
        }
        const ZeroBest zb = p.st.zero[pi];
        if (zb.e_local != FI_NO_ENDPOINT && better(zb.score, zb.e_local, b)) {
          b.score = zb.score;
          b.e = zb.e_local;
          b.m = 0;
        }
        if (pi == (int)p.pd_decode) {
          dec_e = b.e;
          dec_m = b.m;
        }
        if (lane == 0) {
          fi_pick pk;
          const bool none = b.e == FI_NO_ENDPOINT;
          pk.endpoint = none ? FI_NO_ENDPOINT : b.e + p.ep_begin;
          pk.match_blocks = none ? 0 : (uint16_t)b.m;
          pk.n_blocks = (uint16_t)n;
          pk.score = none ? 0.0 : b.score;
          p.out[(uint64_t)r * P + pi] = pk;
        }
      }
    }
    if (lane == 0) {
      if (p.apply_pd) {  // pd-profile-handler: the prefill pick stands only if the threshold test passes
        const uint64_t len = p.offsets[r + 1] - p.offsets[r];
        if (!pd_prefill_runs(dec_e, dec_m, n, len, p.pd_threshold)) {
          fi_pick pk;
          pk.endpoint = FI_NO_ENDPOINT;
          pk.match_blocks = 0;
          pk.n_blocks = (uint16_t)n;
          pk.score = 0.0;
          p.out[(uint64_t)r * P + p.pd_prefill] = pk;
        }
      }
      if (p.probed_blocks) atomicAdd(p.probed_blocks, (unsigned long long)(matched_rows + (real_miss ? 1 : 0)));
    }
    __syncwarp();  // s_chain is rewritten by the next request
  }
}

// presence mask of every block of every request on this rank (sharded upstream mode)
__global__ void __launch_bounds__(kWarps * 32) probe_mask_kernel(const MatchParams p, uint32_t* __restrict__ mask_out) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (uint32_t r = blockIdx.x * kWarps + warp; r < p.R; r += gridDim.x * kWarps) {
    const uint32_t n = p.nblocks[r];
    const uint64_t* crow = p.chain + (uint64_t)r * p.MP;
    for (uint32_t c = 0; c < p.mask_words; ++c) {
      const uint32_t idx = c * 32 + lane;
      bool present = false;
      if (idx < n) present = index_find(p.ix, crow[idx]) != SLOT_MISS;
      const unsigned m = __ballot_sync(FULL, present);
      if (lane == 0) mask_out[(uint64_t)r * p.mask_words + c] = m;
    }
  }
}

// multi-GPU: reduce the ranks' local picks (score desc, endpoint asc), then the PD rule
__global__ void __launch_bounds__(256) merge_picks_kernel(const MergeParams p) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  fi_pick best[FI_EPP_MAX_PROFILES];
  for (uint32_t pi = 0; pi < p.P; ++pi) {
    fi_pick b;
    b.endpoint = FI_NO_ENDPOINT;
    b.match_blocks = 0;
    b.n_blocks = (uint16_t)p.nblocks[r];
    b.score = 0.0;
    for (uint32_t rk = 0; rk < p.ranks; ++rk) {
      const fi_pick c = p.gathered[((uint64_t)rk * p.R + r) * p.P + pi];
      if (c.endpoint == FI_NO_ENDPOINT) continue;
      if (b.endpoint == FI_NO_ENDPOINT || c.score > b.score || (c.score == b.score && c.endpoint < b.endpoint)) b = c;
    }
    best[pi] = b;
  }
  if (p.apply_pd) {
    const fi_pick d = best[p.pd_decode];
    const uint64_t len = p.offsets[r + 1] - p.offsets[r];
    if (!pd_prefill_runs(d.endpoint, d.match_blocks, p.nblocks[r], len, p.pd_threshold)) {
      best[p.pd_prefill].endpoint = FI_NO_ENDPOINT;
      best[p.pd_prefill].match_blocks = 0;
      best[p.pd_prefill].score = 0.0;
    }
  }
  for (uint32_t pi = 0; pi < p.P; ++pi) p.out[(uint64_t)r * p.P + pi] = best[pi];
}

// Per-batch constants of the non-prefix scorers: eligibility
// words, clamp01'd kv / queue scores of the local endpoints, and the best endpoint
// of each profile when nothing matches.  One CTA.
__global__ void __launch_bounds__(1024) prepare_endpoints_kernel(const EndpointDev* __restrict__ eps, uint32_t E_global,
                                                                 uint32_t ep_begin, uint32_t ep_count, ScoreTables st,
                                                                 double* __restrict__ sc, uint32_t* __restrict__ elig,
                                                                 ZeroBest* __restrict__ zero) {
  __shared__ int s_min[32], s_max[32];
  __shared__ double s_bs[32];
  __shared__ uint32_t s_be[32];
  __shared__ int s_minq, s_maxq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t W = st.Epad / 32;
  for (uint32_t pi = 0; pi < st.n_profiles; ++pi) {
    const ProfileDev pr = st.prof[pi];
    // queue min/max over the eligible endpoints of the WHOLE pool
    int mn = INT_MAX, mx = INT_MIN, any = 0;
    for (uint32_t e = tid; e < E_global; e += blockDim.x) {
      const EndpointDev s = eps[e];
      if ((s.flags & FI_ENDPOINT_ALIVE) && (pr.role_mask == 0 || (s.role_mask & pr.role_mask))) {
        mn = min(mn, s.queue_depth);
        mx = max(mx, s.queue_depth);
        any = 1;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn = min(mn, __shfl_xor_sync(FULL, mn, d));
      mx = max(mx, __shfl_xor_sync(FULL, mx, d));
      any |= __shfl_xor_sync(FULL, any, d);
    }
    if (lane == 0) {
      s_min[warp] = mn;
      s_max[warp] = mx;
    }
    __syncthreads();
    if (tid == 0) {
      int a = INT_MAX, b = INT_MIN;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        a = min(a, s_min[w]);
        b = max(b, s_max[w]);
      }
      s_minq = a;
      s_maxq = b;
    }
    __syncthreads();
    const int minq = s_minq, maxq = s_maxq;
    double* sc_p = sc + (uint64_t)pi * FI_EPP_MAX_SCORERS * st.Epad;
    Best b;
    b.score = -1.0;
    b.e = FI_NO_ENDPOINT;
    b.m = 0;
    for (uint32_t e = tid; e < st.Epad; e += blockDim.x) {  // blockDim multiple of 32, Epad multiple of 32
      bool ok = false;
      EndpointDev s;
      s.kv_util = 0.0;
      s.queue_depth = 0;
      s.role_mask = 0;
      s.flags = 0;
      if (e < ep_count) {
        s = eps[ep_begin + e];
        ok = (s.flags & FI_ENDPOINT_ALIVE) && (pr.role_mask == 0 || (s.role_mask & pr.role_mask));
      }
      const unsigned word = __ballot_sync(FULL, ok);
      if (lane == 0) elig[(uint64_t)pi * W + e / 32] = word;
      double tot = 0.0;
      for (uint32_t k = 0; k < pr.n_scorers; ++k) {
        double v = 0.0;
        if (pr.kind[k] == FI_SCORER_KV_UTIL) {
          v = __dsub_rn(1.0, s.kv_util);
        } else if (pr.kind[k] == FI_SCORER_QUEUE) {
          v = (maxq == minq) ? 1.0
                             : __ddiv_rn((double)((long long)maxq - (long long)s.queue_depth),
                                         (double)((long long)maxq - (long long)minq));
        }
        v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
        if (!ok) v = 0.0;
        sc_p[(uint64_t)k * st.Epad + e] = v;
        tot = __dadd_rn(tot, __dmul_rn(v, pr.weight[k]));  // prefix scorer contributes 0·w at zero match
      }
      if (ok && better(tot, e, b)) {
        b.score = tot;
        b.e = e;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double os = __shfl_xor_sync(FULL, b.score, d);
      const uint32_t oe = __shfl_xor_sync(FULL, b.e, d);
      if (better(os, oe, b)) {
        b.score = os;
        b.e = oe;
      }
    }
    if (lane == 0) {
      s_bs[warp] = b.score;
      s_be[warp] = b.e;
    }
    __syncthreads();
    if (tid == 0) {
      Best z;
      z.score = -1.0;
      z.e = FI_NO_ENDPOINT;
      z.m = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
        if (better(s_bs[w], s_be[w], z)) {
          z.score = s_bs[w];
          z.e = s_be[w];
        }
      ZeroBest zb;
      zb.score = z.e == FI_NO_ENDPOINT ? 0.0 : z.score;
      zb.e_local = z.e;
      zb.pad = 0;
      zero[pi] = zb;
    }
    __syncthreads();
  }
}

template <int L, int WPL>
cudaError_t launch_match_t(const MatchParams& p, int sm_count, cudaStream_t s) {
  const size_t smem = (size_t)kWarps * p.MP * sizeof(uint64_t);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (p.R + kWarps - 1) / kWarps;
    const uint32_t cap = (uint32_t)sm_count * (uint32_t)per_sm;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    kern<<<grid, kWarps * 32, smem, s>>>(p);
    return cudaGetLastError();
  };
  const bool lpm = p.lpm == FI_MATCH_LPM;
  const bool gm = p.gmask != nullptr;
  if (lpm) return gm ? go(match_pick_kernel<L, WPL, true, true>) : go(match_pick_kernel<L, WPL, true, false>);
  return gm ? go(match_pick_kernel<L, WPL, false, true>) : go(match_pick_kernel<L, WPL, false, false>);
}

}  // namespace

cudaError_t launch_match_pick(const MatchParams& p, int sm_count, cudaStream_t s) {
  if (p.R == 0) return cudaSuccess;
  switch (p.ix.W) {
    case 1: return launch_match_t<1, 1>(p, sm_count, s);
    case 2: return launch_match_t<2, 1>(p, sm_count, s);
    case 4: return launch_match_t<4, 1>(p, sm_count, s);
    case 8: return launch_match_t<8, 1>(p, sm_count, s);
    case 16: return launch_match_t<16, 1>(p, sm_count, s);
    case 32: return launch_match_t<32, 1>(p, sm_count, s);
    case 64: return launch_match_t<32, 2>(p, sm_count, s);
    case 128: return launch_match_t<32, 4>(p, sm_count, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_probe_mask(const MatchParams& p, uint32_t* mask_out, int sm_count, cudaStream_t s) {
  if (p.R == 0) return cudaSuccess;
  uint32_t grid = (p.R + kWarps - 1) / kWarps;
  const uint32_t cap = (uint32_t)sm_count * 8;
  if (grid > cap) grid = cap;
  probe_mask_kernel<<<grid, kWarps * 32, 0, s>>>(p, mask_out);
  return cudaGetLastError();
}

cudaError_t launch_merge_picks(const MergeParams& p, cudaStream_t s) {
  if (p.R == 0) return cudaSuccess;
  merge_picks_kernel<<<(p.R + 255) / 256, 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_prepare_endpoints(const EndpointDev* eps, uint32_t E_global, uint32_t ep_begin, uint32_t ep_count,
                                     ScoreTables st, double* sc, uint32_t* elig, ZeroBest* zero, cudaStream_t s) {
  prepare_endpoints_kernel<<<1, 1024, 0, s>>>(eps, E_global, ep_begin, ep_count, st, sc, elig, zero);
  return cudaGetLastError();
}

}  // namespace fi
