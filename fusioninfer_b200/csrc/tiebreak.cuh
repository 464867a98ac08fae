// tiebreak.cuh — deterministic per-request rotation that breaks score ties (shared by the kernels and the
// host build of the arithmetic; the oracle restates it independently).
//
// Upstream MaxScorePicker shuffles the candidates before its stable sort, so equal
// totals are resolved at random and load spreads over the tied pods.  A fixed "lowest index" rule would send
// every request without a cached prefix to endpoint 0 under the reference's default prefix-only profile
// (/root/reference/pkg/router/strategy.go:51-68).  Here the order among tied endpoints is a ROTATION of the
// pool that starts at a position derived from the request itself:
//     seed  = n_blocks > 0 ? h_1 (first chained block hash) : h0 ^ (r + 1)·0x9E3779B97F4A7C15
//     start = (mix64(seed) >> 32) · E >> 32                      (E = num_endpoints, the WHOLE pool)
//     rot(e) = (e − start) mod E;   among equal totals the smallest rot wins.
// Requests that share their first block rotate alike (cold requests of one prefix land on one pod, which then
// caches it); distinct prefixes spread uniformly.  Every rank of a sharded pool derives the same rotation.
#pragma once
#include <stdint.h>

#include "xxh64.cuh"  // FI_HD

namespace fi {

FI_HD uint64_t tie_mix(uint64_t x) {  // SplitMix64 finaliser
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBULL;
  x ^= x >> 31;
  return x;
}
FI_HD uint64_t tie_seed(uint32_t n_blocks, uint64_t first_hash, uint64_t h0, uint32_t r) {
  return n_blocks ? first_hash : (h0 ^ ((uint64_t)(r + 1u) * 0x9E3779B97F4A7C15ULL));
}
FI_HD uint32_t tie_start(uint64_t seed, uint32_t E) { return (uint32_t)(((tie_mix(seed) >> 32) * (uint64_t)E) >> 32); }
FI_HD uint32_t tie_rot(uint32_t e_global, uint32_t start, uint32_t E) {
  return e_global >= start ? e_global - start : e_global + (E - start);
}


// Tie sets of LOCAL endpoints are bit words over one contiguous range of the pool, so their rotation order is
// "ascending from local position p, wrapping to 0" with p = start - ep_begin when the rotation starts strictly
// inside this rank's range, else 0.
FI_HD uint32_t tie_local_origin(uint32_t start, uint32_t ep_begin, uint32_t ep_count) {
  return (start > ep_begin && start - ep_begin < ep_count) ? start - ep_begin : 0u;
}
// smallest rotated distance ((position - p) mod 32·W) among the set bits of word wi, 0xFFFFFFFF if none
FI_HD uint32_t tie_word_min(uint32_t w, uint32_t wi, uint32_t p, uint32_t mask) {
  const uint32_t base = wi * 32u;
  uint32_t hi = w;  // members at or after p within this word
  if (p > base) hi = (p - base < 32u) ? (w & (0xFFFFFFFFu << (p - base))) : 0u;
  const uint32_t pick = hi ? hi : w;
  if (!pick) return 0xFFFFFFFFu;
  uint32_t bit = 0;
  while (!((pick >> bit) & 1u)) ++bit;
  return (base + bit - p) & mask;
}

}  // namespace fi
