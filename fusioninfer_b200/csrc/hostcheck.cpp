// hostcheck.cpp — host build of the arithmetic the kernels share with the CPU
// (xxh64.cuh, bitslice.cuh) and of the host LRU, exported for CPU unit tests
// (tests/test_host_logic.py).  Not part of libfi_epp.so and never used to serve a pick.
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "bitslice.cuh"
#include "lru.h"
#include "lru_batch.h"
#include "lru_plan.h"
#include "tiebreak.cuh"
#include "xxh64.cuh"

extern "C" {

uint64_t fihc_xxh64(const uint8_t* p, uint32_t len) { return fi::xxh64_bytes(p, len); }

// chain via the generic virtual-message path (any block size)
uint32_t fihc_chain_generic(const uint8_t* p, uint64_t len, uint64_t h0, uint32_t B, uint32_t M, uint64_t* out) {
  uint64_t nb = len / B;
  if (nb > M) nb = M;
  uint64_t h = h0;
  for (uint64_t i = 0; i < nb; ++i) {
    fi::ChainMsg m{p + i * B, B, h, true};
    h = fi::xxh64_msg(m);
    out[i] = h;
  }
  return (uint32_t)nb;
}

// chain via the split pre-state / chain_step path (B % 32 == 0), as the GPU fast path does
uint32_t fihc_chain_split(const uint8_t* p, uint64_t len, uint64_t h0, uint32_t B, uint32_t M, uint64_t* out) {
  if (B % 32) return 0;
  uint64_t nb = len / B;
  if (nb > M) nb = M;
  uint64_t h = h0;
  for (uint64_t i = 0; i < nb; ++i) {
    fi::XAcc a = fi::xacc_init();
    for (uint32_t s = 0; s < B / 32; ++s) {
      uint64_t w[4];
      std::memcpy(w, p + i * B + 32 * s, 32);
      fi::xacc_stripe(a, w[0], w[1], w[2], w[3]);
    }
    const uint64_t pre = fi::xacc_finish(a, (uint64_t)B + 8);
    h = fi::chain_step(pre, h);
    out[i] = h;
  }
  return (uint32_t)nb;
}

// bit-sliced counting: add n words K at a time (zero padded), return the 32 counts
void fihc_bitcount(const uint32_t* words, uint32_t n, uint32_t K, uint32_t* counts) {
  fi::BitCounter b;
  fi::bc_clear(b);
  for (uint32_t i = 0; i < n; i += K) {
    uint32_t w[16] = {0};
    for (uint32_t j = 0; j < K && i + j < n; ++j) w[j] = words[i + j];
    switch (K) {
      case 1: { uint32_t t[1] = {w[0]}; fi::bc_add<1>(b, t); break; }
      case 2: { uint32_t t[2] = {w[0], w[1]}; fi::bc_add<2>(b, t); break; }
      case 4: { uint32_t t[4]; std::memcpy(t, w, sizeof(t)); fi::bc_add<4>(b, t); break; }
      case 8: { uint32_t t[8]; std::memcpy(t, w, sizeof(t)); fi::bc_add<8>(b, t); break; }
      default: { uint32_t t[16]; std::memcpy(t, w, sizeof(t)); fi::bc_add<16>(b, t); break; }
    }
  }
  for (uint32_t bit = 0; bit < 32; ++bit) counts[bit] = fi::bc_get(b, bit);
}

// merge of two counters built from two word streams
void fihc_bitcount_merge(const uint32_t* wa, uint32_t na, const uint32_t* wb, uint32_t nb, uint32_t* counts,
                         uint32_t* nonzero) {
  fi::BitCounter a, b;
  fi::bc_clear(a);
  fi::bc_clear(b);
  for (uint32_t i = 0; i < na; ++i) { uint32_t t[1] = {wa[i]}; fi::bc_add<1>(a, t); }
  for (uint32_t i = 0; i < nb; ++i) { uint32_t t[1] = {wb[i]}; fi::bc_add<1>(b, t); }
  fi::bc_merge(a, b);
  for (uint32_t bit = 0; bit < 32; ++bit) counts[bit] = fi::bc_get(a, bit);
  *nonzero = fi::bc_nonzero(a);
}

// tie rotation (tiebreak.cuh): start of a request's rotation, rotated distance of an endpoint, and the first
// member of a local tie set (bit words) in rotation order — the per-word arithmetic the match kernel uses
uint32_t fihc_tie_start(uint32_t n_blocks, uint64_t first_hash, uint64_t h0, uint32_t r, uint32_t E) {
  return fi::tie_start(fi::tie_seed(n_blocks, first_hash, h0, r), E);
}
uint32_t fihc_tie_rot(uint32_t e, uint32_t start, uint32_t E) { return fi::tie_rot(e, start, E); }
uint32_t fihc_tie_first_local(const uint32_t* words, uint32_t W, uint32_t start, uint32_t ep_begin, uint32_t ep_count) {
  const uint32_t p = fi::tie_local_origin(start, ep_begin, ep_count);
  const uint32_t mask = W * 32u - 1u;
  uint32_t best = 0xFFFFFFFFu;
  for (uint32_t wi = 0; wi < W; ++wi) {
    const uint32_t d = fi::tie_word_min(words[wi], wi, p, mask);
    if (d < best) best = d;
  }
  return best == 0xFFFFFFFFu ? 0xFFFFFFFFu : ((best + p) & mask);
}

// LRU trace: for each key report (inserted, did_evict, evicted)
void* fihc_lru_new(uint32_t cap) { return new fi::LruSet(cap); }
void fihc_lru_free(void* l) { delete (fi::LruSet*)l; }
uint32_t fihc_lru_size(void* l) { return ((fi::LruSet*)l)->size(); }
int fihc_lru_contains(void* l, uint64_t k) { return ((fi::LruSet*)l)->contains(k) ? 1 : 0; }
void fihc_lru_touch(void* l, const uint64_t* keys, uint32_t n, uint8_t* inserted, uint8_t* did_evict, uint64_t* evicted) {
  fi::LruSet* s = (fi::LruSet*)l;
  for (uint32_t i = 0; i < n; ++i) {
    bool d = false;
    uint64_t ev = 0;
    inserted[i] = s->touch(keys[i], &ev, &d) ? 1 : 0;
    did_evict[i] = d ? 1 : 0;
    evicted[i] = ev;
  }
}

// fi_epp_index_add_chains' host phase (lru_batch.h) against the sequential definition: walk the batch on
// `workers` threads, apply the resulting ops segment by segment (all SETs of a segment, then all its CLEARs — the
// way the GPU applies a group) to a membership set, and compare that set and the LRU contents with one LRU per
// endpoint touched request after request.  Returns 0 if identical, else a code; *segments = segments used.
int fihc_lru_batch_check(uint32_t E, uint32_t cap, const uint32_t* endpoints, const uint64_t* chains, uint32_t pitch,
                         const uint32_t* nblocks, uint32_t R, uint32_t batches, uint32_t workers, uint32_t* segments) {
  std::vector<fi::LruSet> par(E, fi::LruSet(cap)), seq(E, fi::LruSet(cap));
  fi::WorkerPool pool(workers);
  std::vector<fi::WorkerOps> outs;
  std::vector<std::vector<uint64_t>> member_par(E), member_seq(E);  // sorted membership per endpoint
  auto has = [](std::vector<uint64_t>& v, uint64_t k) { return std::binary_search(v.begin(), v.end(), k); };
  auto add = [&](std::vector<uint64_t>& v, uint64_t k) {
    auto it = std::lower_bound(v.begin(), v.end(), k);
    if (it == v.end() || *it != k) v.insert(it, k);
  };
  auto del = [&](std::vector<uint64_t>& v, uint64_t k) {
    auto it = std::lower_bound(v.begin(), v.end(), k);
    if (it != v.end() && *it == k) v.erase(it);
  };
  (void)has;
  uint32_t max_seg = 0;
  for (uint32_t b = 0; b < batches; ++b) {
    const uint32_t* ep = endpoints + (size_t)b * R;
    const uint64_t* ch = chains + (size_t)b * R * pitch;
    const uint32_t* nb = nblocks + (size_t)b * R;
    const size_t nseg = fi::lru_walk_batch(par, 0, E, ep, ch, pitch, nb, R, pool, outs);
    if (nseg > max_seg) max_seg = (uint32_t)nseg;
    for (size_t sgi = 0; sgi < nseg; ++sgi) {
      for (auto& o : outs)
        if (sgi < o.nseg)
          for (auto& op : o.sets[sgi]) add(member_par[op.endpoint], op.hash);
      for (auto& o : outs)
        if (sgi < o.nseg)
          for (auto& op : o.clears[sgi]) del(member_par[op.endpoint], op.hash);
    }
    for (uint32_t r = 0; r < R; ++r) {
      if (ep[r] == FI_NO_ENDPOINT || ep[r] >= E) continue;
      for (uint32_t i = 0; i < nb[r]; ++i) {
        uint64_t ev = 0;
        bool did = false;
        const uint64_t k = ch[(size_t)r * pitch + i];
        const bool ins = seq[ep[r]].touch(k, &ev, &did);
        if (ins) add(member_seq[ep[r]], k);
        if (did) del(member_seq[ep[r]], ev);
      }
    }
    for (uint32_t e = 0; e < E; ++e) {
      if (member_par[e] != member_seq[e]) return 1;
      if (par[e].size() != seq[e].size() || par[e].size() != member_seq[e].size()) return 2;
      for (uint64_t k : member_seq[e])
        if (!par[e].contains(k)) return 3;
    }
  }
  if (segments) *segments = max_seg;
  return 0;
}

// The device LRU's batch rule (lru_kernels.cu) on the CPU: plan the batch with lru_plan_batch (at most plan_cap
// touches per endpoint and sub-batch: the LRU capacity in the conservative pass, unlimited in the optimistic one),
// apply every sub-batch AT ONCE — per endpoint: the keys touched move behind everything else in the order of
// their LAST touch, then the oldest entries beyond the capacity go — and compare recency order and content with
// one LRU per endpoint touched request after request.  0 if identical; *subs = sub-batches of the last batch.
int fihc_lru_plan_check(uint32_t E, uint32_t cap, const uint32_t* endpoints, const uint64_t* chains, uint32_t pitch,
                        const uint32_t* nblocks, uint32_t R, uint32_t batches, uint32_t plan_cap, uint64_t cap_touches,
                        uint32_t cap_requests, uint32_t* subs) {
  std::vector<std::vector<uint64_t>> order(E);  // oldest first
  std::vector<fi::LruSet> seq(E, fi::LruSet(cap));
  std::vector<std::vector<uint64_t>> seq_order(E);
  fi::LruPlan pl;
  for (uint32_t b = 0; b < batches; ++b) {
    const uint32_t* ep = endpoints + (size_t)b * R;
    const uint64_t* ch = chains + (size_t)b * R * pitch;
    const uint32_t* nb = nblocks + (size_t)b * R;
    fi::lru_plan_batch(ep, nb, R, 0, E, plan_cap, cap_touches, cap_requests, &pl);
    if (subs) *subs = (uint32_t)pl.subs.size();
    size_t covered = 0;
    for (size_t sb = 0; sb < pl.subs.size(); ++sb) {
      const fi::LruSubBatch& s = pl.subs[sb];
      if (s.k_end - s.k_begin > cap_requests || s.touches > cap_touches) return 10;
      const uint32_t* st = pl.ep_start.data() + sb * ((size_t)E + 1);
      const uint32_t* inc = pl.inc.data() + sb * (size_t)E;
      for (uint32_t e = 0; e < E; ++e) {
        // touches of endpoint e in this sub-batch, in request order
        std::vector<uint64_t> t;
        uint32_t last_k = 0;
        for (uint32_t i = st[e]; i < st[e + 1]; ++i) {
          const uint32_t k = s.k_begin + pl.ep_list[s.k_begin + i];
          if (i > st[e] && k <= last_k) return 11;  // ascending
          last_k = k;
          if (pl.req_ep[k] != e) return 12;
          const uint64_t* c = ch + (size_t)pl.req_id[k] * pitch;
          t.insert(t.end(), c, c + pl.req_n[k]);
        }
        if (t.size() != inc[e] || t.size() > plan_cap) return 13;
        covered += t.size();
        std::unordered_set<uint64_t> seen;
        std::vector<uint64_t> winners;  // keys by last touch, newest first
        for (size_t i = t.size(); i-- > 0;)
          if (seen.insert(t[i]).second) winners.push_back(t[i]);
        std::vector<uint64_t>& o = order[e];
        o.erase(std::remove_if(o.begin(), o.end(), [&](uint64_t k) { return seen.count(k) != 0; }), o.end());
        o.insert(o.end(), winners.rbegin(), winners.rend());
        if (o.size() > cap) o.erase(o.begin(), o.begin() + (o.size() - cap));
      }
    }
    size_t want_cov = 0;
    for (uint32_t r = 0; r < R; ++r) {
      if (ep[r] >= E) continue;
      want_cov += nb[r];
      for (uint32_t i = 0; i < nb[r]; ++i) {
        const uint64_t k = ch[(size_t)r * pitch + i];
        uint64_t ev = 0;
        bool did = false;
        const bool ins = seq[ep[r]].touch(k, &ev, &did);
        std::vector<uint64_t>& so = seq_order[ep[r]];
        if (!ins) so.erase(std::find(so.begin(), so.end(), k));
        so.push_back(k);
        if (did) so.erase(std::find(so.begin(), so.end(), ev));
      }
    }
    if (covered != want_cov) return 14;
    for (uint32_t e = 0; e < E; ++e)
      if (order[e] != seq_order[e]) return 1;
  }
  return 0;
}

}  // extern "C"
