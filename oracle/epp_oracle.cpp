// oracle/epp_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the prefix-cache-aware Endpoint Picker algorithm that the
// reference deploys.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the
// product path (fusioninfer_b200/) never links, imports or calls it.
//
// PARITY STATUS: "parity unpinned by reference tests".  The arithmetic lives in
// a third-party module that is absent from /root/reference:
//   sigs.k8s.io/gateway-api-inference-extension v1.2.1   (/root/reference/go.mod:16,
//     consumed as image …/epp:v1.2.1, pkg/router/epp.go:46)
//   github.com/cespare/xxhash/v2 v2.3.0                   (/root/reference/go.mod:27)
// and the reference's own tests pin only YAML substrings
// (pkg/router/strategy_test.go:54-58,126-150).  What IS pinned: XXH64 against the
// published known answers and python `xxhash` 3.7.0 (tests/golden/, tests/
// test_oracle_xxh64.py), and the hash chain against an independent python
// restatement built on python `xxhash` (tests/golden/make_golden.py).
//
// Each function cites the SURVEY.md appendix paragraph it follows and the
// reference file:line that parameterises it.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (oracle/Makefile).

#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/fi_epp.h"

namespace {

// ----------------------------------------------------------------------------
// XXH64 — SURVEY.md Appendix A.7 (public xxHash specification; the Go module
// cespare/xxhash/v2 implements the same function).
// ----------------------------------------------------------------------------
constexpr uint64_t P1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t P2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t P3 = 0x165667B19E3779F9ULL;
constexpr uint64_t P4 = 0x85EBCA77C2B2AE63ULL;
constexpr uint64_t P5 = 0x27D4EB2F165667C5ULL;

inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t rd64(const uint8_t* p) {
  uint64_t v;
  std::memcpy(&v, p, 8);
  return v;  // little-endian host (x86-64)
}
inline uint32_t rd32(const uint8_t* p) {
  uint32_t v;
  std::memcpy(&v, p, 4);
  return v;
}
inline uint64_t xround(uint64_t acc, uint64_t x) { return rotl64(acc + x * P2, 31) * P1; }
inline uint64_t xmerge(uint64_t h, uint64_t v) { return (h ^ xround(0, v)) * P1 + P4; }

uint64_t xxh64(const uint8_t* p, size_t len, uint64_t seed) {
  const uint8_t* const end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t* const limit = end - 32;
    do {
      v1 = xround(v1, rd64(p));
      v2 = xround(v2, rd64(p + 8));
      v3 = xround(v3, rd64(p + 16));
      v4 = xround(v4, rd64(p + 24));
      p += 32;
    } while (p <= limit);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xmerge(h, v1);
    h = xmerge(h, v2);
    h = xmerge(h, v3);
    h = xmerge(h, v4);
  } else {
    h = seed + P5;
  }
  h += (uint64_t)len;
  while (p + 8 <= end) {
    h ^= xround(0, rd64(p));
    h = rotl64(h, 27) * P1 + P4;
    p += 8;
  }
  if (p + 4 <= end) {
    h ^= (uint64_t)rd32(p) * P1;
    h = rotl64(h, 23) * P2 + P3;
    p += 4;
  }
  while (p < end) {
    h ^= (uint64_t)(*p) * P5;
    h = rotl64(h, 11) * P1;
    ++p;
  }
  h ^= h >> 33;
  h *= P2;
  h ^= h >> 29;
  h *= P3;
  h ^= h >> 32;
  return h;
}

// ----------------------------------------------------------------------------
// Hash chain — SURVEY.md Appendix A.1 (upstream prefix.hashPrompt), sized by
// blockSize/hashBlockSize and maxPrefixBlocksToMatch of
// /root/reference/pkg/router/strategy.go:57-58,147-148.
//   h_i = XXH64(0, block_i ‖ LE64(h_{i-1})), h_0 = XXH64(0, model ‖ salt);
//   trailing partial block dropped; at most max_blocks blocks.
// ----------------------------------------------------------------------------
uint32_t hash_prompt(const uint8_t* p, uint64_t len, uint64_t h0, uint32_t B, uint32_t M, uint64_t* out,
                     std::vector<uint8_t>& tmp) {
  uint64_t nb = len / B;
  if (nb > M) nb = M;
  tmp.resize((size_t)B + 8);
  uint64_t prev = h0;
  for (uint64_t i = 0; i < nb; ++i) {
    std::memcpy(tmp.data(), p + i * B, B);
    std::memcpy(tmp.data() + B, &prev, 8);  // LE64(h_{i-1})
    prev = xxh64(tmp.data(), (size_t)B + 8, 0);
    out[i] = prev;
  }
  return (uint32_t)nb;
}

// ----------------------------------------------------------------------------
// Index — SURVEY.md Appendix A.2 (upstream indexer: hashToPods + per-pod LRU).
// hash → small endpoint set, open addressing; a key whose set became empty
// stays in the table and reads as ∅.
// ----------------------------------------------------------------------------
struct Slot {
  uint64_t key;
  uint32_t used;
  uint32_t n;
  uint32_t inl[2];
  std::vector<uint32_t>* ext;  // holds ALL members once n > 2
};

class PodIndex {
 public:
  PodIndex() { slots_.assign(1024, Slot{0, 0, 0, {0, 0}, nullptr}); }
  ~PodIndex() {
    for (auto& s : slots_) delete s.ext;
  }
  void reserve(uint64_t keys) {
    uint64_t want = 1024;
    while (want < keys * 2) want <<= 1;
    if (want > slots_.size()) rehash(want);
  }
  // returns pointer to members and count (count 0 == ∅)
  inline uint32_t get(uint64_t h, const uint32_t** members) const {
    uint64_t mask = slots_.size() - 1;
    uint64_t i = mix(h) & mask;
    for (;;) {
      const Slot& s = slots_[i];
      if (!s.used) return 0;
      if (s.key == h) {
        *members = s.ext ? s.ext->data() : s.inl;
        return s.n;
      }
      i = (i + 1) & mask;
    }
  }
  void set(uint64_t h, uint32_t e) {
    if ((used_ + 1) * 10 > slots_.size() * 6) rehash(slots_.size() * 2);
    Slot& s = find_or_insert(h);
    uint32_t* m = s.ext ? s.ext->data() : s.inl;
    for (uint32_t j = 0; j < s.n; ++j)
      if (m[j] == e) return;
    if (s.ext) {
      s.ext->push_back(e);
    } else if (s.n < 2) {
      s.inl[s.n] = e;
    } else {
      s.ext = new std::vector<uint32_t>(s.inl, s.inl + 2);
      s.ext->push_back(e);
    }
    ++s.n;
  }
  void clear(uint64_t h, uint32_t e) {
    uint64_t mask = slots_.size() - 1;
    uint64_t i = mix(h) & mask;
    for (;;) {
      Slot& s = slots_[i];
      if (!s.used) return;
      if (s.key == h) {
        uint32_t* m = s.ext ? s.ext->data() : s.inl;
        for (uint32_t j = 0; j < s.n; ++j) {
          if (m[j] == e) {
            m[j] = m[s.n - 1];
            --s.n;
            if (s.ext) s.ext->pop_back();
            return;
          }
        }
        return;
      }
      i = (i + 1) & mask;
    }
  }
  uint64_t keys() const { return used_; }

 private:
  static inline uint64_t mix(uint64_t h) {
    h ^= h >> 32;
    return h * 0x9E3779B97F4A7C15ULL >> 20;
  }
  Slot& find_or_insert(uint64_t h) {
    uint64_t mask = slots_.size() - 1;
    uint64_t i = mix(h) & mask;
    for (;;) {
      Slot& s = slots_[i];
      if (!s.used) {
        s.used = 1;
        s.key = h;
        s.n = 0;
        s.ext = nullptr;
        ++used_;
        return s;
      }
      if (s.key == h) return s;
      i = (i + 1) & mask;
    }
  }
  void rehash(uint64_t n) {
    std::vector<Slot> old;
    old.swap(slots_);
    slots_.assign(n, Slot{0, 0, 0, {0, 0}, nullptr});
    uint64_t mask = n - 1;
    for (auto& s : old) {
      if (!s.used) continue;
      uint64_t i = mix(s.key) & mask;
      while (slots_[i].used) i = (i + 1) & mask;
      slots_[i] = s;
    }
  }
  std::vector<Slot> slots_;
  uint64_t used_ = 0;
};

// Per-endpoint LRU — upstream podToLRU (hashicorp/golang-lru semantics: Add of an
// existing key moves it to the front; Add of a new key pushes front and, when
// len > capacity, removes the oldest).  Capacity = lruCapacityPerServer
// (/root/reference/pkg/router/strategy.go:59,149).
struct PodLRU {
  std::list<uint64_t> order;  // front = most recent
  std::unordered_map<uint64_t, std::list<uint64_t>::iterator> pos;
};

struct EpState {
  uint32_t role_mask = 0;
  double kv_util = 0.0;
  int32_t queue_depth = 0;
  uint32_t flags = 0;
  // lora-affinity-scorer inputs (upstream pod metrics ActiveModels / WaitingModels / MaxActiveModels)
  std::vector<uint64_t> active, waiting;
  uint32_t max_active = 0;
};

// upstream lora-affinity-scorer — SURVEY.md §8a row a11 (plugin type of
// /root/reference/pkg/router/strategy.go:100-113): 1.0 if the target adapter is active on the pod,
// 0.8 if the pod has room for another adapter, 0.6 if the adapter is queued there, else 0.
inline double lora_score(const EpState& e, uint64_t adapter) {
  bool active = false, waiting = false;
  for (uint64_t a : e.active) active |= a == adapter;
  for (uint64_t a : e.waiting) waiting |= a == adapter;
  if (active) return 1.0;
  if (e.active.size() + e.waiting.size() < e.max_active) return 0.8;
  return waiting ? 0.6 : 0.0;
}

struct Oracle {
  fi_epp_config cfg;
  std::vector<EpState> eps;
  PodIndex index;
  std::vector<PodLRU> lrus;
  std::string err;
};

// upstream by-label filter plugins chain: a pod stays a candidate iff it is alive and every filter of the
// profile finds one of its validValues among the pod's labels (/root/reference/pkg/router/strategy.go:135-144)
inline bool eligible(const EpState& e, const fi_profile& p) {
  if (!(e.flags & FI_ENDPOINT_ALIVE)) return false;
  if (p.role_mask != 0 && (e.role_mask & p.role_mask) == 0) return false;
  for (uint32_t f = 0; f < p.n_more_filters; ++f)
    if ((e.role_mask & p.more_filters[f]) == 0) return false;
  return true;
}

inline double clamp01(double s) { return s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s); }

// Per-profile, per-batch constants of the queue scorer — SURVEY.md Appendix A.4:
// min/max waiting-queue size over the FILTERED candidate set.
struct ProfileCtx {
  int32_t min_q = 0, max_q = 0;
  bool any = false;
};

ProfileCtx make_ctx(const Oracle& o, const fi_profile& p) {
  ProfileCtx c;
  for (const EpState& e : o.eps) {
    if (!eligible(e, p)) continue;
    if (!c.any) {
      c.min_q = c.max_q = e.queue_depth;
      c.any = true;
    } else {
      c.min_q = std::min(c.min_q, e.queue_depth);
      c.max_q = std::max(c.max_q, e.queue_depth);
    }
  }
  return c;
}

// total[e] = Σ_scorers clamp01(score)·(double)weight,
// accumulated in profile order in fp64 without FMA contraction
// (weights: /root/reference/pkg/router/strategy.go:66,82,97,157,163).
inline double total_score(const fi_profile& p, const ProfileCtx& c, const EpState& e, uint32_t match,
                          uint32_t n_blocks, uint64_t adapter) {
  double total = 0.0;
  for (uint32_t s = 0; s < p.n_scorers; ++s) {
    double sc = 0.0;
    switch (p.scorers[s].kind) {
      case FI_SCORER_PREFIX:  // upstream prefix Plugin.Score: matchLen / total blocks
        sc = n_blocks ? (double)match / (double)n_blocks : 0.0;
        break;
      case FI_SCORER_KV_UTIL:  // upstream kv-cache-utilization-scorer: 1 − usage
        sc = 1.0 - e.kv_util;
        break;
      case FI_SCORER_QUEUE:  // upstream queue-scorer: (max − q)/(max − min), 1.0 if all equal
        sc = (c.max_q == c.min_q) ? 1.0
                                  : (double)((int64_t)c.max_q - (int64_t)e.queue_depth) /
                                        (double)((int64_t)c.max_q - (int64_t)c.min_q);
        break;
      case FI_SCORER_LORA:
        sc = lora_score(e, adapter);
        break;
      default:
        sc = 0.0;
    }
    total = total + clamp01(sc) * (double)p.scorers[s].weight;
  }
  return total;
}

// ----------------------------------------------------------------------------
// Tie order — SURVEY.md Appendix A.5.  Upstream's MaxScorePicker shuffles the candidates before its stable
// sort, so equal totals are resolved at random.  To stay reproducible (and comparable bit for bit with the GPU
// path) the order among tied endpoints is a rotation of the pool whose start is derived from the request —
// the rule stated in include/fi_epp.h ("Ties"), restated here independently of the kernels' tiebreak.cuh:
//   seed  = n_blocks > 0 ? first chained block hash : h0 ^ (r + 1)·0x9E3779B97F4A7C15   (r = index in the call)
//   start = (splitmix64_finalise(seed) >> 32) · E >> 32
//   the tied endpoint with the smallest (e − start) mod E wins.
// ----------------------------------------------------------------------------
inline uint32_t tie_rotation_start(uint32_t n_blocks, uint64_t first_hash, uint64_t h0, uint32_t r, uint32_t E) {
  uint64_t x = n_blocks ? first_hash : (h0 ^ ((uint64_t)r + 1) * 0x9E3779B97F4A7C15ULL);
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  x = x ^ (x >> 31);
  return (uint32_t)(((x >> 32) * (uint64_t)E) >> 32);
}
inline uint32_t tie_distance(uint32_t e, uint32_t start, uint32_t E) { return (e + E - start) % E; }

struct Scratch {
  std::vector<uint16_t> match;   // per endpoint
  std::vector<uint32_t> touched;  // endpoints with match > 0
  std::vector<uint32_t> alive, alive2;
  std::vector<uint64_t> chain;
  std::vector<uint8_t> tmp;
};

// One request: hash → match (A.3) → score (A.4) → pick (A.5) → PD (A.6).
void pick_one(const Oracle& o, const std::vector<ProfileCtx>& ctx, const uint8_t* prompt, uint64_t len,
              uint64_t h0, uint64_t adapter, uint32_t r, fi_pick* out, uint64_t* chain_out, Scratch& sc) {
  const fi_epp_config& cfg = o.cfg;
  const uint32_t E = cfg.num_endpoints;
  sc.chain.resize(cfg.max_blocks);
  uint32_t n = hash_prompt(prompt, len, h0, cfg.block_bytes, cfg.max_blocks, sc.chain.data(), sc.tmp);
  if (chain_out) {
    std::memcpy(chain_out, sc.chain.data(), (size_t)n * 8);
    for (uint32_t i = n; i < cfg.max_blocks; ++i) chain_out[i] = 0;
  }
  if (sc.match.size() != E) sc.match.assign(E, 0);
  sc.touched.clear();

  if (cfg.match_mode == FI_MATCH_UPSTREAM) {
    // upstream Plugin.matchLongestPrefix: for i: s = Get(h_i); if ∅ break; res[p]++ ∀ p ∈ s
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t* m = nullptr;
      uint32_t cnt = o.index.get(sc.chain[i], &m);
      if (cnt == 0) break;
      for (uint32_t j = 0; j < cnt; ++j) {
        if (sc.match[m[j]]++ == 0) sc.touched.push_back(m[j]);
      }
    }
  } else {
    // LPM: match[e] = max{m : e ∈ Get(h_i) ∀ i < m}
    sc.alive.clear();
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t* m = nullptr;
      uint32_t cnt = o.index.get(sc.chain[i], &m);
      if (cnt == 0) break;
      if (i == 0) {
        sc.alive.assign(m, m + cnt);
      } else {
        sc.alive2.clear();
        for (uint32_t a : sc.alive)
          for (uint32_t j = 0; j < cnt; ++j)
            if (m[j] == a) {
              sc.alive2.push_back(a);
              break;
            }
        sc.alive.swap(sc.alive2);
      }
      if (sc.alive.empty()) break;
      for (uint32_t a : sc.alive)
        if (sc.match[a]++ == 0) sc.touched.push_back(a);
    }
  }

  const uint32_t start = tie_rotation_start(n, n ? sc.chain[0] : 0, h0, r, E);
  for (uint32_t p = 0; p < cfg.n_profiles; ++p) {
    const fi_profile& prof = cfg.profiles[p];
    double best = 0.0;
    uint32_t best_e = FI_NO_ENDPOINT;
    for (uint32_t e = 0; e < E; ++e) {  // max total; equal totals: the request's tie rotation decides
      const EpState& es = o.eps[e];
      if (!eligible(es, prof)) continue;
      double t = total_score(prof, ctx[p], es, sc.match[e], n, adapter);
      if (best_e == FI_NO_ENDPOINT || t > best ||
          (t == best && tie_distance(e, start, E) < tie_distance(best_e, start, E))) {
        best = t;
        best_e = e;
      }
    }
    fi_pick& pk = out[p];
    pk.endpoint = best_e;
    pk.match_blocks = best_e == FI_NO_ENDPOINT ? 0 : sc.match[best_e];
    pk.n_blocks = (uint16_t)n;
    pk.score = best_e == FI_NO_ENDPOINT ? 0.0 : best;
  }

  // SURVEY.md Appendix A.6 (pd-profile-handler, /root/reference/pkg/router/strategy.go:129-133):
  // decode first; prefill runs iff (1 − hit)·len(prompt) ≥ threshold.
  if (cfg.pd_enabled) {
    const fi_pick& d = out[cfg.pd_decode_profile];
    double hit = (d.endpoint != FI_NO_ENDPOINT && n) ? (double)d.match_blocks / (double)n : 0.0;
    double miss_bytes = (1.0 - hit) * (double)len;
    if (!(miss_bytes >= cfg.pd_threshold)) {
      fi_pick& pf = out[cfg.pd_prefill_profile];
      pf.endpoint = FI_NO_ENDPOINT;
      pf.match_blocks = 0;
      pf.score = 0.0;
    }
  }

  for (uint32_t e : sc.touched) sc.match[e] = 0;
}

bool validate(const fi_epp_config& c, std::string& err) {
  if (c.struct_size != sizeof(fi_epp_config)) return err = "struct_size mismatch", false;
  if (c.block_bytes == 0 || c.max_blocks == 0 || c.max_blocks > FI_EPP_MAX_BLOCKS)
    return err = "block_bytes/max_blocks out of range", false;
  if (c.num_endpoints == 0) return err = "num_endpoints == 0", false;
  if (c.n_profiles == 0 || c.n_profiles > FI_EPP_MAX_PROFILES) return err = "n_profiles out of range", false;
  for (uint32_t p = 0; p < c.n_profiles; ++p) {
    if (c.profiles[p].n_scorers > FI_EPP_MAX_SCORERS) return err = "n_scorers out of range", false;
    for (uint32_t s = 0; s < c.profiles[p].n_scorers; ++s) {
      uint32_t k = c.profiles[p].scorers[s].kind;
      if (k != FI_SCORER_PREFIX && k != FI_SCORER_KV_UTIL && k != FI_SCORER_QUEUE && k != FI_SCORER_LORA)
        return err = "unsupported scorer kind", false;
      if (c.profiles[p].scorers[s].weight < 0) return err = "negative weight", false;
    }
  }
  if (c.pd_enabled && (c.pd_decode_profile >= c.n_profiles || c.pd_prefill_profile >= c.n_profiles))
    return err = "pd profile index out of range", false;
  return true;
}

}  // namespace

extern "C" {

// The defaults of generatePrefixCacheConfig (/root/reference/pkg/router/strategy.go:51-68) with 64-byte blocks
//  — so that the CPU legs can build their configuration without loading the product library.
int epo_config_default(fi_epp_config* c) {
  if (!c) return FI_ERR_INVALID;
  std::memset(c, 0, sizeof(*c));
  c->struct_size = sizeof(*c);
  c->abi_version = FI_EPP_ABI_VERSION;
  c->block_bytes = 64;
  c->max_blocks = 256;      // strategy.go:58
  c->lru_capacity = 31250;  // strategy.go:59
  c->num_endpoints = 1;
  c->endpoint_count = 1;
  c->match_mode = FI_MATCH_UPSTREAM;
  c->max_batch = 1024;
  c->n_profiles = 1;
  std::snprintf(c->profiles[0].name, sizeof(c->profiles[0].name), "default");
  c->profiles[0].n_scorers = 1;
  c->profiles[0].scorers[0].kind = FI_SCORER_PREFIX;
  c->profiles[0].scorers[0].weight = 100;  // strategy.go:66
  return FI_OK;
}

uint64_t epo_xxh64(const void* data, uint64_t len, uint64_t seed) {
  return xxh64((const uint8_t*)data, (size_t)len, seed);
}

void* epo_create(const fi_epp_config* cfg) {
  std::string err;
  if (!cfg || !validate(*cfg, err)) {
    std::fprintf(stderr, "epo_create: %s\n", err.c_str());
    return nullptr;
  }
  Oracle* o = new Oracle();
  o->cfg = *cfg;
  o->eps.assign(cfg->num_endpoints, EpState{});
  if (cfg->lru_capacity) o->lrus.resize(cfg->num_endpoints);
  return o;
}

void epo_destroy(void* h) { delete (Oracle*)h; }

int epo_endpoints_update(void* h, const fi_endpoint_state* s, uint32_t n) {
  Oracle* o = (Oracle*)h;
  for (uint32_t i = 0; i < n; ++i) {
    if (s[i].endpoint >= o->cfg.num_endpoints) return FI_ERR_INVALID;
    EpState& e = o->eps[s[i].endpoint];
    e.role_mask = s[i].role_mask;
    e.kv_util = s[i].kv_util;
    e.queue_depth = s[i].queue_depth;
    e.flags = s[i].flags;
  }
  return FI_OK;
}

int epo_endpoints_lora_update(void* h, const fi_endpoint_lora* s, uint32_t n) {
  Oracle* o = (Oracle*)h;
  for (uint32_t i = 0; i < n; ++i) {
    if (s[i].endpoint >= o->cfg.num_endpoints || s[i].n_active > FI_EPP_MAX_LORA || s[i].n_waiting > FI_EPP_MAX_LORA)
      return FI_ERR_INVALID;
    EpState& e = o->eps[s[i].endpoint];
    e.active.assign(s[i].active, s[i].active + s[i].n_active);
    e.waiting.assign(s[i].waiting, s[i].waiting + s[i].n_waiting);
    e.max_active = s[i].max_active;
  }
  return FI_OK;
}

int epo_index_reserve(void* h, uint64_t keys) {
  ((Oracle*)h)->index.reserve(keys);
  return FI_OK;
}

int epo_index_apply(void* h, const fi_index_op* ops, uint64_t n) {
  Oracle* o = (Oracle*)h;
  for (uint64_t i = 0; i < n; ++i) {
    if (ops[i].endpoint >= o->cfg.num_endpoints) return FI_ERR_INVALID;
    if (ops[i].op == FI_OP_SET)
      o->index.set(ops[i].hash, ops[i].endpoint);
    else if (ops[i].op == FI_OP_CLEAR)
      o->index.clear(ops[i].hash, ops[i].endpoint);
    else
      return FI_ERR_INVALID;
  }
  return FI_OK;
}

// upstream indexer.Add(hashes, pod) — SURVEY.md Appendix A.2
int epo_index_add_chain(void* h, uint32_t endpoint, const uint64_t* hashes, uint32_t n) {
  Oracle* o = (Oracle*)h;
  if (!o->cfg.lru_capacity || endpoint >= o->cfg.num_endpoints) return FI_ERR_INVALID;
  PodLRU& l = o->lrus[endpoint];
  for (uint32_t i = 0; i < n; ++i) {
    uint64_t k = hashes[i];
    auto it = l.pos.find(k);
    if (it != l.pos.end()) {
      l.order.splice(l.order.begin(), l.order, it->second);
      continue;
    }
    l.order.push_front(k);
    l.pos[k] = l.order.begin();
    o->index.set(k, endpoint);
    if (l.order.size() > o->cfg.lru_capacity) {
      uint64_t old = l.order.back();
      l.order.pop_back();
      l.pos.erase(old);
      o->index.clear(old, endpoint);
    }
  }
  return FI_OK;
}

// a batch of decisions, sequentially in request order (what fi_epp_index_add_chains must equal)
int epo_index_add_chains(void* h, const uint32_t* endpoints, const uint64_t* chains, uint32_t pitch,
                         const uint32_t* nblocks, uint32_t R) {
  for (uint32_t r = 0; r < R; ++r) {
    if (endpoints[r] == FI_NO_ENDPOINT || nblocks[r] == 0) continue;
    int rc = epo_index_add_chain(h, endpoints[r], chains + (size_t)r * pitch, nblocks[r]);
    if (rc != FI_OK) return rc;
  }
  return FI_OK;
}

uint64_t epo_index_keys(void* h) { return ((Oracle*)h)->index.keys(); }

// membership probe for tests: 1 if (endpoint, hash) is in the logical index
int epo_index_contains(void* h, uint32_t endpoint, uint64_t hash) {
  const uint32_t* m = nullptr;
  uint32_t cnt = ((Oracle*)h)->index.get(hash, &m);
  for (uint32_t j = 0; j < cnt; ++j)
    if (m[j] == endpoint) return 1;
  return 0;
}

int epo_hash_batch(void* h, const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0, uint32_t R,
                   uint64_t* chains_out, uint32_t* nblocks_out) {
  Oracle* o = (Oracle*)h;
  std::vector<uint8_t> tmp;
  std::vector<uint64_t> chain(o->cfg.max_blocks);
  for (uint32_t r = 0; r < R; ++r) {
    uint32_t n = hash_prompt(prompts + offsets[r], offsets[r + 1] - offsets[r], h0[r], o->cfg.block_bytes,
                             o->cfg.max_blocks, chain.data(), tmp);
    if (nblocks_out) nblocks_out[r] = n;
    if (chains_out) {
      uint64_t* dst = chains_out + (size_t)r * o->cfg.max_blocks;
      std::memcpy(dst, chain.data(), (size_t)n * 8);
      for (uint32_t i = n; i < o->cfg.max_blocks; ++i) dst[i] = 0;
    }
  }
  return FI_OK;
}

// host cores this process may actually use: the affinity mask, capped by the cgroup CPU quota if there is one
static unsigned usable_cores() {
  unsigned n = std::thread::hardware_concurrency();
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) n = (unsigned)CPU_COUNT(&set);
  if (FILE* f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {  // cgroup v2: "<quota|max> <period>"
    char q[64] = {0};
    long long period = 0;
    if (std::fscanf(f, "%63s %lld", q, &period) == 2 && std::strcmp(q, "max") != 0 && period > 0) {
      const long long quota = std::atoll(q);
      if (quota > 0) n = std::min<unsigned>(n, (unsigned)std::max<long long>(1, (quota + period - 1) / period));
    }
    std::fclose(f);
  }
  return n ? n : 1;
}

// Persistent workers (one pool per process): the timing legs call the batch entry points many times and must
// not pay a thread spawn per call.
class Pool {
 public:
  static Pool& get(unsigned n) {
    static std::mutex mu;
    static std::unique_ptr<Pool> inst;
    std::lock_guard<std::mutex> lk(mu);
    if (!inst || inst->n_ != n) inst.reset(new Pool(n));
    return *inst;
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void run(const std::function<void(unsigned)>& fn) {  // fn(worker) on every worker, the caller is worker 0
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      busy_ = n_ - 1;
      ++gen_;
    }
    cv_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return busy_ == 0; });
  }

 private:
  explicit Pool(unsigned n) : n_(n < 1 ? 1 : n) {
    for (unsigned w = 1; w < n_; ++w) th_.emplace_back([this, w] {
      uint64_t seen = 0;
      for (;;) {
        {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return gen_ != seen; });
          seen = gen_;
          if (stop_) return;
        }
        (*fn_)(w);
        std::lock_guard<std::mutex> lk(mu_);
        if (--busy_ == 0) done_.notify_one();
      }
    });
  }
  unsigned n_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(unsigned)>* fn_ = nullptr;
  unsigned busy_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// The whole path for R requests; requests are sharded over `nthreads` host
// threads (the index is read-only during a batch).  out: R*n_profiles picks.
static int pick_batch_impl(void* h, const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0,
                           const uint64_t* adapters, uint32_t R, fi_pick* out, uint64_t* chains_out, uint32_t nthreads,
                           uint32_t repeat) {
  Oracle* o = (Oracle*)h;
  const uint32_t P = o->cfg.n_profiles;
  std::vector<ProfileCtx> ctx(P);
  for (uint32_t p = 0; p < P; ++p) ctx[p] = make_ctx(*o, o->cfg.profiles[p]);
  if (nthreads == 0) nthreads = 1;
  if (nthreads > R) nthreads = R ? R : 1;
  auto work = [&](uint32_t lo, uint32_t hi) {
    Scratch sc;
    for (uint32_t rep = 0; rep < repeat; ++rep)
      for (uint32_t r = lo; r < hi; ++r) {
        pick_one(*o, ctx, prompts + offsets[r], offsets[r + 1] - offsets[r], h0[r], adapters ? adapters[r] : 0, r,
                 out + (size_t)r * P,
                 chains_out ? chains_out + (size_t)r * o->cfg.max_blocks : nullptr, sc);
      }
  };
  if (nthreads == 1) {
    work(0, R);
    return FI_OK;
  }
  const uint32_t per = (R + nthreads - 1) / nthreads;
  Pool::get(nthreads).run([&](unsigned t) {
    const uint32_t lo = t * per, hi = std::min(R, lo + per);
    if (lo < hi) work(lo, hi);
  });
  return FI_OK;
}

// cores the CPU legs may use (affinity mask ∩ cgroup quota) — what bench.py reports as `cores`
uint32_t epo_usable_cores(void) { return usable_cores(); }

int epo_pick_batch(void* h, const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0, uint32_t R,
                   fi_pick* out, uint64_t* chains_out, uint32_t nthreads) {
  return pick_batch_impl(h, prompts, offsets, h0, nullptr, R, out, chains_out, nthreads, 1);
}

int epo_pick_batch_lora(void* h, const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0,
                        const uint64_t* adapters, uint32_t R, fi_pick* out, uint64_t* chains_out, uint32_t nthreads) {
  return pick_batch_impl(h, prompts, offsets, h0, adapters, R, out, chains_out, nthreads, 1);
}

// Timing variant: every thread processes its shard `repeat` times (same results), so that thread
// start-up does not dominate a bounded sample on a many-core host.  Decisions = R * repeat.
int epo_pick_batch_repeat(void* h, const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0, uint32_t R,
                          fi_pick* out, uint32_t nthreads, uint32_t repeat) {
  return pick_batch_impl(h, prompts, offsets, h0, nullptr, R, out, nullptr, nthreads, repeat ? repeat : 1);
}

}  // extern "C"
