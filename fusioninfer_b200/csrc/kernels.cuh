// kernels.cuh — device-side data layout and launcher declarations of libfi_epp.
//
// HBM layout (DESIGN.md "Data layout"):
//   prompts   concatenated prompt bytes, request r = [offsets[r], offsets[r+1])
//   pre       tiled u64     block pre-states (hash_blocks → chain_finalize): 16-byte unit u of
//                           request r at ((r/32)*MP/2 + u)*32 + r%32 — coalesced for the chain walker
//   chain     [R][MP] u64   chained block hashes h_1..h_n
//   index     keys    [C]     u64 table, buckets of 4 keys = one 32 B sector; 0 = empty, ~0 = tombstone
//             node_of [C]     u32 node of the key in that slot
//             klog    [C+3]   u64 key of node n (0: free / retired); nodes are numbered in insertion order,
//                             nodes C, C+1 belong to the hashes 0 and ~0, node C+2 is a never-written row
//             rows    [C+3][W] u32, row n = membership bitset of node n over the local endpoints
//                             (bit e%32 of word e/32)
//             cnt     [C+3]   u32 popcount of the row (this rank's endpoints holding the block)
//             rmask   [C+3]   u32 bit g set ⇔ rank g's row of this key is non-empty; key present ⇔ rmask != 0.
//                             One rank: rmask = (cnt > 0).  Endpoint-range shards: every rank's table is a
//                             DIRECTORY of the whole pool's keys (rows only for its own endpoints), kept exact by
//                             gossiping the owners' APPEAR / VANISH transitions (index_kernels.cu), so the first
//                             block NO endpoint holds — upstream's stopping point — is found locally.
//   picks     [R][P]        fi_pick
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fi_epp.h"

namespace fi {

constexpr uint32_t SLOT_MISS = 0xFFFFFFFFu;
constexpr uint64_t KEY_EMPTY = 0ull;
constexpr uint64_t KEY_TOMB = ~0ull;
constexpr int BUCKET_KEYS = 4;  // one 32-byte sector (64-byte buckets of 8 were measured slower: DESIGN.md)

struct IndexView {
  uint64_t* keys;
  uint32_t* node_of;
  uint64_t* klog;
  uint32_t* rows;
  uint32_t* cnt;
  uint32_t* rmask;  // per node: ranks whose local row is non-empty (bit = rank); key present ⇔ rmask != 0
  uint64_t bmask;   // buckets - 1
  uint64_t C;      // regular slots = buckets * BUCKET_KEYS
  uint32_t W;      // words per row (power of two)
  uint32_t logW;
};

// device-side counters of the index (one cache line)
struct IndexCounters {
  unsigned long long used;        // slots claimed = nodes allocated (keys + tombstones)
  unsigned long long tombstones;  // keys retired because no rank holds them any more
  unsigned long long overflow;    // != 0: an insert found no free slot
  unsigned long long pad;
};

// Transition log of one gossip round (sharded pools): the hashes whose LOCAL row went empty -> non-empty
// (appear) or non-empty -> empty (vanish) while this rank applied its own SET / CLEAR ops.  The other ranks
// replay them into their directories (rmask bit of this rank).
struct GossipLog {
  unsigned long long* n_appear;  // device counters (null: single rank, nothing is logged)
  unsigned long long* n_vanish;
  uint64_t* appear;              // [cap]
  uint64_t* vanish;              // [cap]
  uint64_t cap;
};

struct ProfileDev {
  uint32_t n_scorers;
  uint32_t n_filters;                    // by-label filters, ANDed (fi_profile: role_mask first, then more_filters)
  uint32_t filter[FI_EPP_MAX_FILTERS];
  uint32_t kind[FI_EPP_MAX_SCORERS];
  double weight[FI_EPP_MAX_SCORERS];
};

// alive and carrying, for every filter of the profile, at least one of its label bits
__host__ __device__ inline bool profile_admits(const ProfileDev& pr, uint32_t ep_flags, uint32_t ep_role_mask) {
  if (!(ep_flags & FI_ENDPOINT_ALIVE)) return false;
  for (uint32_t f = 0; f < pr.n_filters; ++f)
    if (!(ep_role_mask & pr.filter[f])) return false;
  return true;
}

// best total of a profile when no prefix block matches (per batch constant); the endpoints that attain it
// are the bit words ScoreTables::ztie — which of them wins depends on the request's tie rotation
struct ZeroBest {
  double score;
  uint32_t any;  // 0: no eligible local endpoint
  uint32_t pad;
};

struct EndpointDev {  // raw state of one endpoint of the GLOBAL pool
  double kv_util;
  int32_t queue_depth;
  uint32_t role_mask;
  uint32_t flags;
  uint32_t pad;
};

struct LoraDev {  // adapter residency of one LOCAL endpoint (lora-affinity-scorer)
  uint64_t active[FI_EPP_MAX_LORA];
  uint64_t waiting[FI_EPP_MAX_LORA];
  uint32_t n_active, n_waiting, max_active, pad;
};

struct ScoreTables {
  ProfileDev prof[FI_EPP_MAX_PROFILES];
  uint32_t n_profiles;
  uint32_t Epad;        // W * 32
  const double* sc;     // [P][S][Epad] clamp01'd per-endpoint scores of the non-prefix scorers
  const uint32_t* elig; // [P][W] eligibility bit words
  const ZeroBest* zero; // [P]
  const uint32_t* ztie; // [P][W] eligible local endpoints whose zero-match total equals zero[p].score
  const LoraDev* lora;  // [Epad] or null
  uint32_t has_lora;    // some profile has a lora-affinity-scorer: scores depend on the request's adapter,
  uint32_t pad;         // so every eligible endpoint is scored per request (no zero-match shortcut)
};

// Peer-memory exchange of the endpoint-range sharded mode (one buffer per rank, every rank's buffer
// mapped into every process over NVLink / CUDA IPC).  Low-latency protocol: every 32-bit datum travels
// in one aligned 64-bit store together with a 32-bit tag = the step number of the pick call, so a word is
// valid exactly when its tag matches — no fences, no flags, no barrier between the ranks.  The producer
// (match_pick_kernel: this rank's local pick) stores each word into slot [parity][own rank] of EVERY rank's
// buffer as soon as it exists; the consumer (merge_picks_kernel) polls only the words of the request it is
// about to reduce, so the transfer of later requests overlaps the work on earlier ones.  Parity double-buffers consecutive steps (a rank can
// be at most one step ahead of a peer, because its merge needs that peer's picks of the previous step).
constexpr int FI_MAX_RANKS = 16;
struct PeerXchg {
  uint32_t world, rank;
  uint32_t step;                // tag of this pick call (monotonic, starts at 1; buffers start zeroed)
  uint32_t enabled;             // 0: the NCCL all-gather path is used instead
  uint8_t* base[FI_MAX_RANKS];  // base[k] = rank k's exchange buffer as mapped here (base[rank] is local)
  uint64_t off_pick[2];         // u64 {tag:word}       [world][R][P][4]  endpoint, match_blocks, score lo, hi
  uint32_t* err;                // local: set to 1 if a poll timed out
};

struct MatchParams {
  const uint64_t* chain;
  const uint32_t* nblocks;
  const uint64_t* offsets;  // [R+1], prompt byte offsets (PD threshold); may be null if !apply_pd
  const uint64_t* adapters; // [R] target adapter id per request, or null (= id 0)
  uint32_t R;
  uint32_t MP;  // pitch of chain rows (multiple of 4)
  IndexView ix;
  ScoreTables st;
  uint32_t ep_begin;
  uint32_t ep_count;
  uint32_t E_global;  // pool size: ties rotate over the whole pool (tie_start)
  uint32_t r_base;    // index of this launch's request 0 within the caller's batch (sliced host feed)
  const uint64_t* h0; // [R] chain seeds (tie rotation of prompts shorter than one block)
  uint32_t lpm;  // fi_match_mode
  // pd-profile-handler
  uint32_t apply_pd, pd_decode, pd_prefill;
  double pd_threshold;
  fi_pick* out;                       // [R][P]
  unsigned long long* probed_blocks;  // optional Σ N_probe
  uint32_t* work_counter;             // dynamic request queue of the launch
  uint32_t zero_work_counter;         // launcher zeroes it first (0: the caller already did)
  uint32_t max_ctas_per_sm;           // 0: as many as fit; else a cap (pipelined API: leave room for hash_blocks)
  uint32_t lane_zero;                 // always 0: makes the ticket address formally lane-dependent (match_kernels.cu take_ticket)
  PeerXchg px;                        // sharded mode, peer-memory exchange (px.enabled)
};

struct MergeParams {
  const fi_pick* gathered;  // [ranks][R][P]
  uint32_t ranks, R, P;
  const uint32_t* nblocks;
  const uint64_t* offsets;
  const uint64_t* chain;  // [R][MP]  (tie rotation: first block hash)
  const uint64_t* h0;     // [R]
  uint32_t MP, E_global;
  uint32_t apply_pd, pd_decode, pd_prefill;
  double pd_threshold;
  fi_pick* out;  // [R][P]
  PeerXchg px;   // px.enabled: poll every rank's tagged pick words in-kernel instead of after an all-gather
};

// ---- launchers (each returns the cudaGetLastError() of its launch) -----------
// grid_cap: 0 = one CTA per request; else at most that many CTAs (the pipelined API shares the SMs with match_pick)
// zero_word: optional device word the kernel clears (the pipelined path's request-queue counter of the batch)
cudaError_t launch_hash_blocks(const uint8_t* prompts, const uint64_t* offsets, uint32_t R, uint32_t B,
                               uint32_t M, uint32_t MP, uint64_t* pre, uint32_t* nblocks, uint32_t grid_cap,
                               cudaStream_t s, uint32_t* zero_word = nullptr);
// compact: the 64-register / 24 KB shape whose 128 CTAs all fit on a 16-SM partition (pipelined path)
cudaError_t launch_chain_finalize(const uint64_t* pre, const uint32_t* nblocks, const uint64_t* h0,
                                  uint32_t R, uint32_t MP, uint64_t* chain, bool compact, cudaStream_t s);
cudaError_t launch_hash_generic(const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0,
                                uint32_t R, uint32_t B, uint32_t M, uint32_t MP, uint64_t* chain,
                                uint32_t* nblocks, cudaStream_t s);

cudaError_t launch_index_set(IndexView ix, IndexCounters* ctr, const fi_index_op* ops, uint64_t n,
                             uint32_t ep_begin, uint32_t ep_count, uint32_t rank, GossipLog log, cudaStream_t s);
cudaError_t launch_index_clear(IndexView ix, IndexCounters* ctr, const fi_index_op* ops, uint64_t n,
                               uint32_t ep_begin, uint32_t ep_count, uint32_t rank, GossipLog log, cudaStream_t s);
// the same with the op count read from device memory (ops produced by a kernel; cap = buffer capacity)
cudaError_t launch_index_clear_counted(IndexView ix, IndexCounters* ctr, const fi_index_op* ops, uint64_t cap,
                                       const unsigned long long* n_dev, uint32_t ep_begin, uint32_t ep_count, uint32_t rank,
                                       GossipLog log, cudaStream_t s);
// replay another rank's transitions into this rank's directory (n hashes; bit = that rank)
cudaError_t launch_index_remote_appear(IndexView ix, IndexCounters* ctr, const uint64_t* hashes, uint64_t n,
                                       uint32_t rank, cudaStream_t s);
cudaError_t launch_index_remote_vanish(IndexView ix, IndexCounters* ctr, const uint64_t* hashes, uint64_t n,
                                       uint32_t rank, cudaStream_t s);
cudaError_t launch_index_rebuild(IndexView from, IndexView to, IndexCounters* ctr, cudaStream_t s);
cudaError_t launch_index_contains(IndexView ix, const fi_index_op* q, uint64_t n, uint32_t ep_begin,
                                  uint32_t ep_count, uint8_t* out, cudaStream_t s);

cudaError_t launch_prepare_endpoints(const EndpointDev* eps, uint32_t E_global, uint32_t ep_begin,
                                     uint32_t ep_count, ScoreTables st, double* sc, uint32_t* elig,
                                     ZeroBest* zero, uint32_t* ztie, cudaStream_t s);

cudaError_t launch_match_pick(const MatchParams& p, int sm_count, cudaStream_t s);
cudaError_t launch_merge_picks(const MergeParams& p, cudaStream_t s);

}  // namespace fi
