/*
 * fi_epp.h — C ABI of the B200-native prefix-cache-aware Endpoint Picker.
 *
 * This is the drop-in boundary for the ONE hot path this repo implements
 * (BASELINE.json north_star): per request, hash the prompt into
 * chained fixed-size block keys, look the chain up against every candidate
 * endpoint's cached-prefix set, combine the match length with KV-cache /
 * queue load into a weighted fp64 score, and argmax to choose a pod.
 *
 * What it replaces.  FusionInfer (the reference, /root/reference) does not
 * contain this loop: its router role only *configures and deploys* the
 * Gateway-API-Inference-Extension EPP image
 *   pkg/router/epp.go:46        DefaultEPPImage  …/epp:v1.2.1
 *   pkg/router/epp.go:125-129   args --pool-name --pool-namespace --config-file
 *   pkg/router/strategy.go:51-68,115-165  the EndpointPickerConfig YAML
 * and the loop itself runs inside that image (upstream Go packages
 * pkg/epp/scheduling/framework/plugins/{multi/prefix,scorer,picker}).  There
 * is no cgo/FFI in the reference (Dockerfile:24 builds CGO_ENABLED=0), so the
 * entry points below are shaped after the upstream plugin seams a Go EPP
 * would bind through cgo:
 *
 *   upstream seam (Go, module sigs.k8s.io/gateway-api-inference-extension v1.2.1,
 *   go.mod:16)                                      → entry point here
 *   ------------------------------------------------------------------------
 *   config loader for the YAML of strategy.go:52-67 → fi_epp_config_from_yaml
 *   prefix.Plugin hashPrompt (xxhash chain)         → fi_epp_hash_batch
 *   prefix indexer.Add / LRU eviction (PreRequest)  → fi_epp_index_add_chains (a batch
 *                                                     of decisions), fi_epp_index_add_chain,
 *                                                     fi_epp_index_apply
 *   datastore pod metrics refresh (kv, queue, role) → fi_epp_endpoints_update
 *   SchedulerProfile.Run: filter → scorers → picker → fi_epp_pick_batch
 *   pd-profile-handler (decode then prefill)        → fi_epp_pick_batch with
 *                                                     cfg.pd_enabled
 *
 * Conventions (cgo-safe): every function returns 0 (FI_OK) or a negative
 * fi_status; nothing throws across the ABI; the caller owns every buffer and
 * no pointer is retained after a call returns; one fi_epp handle is internally
 * serialised by a mutex (concurrency comes from batching); library threads
 * never call back into the host language.  There is NO CPU fallback: without
 * a CUDA device fi_epp_create fails with FI_ERR_CUDA.
 *
 * Ties.  Upstream's MaxScorePicker shuffles the candidates before its stable sort, i.e. equal totals are
 * resolved at random.  Here the order among endpoints with equal totals is a rotation
 * of the pool that starts at a position derived from the request, so that picks are reproducible (and
 * bit-comparable with the CPU oracle) yet spread over the tied pods the way upstream's shuffle does:
 *     seed  = n_blocks > 0 ? h_1 (the first chained block hash) : h0[r] ^ (r + 1) * 0x9E3779B97F4A7C15
 *     x     = seed;  x ^= x >> 30;  x *= 0xBF58476D1CE4E5B9;  x ^= x >> 27;  x *= 0x94D049BB133111EB;  x ^= x >> 31
 *     start = ((x >> 32) * num_endpoints) >> 32
 *     among equal totals the endpoint with the smallest (endpoint - start) mod num_endpoints wins
 * (r = index of the request within the call).
 */
#ifndef FI_EPP_H_
#define FI_EPP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FI_EPP_ABI_VERSION 2u
#define FI_EPP_MAX_PROFILES 4u
#define FI_EPP_MAX_SCORERS 4u
#define FI_EPP_MAX_FILTERS 4u /* by-label filters per profile */
#define FI_EPP_MAX_LABELS 24u /* (label, value) pairs a configuration can filter on */
#define FI_EPP_MAX_BLOCKS 1023u /* counts are kept in 10 bit-planes on the GPU */
#define FI_NO_ENDPOINT 0xFFFFFFFFu
#define FI_EPP_UNIQUE_ID_BYTES 128u

typedef enum fi_status {
  FI_OK = 0,
  FI_ERR_INVALID = -1,  /* bad argument / config */
  FI_ERR_CUDA = -2,     /* CUDA runtime error, or no device (there is no CPU fallback) */
  FI_ERR_NOMEM = -3,
  FI_ERR_CAPACITY = -4, /* batch, prompt bytes or index larger than configured */
  FI_ERR_STATE = -5,
  FI_ERR_COMM = -6,     /* NCCL / peer-memory error */
  FI_ERR_CONFIG = -7    /* EndpointPickerConfig YAML rejected */
} fi_status;

/* SURVEY.md Appendix A.3.  UPSTREAM: stop at the first block no endpoint holds,
 * count per endpoint the blocks it holds before that.  LPM: per-endpoint longest
 * contiguous prefix.  Identical whenever every endpoint's set is prefix-closed. */
typedef enum fi_match_mode { FI_MATCH_UPSTREAM = 0, FI_MATCH_LPM = 1 } fi_match_mode;

/* plugin `type:` names of pkg/router/strategy.go:55,74,89,104 */
typedef enum fi_scorer_kind {
  FI_SCORER_PREFIX = 1,  /* prefix-cache-scorer          */
  FI_SCORER_KV_UTIL = 2, /* kv-cache-utilization-scorer  */
  FI_SCORER_QUEUE = 3,   /* queue-scorer                 */
  FI_SCORER_LORA = 4     /* lora-affinity-scorer         */
} fi_scorer_kind;

/* Label bits of an endpoint (fi_endpoint_state.role_mask) — what the by-label filters of strategy.go:135-144
 * test.  The three `fusioninfer.io/component-type` values (api/core/v1alpha1/inferenceservice_types.go:26-33)
 * have fixed bits; any other (label, value) pair a configuration filters on is given one of the bits from
 * FI_ROLE_FIRST_FREE up by fi_epp_config_from_yaml, which records the assignment in fi_epp_config.labels —
 * the host sets that bit on every pod carrying the label value when it calls fi_epp_endpoints_update. */
#define FI_ROLE_WORKER 1u
#define FI_ROLE_PREFILLER 2u
#define FI_ROLE_DECODER 4u
#define FI_ROLE_FIRST_FREE 8u

#define FI_ENDPOINT_ALIVE 1u

typedef struct fi_scorer {
  uint32_t kind;  /* fi_scorer_kind */
  int32_t weight; /* `weight:` of the pluginRef (strategy.go:66,157,163); >= 0 */
} fi_scorer;

typedef struct fi_profile {
  char name[32];      /* schedulingProfiles[].name */
  uint32_t role_mask; /* first by-label filter (0: none): the endpoint must carry one of these label bits */
  uint32_t n_scorers;
  fi_scorer scorers[FI_EPP_MAX_SCORERS]; /* in profile order: fp64 accumulation order */
  /* further by-label filters of the profile.  Filters chain like upstream's filter plugins: an endpoint is
   * eligible iff it is alive and passes EVERY filter f, i.e. (ep.role_mask & f) != 0. */
  uint32_t n_more_filters;
  uint32_t more_filters[FI_EPP_MAX_FILTERS - 1];
} fi_profile;

/* One (label, value) pair of a by-label filter and the endpoint bit that stands for it. */
typedef struct fi_label_bit {
  char label[64]; /* e.g. "fusioninfer.io/component-type" */
  char value[56]; /* one of the filter's validValues */
  uint32_t bit;   /* single bit of fi_endpoint_state.role_mask */
  uint32_t reserved;
} fi_label_bit;

typedef struct fi_epp_config {
  uint32_t struct_size; /* sizeof(fi_epp_config), checked */
  uint32_t abi_version; /* FI_EPP_ABI_VERSION */
  int32_t device;       /* CUDA ordinal */
  uint32_t block_bytes; /* blockSize | hashBlockSize (strategy.go:57,147); 64 = 16 u32 tokens */
  uint32_t max_blocks;  /* maxPrefixBlocksToMatch (strategy.go:58,148) */
  uint32_t lru_capacity; /* lruCapacityPerServer (strategy.go:59,149); 0: no LRU (fi_epp_index_add_chain* fail) */
  uint32_t num_endpoints;  /* global pool size E */
  uint32_t endpoint_begin; /* this handle's shard [begin, begin+count) of the pool */
  uint32_t endpoint_count;
  uint32_t match_mode; /* fi_match_mode */
  uint32_t max_batch;  /* largest R accepted by one pick/hash call */
  uint32_t reserved0;
  uint64_t max_prompt_bytes; /* largest total prompt bytes per call (device staging) */
  uint64_t index_slots;      /* key slots of the GPU index, power of two; 0 = 2x num_endpoints*lru_capacity (load <= 0.5: a shard
                              * is a directory of the whole pool's keys, with membership rows for its own endpoints) */
  uint32_t n_profiles;
  uint32_t pd_enabled;        /* pd-profile-handler present (strategy.go:129-133) */
  uint32_t pd_decode_profile; /* profile index run first */
  uint32_t pd_prefill_profile;
  double pd_threshold;        /* `threshold:` — prefill runs iff (1-hit)*len(prompt) >= threshold */
  fi_profile profiles[FI_EPP_MAX_PROFILES];
  /* written by fi_epp_config_from_yaml, read by the host: which role_mask bit each filtered (label, value) has */
  uint32_t n_labels;
  uint32_t reserved1;
  fi_label_bit labels[FI_EPP_MAX_LABELS];
} fi_epp_config;

/* One row of the pod datastore the scorers read (upstream metrics refresh). */
typedef struct fi_endpoint_state {
  uint32_t endpoint;   /* global index in [0, num_endpoints) */
  uint32_t role_mask;  /* FI_ROLE_* */
  double kv_util;      /* KVCacheUsagePercent in [0,1] */
  int32_t queue_depth; /* WaitingQueueSize */
  uint32_t flags;      /* FI_ENDPOINT_ALIVE */
} fi_endpoint_state;

/* LoRA adapters resident / queued on one endpoint (upstream pod metrics ActiveModels,
 * WaitingModels, MaxActiveModels) — read by the lora-affinity-scorer
 * (pkg/router/strategy.go:100-113).  Adapter ids are any 64-bit ids the host uses
 * consistently, e.g. fi_epp_model_seed(adapter name). */
#define FI_EPP_MAX_LORA 8u
typedef struct fi_endpoint_lora {
  uint32_t endpoint;   /* global index */
  uint32_t max_active; /* MaxActiveModels */
  uint32_t n_active;   /* <= FI_EPP_MAX_LORA */
  uint32_t n_waiting;  /* <= FI_EPP_MAX_LORA */
  uint64_t active[FI_EPP_MAX_LORA];
  uint64_t waiting[FI_EPP_MAX_LORA];
} fi_endpoint_lora;

typedef enum fi_index_opcode { FI_OP_SET = 1, FI_OP_CLEAR = 2 } fi_index_opcode;

/* One membership change of the logical index {(endpoint, block hash)}. */
typedef struct fi_index_op {
  uint64_t hash;
  uint32_t endpoint; /* global index */
  uint32_t op;       /* fi_index_opcode */
} fi_index_op;

/* One routing decision (16 bytes). */
typedef struct fi_pick {
  uint32_t endpoint;     /* global index, or FI_NO_ENDPOINT */
  uint16_t match_blocks; /* prefix blocks matched at the picked endpoint */
  uint16_t n_blocks;     /* blocks hashed for this request */
  double score;          /* weighted fp64 total of the picked endpoint */
} fi_pick;

typedef struct fi_index_stats {
  uint64_t slots;      /* key slots */
  uint64_t used;       /* slots holding a key or a tombstone */
  uint64_t tombstones; /* keys whose row became empty */
  uint64_t rebuilds;
  uint64_t ops_applied;
  uint64_t lru_entries; /* entries of the per-endpoint LRUs (device-resident or host), summed */
} fi_index_stats;

typedef struct fi_epp_stats {
  uint64_t kernel_launches; /* kernels of this library launched since create/reset */
  uint64_t pick_calls;
  uint64_t requests;
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  /* per-kernel device time, accumulated only while profiling is on */
  double ms_hash_blocks, ms_chain_probe, ms_match_pick, ms_index_apply, ms_other;
  uint64_t n_hash_blocks, n_chain_probe, n_match_pick, n_index_apply, n_other;
  uint64_t probed_blocks; /* sum over requests of N_probe, profiling only */
} fi_epp_stats;

typedef struct fi_epp fi_epp;

uint32_t fi_epp_abi_version(void);
const char* fi_epp_status_string(int status);

/* Fill cfg with the defaults of generatePrefixCacheConfig (strategy.go:51-68)
 * except block_bytes = 64 (16 uint32 tokens). */
int fi_epp_config_default(fi_epp_config* cfg);

/* Parse an EndpointPickerConfig YAML document (exactly the schema
 * strategy.go:52-67,126-164 emits, plus custom passthrough strategy.go:29-31)
 * into cfg's plugin-derived fields (block_bytes, max_blocks, lru_capacity,
 * profiles, pd_*).  Deployment fields (device, endpoints, sizes) are untouched.
 * On FI_ERR_CONFIG a message is written to err (NUL-terminated, truncated). */
int fi_epp_config_from_yaml(const char* yaml, size_t len, fi_epp_config* cfg, char* err, size_t err_len);

int fi_epp_create(const fi_epp_config* cfg, fi_epp** out);
void fi_epp_destroy(fi_epp* h);
const char* fi_epp_last_error(const fi_epp* h);

/* h0 = XXH64(seed 0, model ‖ salt): the chain seed of SURVEY.md Appendix A.1. */
int fi_epp_model_seed(const void* model, size_t model_len, const void* salt, size_t salt_len, uint64_t* h0);

/* Replace the state of the listed endpoints (every rank receives the whole
 * pool: queue min/max are global).  Unlisted endpoints keep their state;
 * endpoints never listed are not alive. */
int fi_epp_endpoints_update(fi_epp* h, const fi_endpoint_state* states, uint32_t n);

/* Adapter residency of the listed endpoints (lora-affinity-scorer).  Endpoints never listed
 * hold no adapter and have max_active 0. */
int fi_epp_endpoints_lora_update(fi_epp* h, const fi_endpoint_lora* states, uint32_t n);

/* Asynchronous, ordered: every op submitted before a pick call is visible to
 * that pick.  Ops for endpoints outside this handle's shard are ignored.
 *
 * Sharded pools (after fi_epp_comm_init with world > 1): every rank's index is a directory of the WHOLE pool's
 * block hashes — membership rows only for its own endpoints — so that "the first block no pod holds" (where
 * upstream's matchLongestPrefix stops) is a local lookup.  The owner of an endpoint applies its ops and the
 * resulting key appear/vanish transitions are exchanged between the ranks inside this call: index updates of
 * a sharded pool are COLLECTIVE — every rank calls fi_epp_index_apply / fi_epp_index_add_chains the same
 * number of times in the same order, each with its own ops (possibly none). */
int fi_epp_index_apply(fi_epp* h, const fi_index_op* ops, uint64_t n);

/* Upstream indexer.Add(hashes, pod): touch each hash in `endpoint`'s LRU
 * (capacity lru_capacity), emit SET for new entries and CLEAR for evicted
 * ones.  Requires lru_capacity > 0.  Single-rank handles only (FI_ERR_STATE on a sharded pool). */
int fi_epp_index_add_chain(fi_epp* h, uint32_t endpoint, const uint64_t* hashes, uint32_t n);

/* The same for a whole batch of routing decisions — upstream's PreRequest step after a pick batch:
 * indexer.Add(chains[r*pitch_blocks .. +nblocks[r]), endpoints[r]) for r = 0..R-1 (FI_NO_ENDPOINT and
 * endpoints of other shards are skipped).  Equal to R fi_epp_index_add_chain calls in request order.  With the
 * device-resident LRU (the default, see below) the chains are copied to the GPU and the whole batch is applied
 * by a handful of kernels; with the host LRU the endpoints' LRUs are walked in parallel on host worker threads
 * (FI_EPP_LRU_THREADS, default = usable cores, at most 128).  `chains` / `nblocks` are what fi_epp_pick_batch
 * returned (chains_out, picks' n_blocks).  Collective on a sharded pool. */
int fi_epp_index_add_chains(fi_epp* h, const uint32_t* endpoints, const uint64_t* chains, uint32_t pitch_blocks,
                            const uint32_t* nblocks, uint32_t R);

/* The same with the chains already in device memory — the chains_out of fi_epp_pick_batch_device, written on
 * `stream` — so that only the two small host arrays cross PCIe.  Served by the device-resident LRU
 * (fusioninfer_b200/csrc/lru_kernels.cu), which is also what fi_epp_index_add_chain(s) use on a handle with
 * lru_capacity >= max_blocks (sharded pools included) unless fi_epp_set_option(h, "device_lru", 0) / FI_EPP_DEVICE_LRU=0
 * selected the host LRU before the first Add; FI_ERR_STATE when the handle runs the host LRU.
 * d_chains == NULL: the chains of this handle's most recent fi_epp_pick_batch / fi_epp_pick_batch_device call,
 * read from the handle's own buffer (pitch_blocks ignored; R <= that call's R) — the PreRequest step right after
 * a pick, with nothing but the decisions crossing PCIe. */
int fi_epp_index_add_chains_device(fi_epp* h, const uint32_t* endpoints, const void* d_chains, uint32_t pitch_blocks,
                                   const uint32_t* nblocks, uint32_t R, void* stream);

/* Diagnostics: the keys of `endpoint` in the device-resident LRU, least recently used first (at most `cap`
 * written, *n_out = how many it holds).  FI_ERR_STATE when the handle runs the host LRU. */
int fi_epp_lru_dump(fi_epp* h, uint32_t endpoint, uint64_t* out, uint32_t cap, uint32_t* n_out);

/* Diagnostics: totals of the device-resident LRU since create — out[0] SETs emitted, [1] CLEARs emitted, [2] keys
 * touched but gone again by the end of their batch, [3] per-endpoint maintenance passes (log compaction + table
 * rebuild), [4] requests deferred to a conservative pass (their endpoint's table could not take the batch's new
 * keys), [5] sub-batches run. */
int fi_epp_lru_counters(fi_epp* h, uint64_t out[6]);

int fi_epp_index_sync(fi_epp* h); /* block until submitted ops are applied */

/* Diagnostics: out[i] = 1 iff (q[i].endpoint, q[i].hash) is in this handle's GPU index. */
int fi_epp_index_contains(fi_epp* h, const fi_index_op* q, uint64_t n, uint8_t* out);
int fi_epp_index_stats(fi_epp* h, fi_index_stats* out);

/* Hash only.  prompts: concatenated prompt bytes; offsets: R+1 byte offsets;
 * h0: R chain seeds.  chains_out: R*max_blocks hashes (row r holds
 * nblocks_out[r] valid entries); either output may be NULL. */
int fi_epp_hash_batch(fi_epp* h, const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0,
                      uint32_t R, uint64_t* chains_out, uint32_t* nblocks_out);

/* The hot path, host buffers (pinned memory from fi_epp_pinned_alloc avoids a
 * staging copy).  out: R*n_profiles picks, out[r*n_profiles + p] for profile p.
 * With pd_enabled the prefill profile's pick is FI_NO_ENDPOINT when the
 * threshold test skips it.  chains_out optional (R*max_blocks). */
int fi_epp_pick_batch(fi_epp* h, const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0,
                      uint32_t R, fi_pick* out, uint64_t* chains_out);

/* Same, every buffer already in device memory of cfg.device; work is ordered
 * after `stream` (a cudaStream_t, may be NULL) and `stream` waits for it. */
int fi_epp_pick_batch_device(fi_epp* h, const void* d_prompts, const void* d_offsets, const void* d_h0,
                             uint32_t R, uint64_t total_prompt_bytes, void* d_out, void* d_chains_out,
                             void* stream);

/* The same two calls with one target adapter id per request (lora-affinity-scorer: 1.0 if the
 * adapter is active on the endpoint, 0.8 if the endpoint has room for another adapter, 0.6 if it
 * is queued there, else 0).  adapters == NULL means "no adapter" (id 0) for every request. */
int fi_epp_pick_batch_lora(fi_epp* h, const uint8_t* prompts, const uint64_t* offsets, const uint64_t* h0,
                           const uint64_t* adapters, uint32_t R, fi_pick* out, uint64_t* chains_out);
int fi_epp_pick_batch_device_lora(fi_epp* h, const void* d_prompts, const void* d_offsets, const void* d_h0,
                                  const void* d_adapters, uint32_t R, uint64_t total_prompt_bytes, void* d_out,
                                  void* d_chains_out, void* stream);

/* Pipelined device path.  fi_epp_pick_submit enqueues one batch exactly like fi_epp_pick_batch_device (inputs
 * ready in `stream` order at the call) but does NOT order `stream` behind the result: batch k+1's block
 * hashing and chain walk run while batch k is still being matched (two batches in flight, internal
 * streams).  The inputs and `d_out` of a submitted batch must stay untouched until a fi_epp_pick_wait
 * issued after it: that call makes `stream` wait (on the device; the host does not block) for every batch
 * submitted so far.  Batches complete in submission order and see the index as of their submit call.
 * Sharded handles and block sizes that are not a multiple of 32 take the stream-ordered path inside. */
int fi_epp_pick_submit(fi_epp* h, const void* d_prompts, const void* d_offsets, const void* d_h0, uint32_t R,
                       uint64_t total_prompt_bytes, void* d_out, void* stream);
int fi_epp_pick_wait(fi_epp* h, void* stream);
/* How the pipelined path runs: out[0] = 1 if the GPU is partitioned (green contexts: the chain walk of batch k+1 on
 * its own out[1] SMs while batch k is matched and batch k+2 hashed on the other out[2] — three batches in flight),
 * 0 if not (driver without green contexts, option "pipe_partition" = 0, sharded pool): two batches in flight on
 * the whole GPU.  Valid after the first fi_epp_pick_submit. */
int fi_epp_pipeline_info(fi_epp* h, int32_t out[3]);

void* fi_epp_pinned_alloc(size_t bytes);
void fi_epp_pinned_free(void* p);

/* Multi-GPU (one handle per GPU, endpoint-range shards).  Rank 0 makes an id,
 * the host distributes it out of band, every rank calls comm_init. */
int fi_epp_comm_unique_id(uint8_t out[FI_EPP_UNIQUE_ID_BYTES]);
int fi_epp_comm_init(fi_epp* h, const uint8_t id[FI_EPP_UNIQUE_ID_BYTES], uint32_t rank, uint32_t world);
/* Must precede the first index update of the handle.  How the sharded pick reduces the ranks' local
 * (score, endpoint) picks: FI_EXCHANGE_NONE (one rank), FI_EXCHANGE_PEER (default: the match kernel stores its
 * pick into every rank's buffer over NVLink peer memory / CUDA IPC and the merge kernel polls tagged words —
 * no collective call for the reduction) or FI_EXCHANGE_NCCL (one ncclAllGather: env FI_EPP_EXCHANGE=nccl,
 * more than 16 ranks, or a rank that cannot map a peer's buffer).  Every rank hashes every prompt (env
 * FI_EPP_SHARD_HASH=split: hashing split over the ranks and the chains all-gathered instead — slower on NVLink-
 * connected B200s: 16 KiB from local HBM cost less than 2 KiB over the link). */
#define FI_EXCHANGE_NONE 0
#define FI_EXCHANGE_PEER 1
#define FI_EXCHANGE_NCCL 2
int fi_epp_comm_exchange(fi_epp* h);

/* Runtime knobs (measurement and tuning; every one has a working default).  Names:
 *   "exchange"     sharded pick reduction: FI_EXCHANGE_PEER | FI_EXCHANGE_NCCL (PEER only if the peers were mapped)
 *   "shard_hash"   sharded hashing: 0 = every rank hashes every prompt (default), 1 = split over the ranks + all-gather of the chains
 *   "device_lru"   1 = per-endpoint LRUs resident in HBM (default when lru_capacity >= max_blocks), 0 = host LRU; before the first Add
 *   "lru_table_slots"  slots per endpoint table of the device LRU (0 = sized by free HBM, 4..32 x lru_capacity)
 *   "feed_slices"  slices of a host-buffer pick's prompt copy, 1..16 (default 8)
 *   "lru_threads"  host worker threads of fi_epp_index_add_chains (takes effect at the next call)
 *   "pipe_partition"  SMs of the chain-walk partition of the pipelined path (default 40; 0 = no partition)
 *   "pipe_hash_ctas", "pipe_match_ctas"  CTAs per SM of the two kernels the pipelined path runs side by side (0 = default:
 *                  partitioned GPU 4 hashing CTAs per SM, otherwise one CTA per request; match_pick as many as fit)
 *                  (fi_epp_pick_submit: batch k+1's block hashing next to batch k's match; 0 = uncapped)
 * FI_ERR_INVALID for an unknown name or a value out of range. */
int fi_epp_set_option(fi_epp* h, const char* name, int64_t value);

int fi_epp_set_profiling(fi_epp* h, int on);
int fi_epp_get_stats(fi_epp* h, fi_epp_stats* out);
int fi_epp_reset_stats(fi_epp* h);

#ifdef __cplusplus
}
#endif
#endif /* FI_EPP_H_ */
