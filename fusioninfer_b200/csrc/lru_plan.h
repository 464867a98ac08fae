// lru_plan.h — host-side planning of one batch of indexer.Add calls for the DEVICE-resident LRU
// (lru_kernels.cu).  Pure C++ (unit-tested on the CPU through hostcheck.cpp).
//
// Upstream's PreRequest step is indexer.Add(chain_r, pod_r) for every routed request, in request order
// (capacity lruCapacityPerServer, /root/reference/pkg/router/strategy.go:59,149).
// The device LRU applies a whole SUB-BATCH of those at once, which is exact as long as an endpoint's LRU
// never receives more than `cap_per_endpoint` touches within one sub-batch (then nothing touched in the
// sub-batch can be evicted before its end: an LRU of capacity C always holds the C most recently touched
// distinct keys).  The planner cuts the request stream — in request order — into such sub-batches and lays out,
// per sub-batch, what the kernels need:
//   req_id[k]    index of the k-th kept request in the caller's arrays (ascending)
//   req_ep[k]    its LOCAL endpoint
//   req_n[k]     its block count
//   req_off[k]   exclusive prefix sum of req_n within the sub-batch (position of its first touch)
//   ep_start[e] .. ep_start[e+1]  range of ep_list holding the k's of local endpoint e, ascending
//   inc[e]       touches endpoint e receives in the sub-batch (upper bound of its new log records)
#pragma once
#include <cstdint>
#include <vector>

namespace fi {

struct LruSubBatch {
  uint32_t k_begin = 0, k_end = 0;  // range of the kept-request arrays
  uint64_t touches = 0;             // sum of req_n over the range
};

struct LruPlan {
  std::vector<uint32_t> req_id, req_ep, req_n, req_off;  // [K]   (req_off restarts at 0 in every sub-batch)
  std::vector<uint32_t> ep_list;                         // [K]   k relative to the sub-batch's k_begin
  std::vector<uint32_t> ep_start;                        // [nsub][EL + 1]
  std::vector<uint32_t> inc;                             // [nsub][EL]
  std::vector<LruSubBatch> subs;
};

// endpoints[r] is a GLOBAL endpoint index (or any value outside [ep_begin, ep_begin + EL): skipped, like
// FI_NO_ENDPOINT and requests with no blocks).  A request with more than cap_per_endpoint blocks cannot be
// planned (the caller rejects it first).  cap_touches / cap_requests bound a sub-batch by the size of the
// device scratch arrays.
inline void lru_plan_batch(const uint32_t* endpoints, const uint32_t* nblocks, uint32_t R, uint32_t ep_begin, uint32_t EL,
                           uint32_t cap_per_endpoint, uint64_t cap_touches, uint32_t cap_requests, LruPlan* out) {
  LruPlan& p = *out;
  p.req_id.clear();
  p.req_ep.clear();
  p.req_n.clear();
  p.req_off.clear();
  p.ep_list.clear();
  p.ep_start.clear();
  p.inc.clear();
  p.subs.clear();
  std::vector<uint32_t> acc(EL, 0);
  std::vector<uint32_t> touched;  // endpoints with acc != 0 in the open sub-batch
  LruSubBatch cur;
  auto close = [&]() {
    if (cur.k_end == cur.k_begin) return;
    // bucket the sub-batch's requests by endpoint (counting sort keeps k ascending within an endpoint)
    const size_t s0 = p.ep_start.size();
    p.ep_start.resize(s0 + EL + 1, 0);
    uint32_t* st = p.ep_start.data() + s0;
    for (uint32_t k = cur.k_begin; k < cur.k_end; ++k) st[p.req_ep[k] + 1]++;
    for (uint32_t e = 0; e < EL; ++e) st[e + 1] += st[e];
    const size_t l0 = p.ep_list.size();
    p.ep_list.resize(l0 + (cur.k_end - cur.k_begin));
    std::vector<uint32_t> fill(st, st + EL);
    for (uint32_t k = cur.k_begin; k < cur.k_end; ++k) p.ep_list[l0 + fill[p.req_ep[k]]++] = k - cur.k_begin;
    const size_t i0 = p.inc.size();
    p.inc.resize(i0 + EL, 0);
    for (uint32_t e : touched) {
      p.inc[i0 + e] = acc[e];
      acc[e] = 0;
    }
    touched.clear();
    p.subs.push_back(cur);
    cur.k_begin = cur.k_end;
    cur.touches = 0;
  };
  for (uint32_t r = 0; r < R; ++r) {
    const uint32_t e = endpoints[r] - ep_begin;
    const uint32_t n = nblocks[r];
    if (e >= EL || n == 0) continue;
    if (acc[e] + (uint64_t)n > cap_per_endpoint || cur.touches + n > cap_touches || cur.k_end - cur.k_begin >= cap_requests)
      close();
    if (acc[e] == 0) touched.push_back(e);
    acc[e] += n;
    p.req_id.push_back(r);
    p.req_ep.push_back(e);
    p.req_n.push_back(n);
    p.req_off.push_back((uint32_t)cur.touches);
    cur.touches += n;
    cur.k_end++;
  }
  close();
}

}  // namespace fi
