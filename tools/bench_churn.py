#!/usr/bin/env python
"""Index-maintenance path under churn: every decision is followed by
upstream's PreRequest step, indexer.Add(chain, picked endpoint), through the host LRU
(fi_epp_index_add_chain) whose SET/CLEAR deltas stream to the GPU index on the side stream and are
ordered before the next pick.  Checks the picks bit-exactly against the oracle doing the same.

    python tools/bench_churn.py [--requests 2048] [--endpoints 256] [--steps 6] [--lru 4000]

Prints one JSON line: decisions/s of pick + add, the split between the two, ops streamed, rebuilds.
This is NOT the headline metric (bench.py): the exact per-endpoint LRU order is pointer chasing on the
host (one list + map touch per block), which is what bounds this mode.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=2048)
    ap.add_argument("--endpoints", type=int, default=256)
    ap.add_argument("--tokens", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--lru", type=int, default=4000)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()

    from fusioninfer_b200 import EndpointPicker, make_config, synth
    from fusioninfer_b200 import _abi as abi

    wl = synth.Workload(R=args.requests, E=args.endpoints, T=args.tokens, seed=synth.SEEDS[2], lru_capacity=args.lru)
    P, K, Q = abi.FI_SCORER_PREFIX, abi.FI_SCORER_KV_UTIL, abi.FI_SCORER_QUEUE
    slots = 4096
    while slots < 4 * wl.E * args.lru:
        slots *= 2
    cfg = make_config(num_endpoints=wl.E, block_bytes=wl.block_bytes, max_blocks=wl.max_blocks, lru_capacity=args.lru,
                      max_batch=wl.R, max_prompt_bytes=wl.R * wl.T * 4, index_slots=slots,
                      profiles=[{"name": "default", "scorers": [(P, 100), (K, 10), (Q, 10)]}])
    gpu = EndpointPicker(cfg)
    gpu.update_endpoints(wl.endpoint_states())
    cpu = None
    if not args.no_oracle:
        from oracle.epp_oracle import Oracle

        cpu = Oracle(cfg)
        cpu.update_endpoints(wl.endpoint_states())

    t_pick = t_add = 0.0
    exact = True
    hits = 0
    for step in range(args.steps):
        tok, offs = wl.prompts(batch=step)
        t0 = time.perf_counter()
        picks, chains = gpu.pick_batch(tok, offs, wl.h0, want_chains=True)
        t1 = time.perf_counter()
        for r in range(wl.R):
            gpu.index_add_chain(int(picks[r, 0]["endpoint"]), chains[r, : int(picks[r, 0]["n_blocks"])])
        gpu.index_sync()
        t2 = time.perf_counter()
        if step:  # step 0 fills an empty index
            t_pick += t1 - t0
            t_add += t2 - t1
        hits += int((picks["match_blocks"] > 0).sum())
        if cpu is not None:
            want, wch = cpu.pick_batch(tok, offs, wl.h0, want_chains=True, nthreads=os.cpu_count() or 1)
            exact = exact and picks.tobytes() == want.tobytes()
            for r in range(wl.R):
                cpu.index_add_chain(int(want[r, 0]["endpoint"]), wch[r, : int(want[r, 0]["n_blocks"])])
    st = gpu.index_stats()
    n = wl.R * (args.steps - 1)
    print(json.dumps({
        "mode": "churn: pick + indexer.Add(chain, picked endpoint) per decision",
        "requests_per_step": wl.R, "endpoints": wl.E, "prompt_tokens": wl.T, "lru_capacity": args.lru,
        "steps_timed": args.steps - 1,
        "decisions_per_s": n / (t_pick + t_add), "pick_ms_per_step": 1e3 * t_pick / (args.steps - 1),
        "add_ms_per_step": 1e3 * t_add / (args.steps - 1),
        "lru_touches_per_s": n * wl.n_blocks / t_add,
        "index": {"slots": st.slots, "used": st.used, "tombstones": st.tombstones, "rebuilds": st.rebuilds,
                  "ops_applied": st.ops_applied, "lru_entries": st.lru_entries},
        "requests_with_prefix_hit": hits, "bit_exact_vs_oracle": (exact if cpu is not None else None),
    }), flush=True)
    gpu.close()


if __name__ == "__main__":
    main()
