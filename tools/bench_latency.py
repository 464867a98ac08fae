#!/usr/bin/env python
"""Small-batch latency of the pick call — the regime of the ~100 us micro-batcher SURVEY.md §8(f)2 puts in
front of fi_epp_pick_batch: R in {1, 16, 256, 1024} requests per call, host (pinned) buffers in, picks out,
p50 / p99 over N calls, against the headline pool (1 024 endpoints x 31 250 entries, 4 096-token prompts).
Beside it: the CPU oracle's latency for the same calls (1 thread for R = 1, all usable cores otherwise).

    python tools/bench_latency.py [--calls 1000] [--cfg 3] [--no-oracle]
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
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--calls", type=int, default=1000)
    ap.add_argument("--sizes", default="1,16,256,1024")
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    from fusioninfer_b200 import EndpointPicker, PinnedBuffer, make_config, synth

    sizes = [int(x) for x in args.sizes.split(",")]
    wl = synth.baseline_workload(args.cfg, R=max(sizes) * 8)
    profiles, pd = synth.baseline_profiles(args.cfg)
    slots = 4096
    while slots < 2 * wl.E * wl.lru_capacity:
        slots *= 2
    cfg = make_config(num_endpoints=wl.E, block_bytes=wl.block_bytes, max_blocks=wl.max_blocks, lru_capacity=0,
                      max_batch=max(sizes), max_prompt_bytes=max(sizes) * wl.T * 4, index_slots=slots, profiles=profiles, pd=pd)
    gpu = EndpointPicker(cfg)
    gpu.update_endpoints(wl.endpoint_states())
    cpu = None
    if not args.no_oracle:
        from oracle import epp_oracle as eo

        cpu = eo.Oracle(cfg)
        cpu.update_endpoints(wl.endpoint_states())
        cpu.index_reserve(wl.E * wl.lru_capacity)
        ncores = eo.usable_cores()
    for ops in wl.index_ops(chunk_endpoints=128):
        gpu.index_apply(ops)
        if cpu is not None:
            cpu.index_apply(ops)
    gpu.index_sync()
    tok, offs = wl.prompts()
    P = gpu.n_profiles
    out = {"workload": f"cfg{args.cfg} pool ({wl.E} endpoints x {wl.lru_capacity} entries), {wl.T}-token prompts, host buffers",
           "calls": args.calls, "sizes": {}}
    for R in sizes:
        pin_tok, pin_off, pin_h0, pin_out = PinnedBuffer(R * wl.T * 4), PinnedBuffer(8 * (R + 1)), PinnedBuffer(8 * R), PinnedBuffer(16 * R * P)
        pin_off.array(np.uint64)[:] = offs[: R + 1]
        pin_h0.array(np.uint64)[:] = np.uint64(wl.h0)
        lat = np.zeros(args.calls)
        nvar = wl.R // R
        for i in range(args.calls + 20):
            k = (i % nvar) * R
            pin_tok.array(np.uint32)[:] = tok[k : k + R].reshape(-1)  # fresh prompts every call (outside the timer)
            t0 = time.perf_counter()
            gpu.pick_batch_raw(pin_tok.ptr, pin_off.ptr, pin_h0.ptr, R, pin_out.ptr)
            if i >= 20:
                lat[i - 20] = time.perf_counter() - t0
        rec = {"gpu_us": {"p50": 1e6 * float(np.median(lat)), "p99": 1e6 * float(np.quantile(lat, 0.99)),
                          "mean": 1e6 * float(lat.mean())},
               "gpu_decisions_per_s_at_p50": R / float(np.median(lat))}
        if cpu is not None:
            nt = 1 if R == 1 else min(ncores, R)
            n_cpu = max(20, min(args.calls, 200))
            cl = np.zeros(n_cpu)
            for i in range(n_cpu + 3):
                k = (i % nvar) * R
                sub = np.ascontiguousarray(tok[k : k + R])
                t0 = time.perf_counter()
                cpu.pick_batch(sub, offs[: R + 1], wl.h0, nthreads=nt)
                if i >= 3:
                    cl[i - 3] = time.perf_counter() - t0
            rec["oracle_us"] = {"p50": 1e6 * float(np.median(cl)), "p99": 1e6 * float(np.quantile(cl, 0.99)), "threads": nt}
        out["sizes"][str(R)] = rec
        for b in (pin_tok, pin_off, pin_h0, pin_out):
            b.free()
    print(json.dumps(out), flush=True)
    gpu.close()


if __name__ == "__main__":
    main()
