#!/usr/bin/env python
"""Index-maintenance path under churn: every pick batch is followed by upstream's
PreRequest step for the whole batch — indexer.Add(chain_r, picked endpoint_r).  Two implementations:
  --lru device (default)  the GPU-resident LRU (lru_kernels.cu): the chains never leave the device
                          (fi_epp_index_add_chains_device with the handle's own chain buffer), only the picked
                          endpoints come back from the host;
  --lru host              fi_epp_index_add_chains on the host LRU (worker pool; SET/CLEAR deltas streamed to the
                          GPU index on the side stream, ordered before the next pick).
Picks are checked bit-exactly against the oracle doing the same work (sequentially, one chain at a time), whose
time is reported beside ours.

    python tools/bench_churn.py [--cfg 3] [--requests N] [--steps 12] [--lru host --threads T] [--no-oracle]

Defaults = BASELINE.json's headline pool: 1 024 endpoints, lruCapacityPerServer 31 250
(/root/reference/pkg/router/strategy.go:59), 16 384 requests of 4 096 tokens per step.  The index starts
full (every endpoint's LRU at capacity), so every new block evicts one.  Prints one JSON line:
decisions/s of pick + add, the split, ops streamed, tombstone growth, rebuilds and the pick-latency
distribution (p50 / p99 / max over the steps — a rebuild shows up there).
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
    ap.add_argument("--requests", type=int, default=0)
    ap.add_argument("--endpoints", type=int, default=0)
    ap.add_argument("--lru-capacity", dest="lru", type=int, default=0)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--lru", dest="lru_impl", choices=["device", "host"], default="device")
    ap.add_argument("--oracle-steps", type=int, default=2, help="steps the oracle mirrors (it is ~100x slower)")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--slot-mult", type=int, default=2)
    args = ap.parse_args()

    from fusioninfer_b200 import EndpointPicker, PinnedBuffer, make_config, synth

    over = {}
    if args.requests:
        over["R"] = args.requests
    if args.endpoints:
        over["E"] = args.endpoints
    if args.lru:
        over["lru_capacity"] = args.lru
    wl = synth.baseline_workload(args.cfg, **over)
    profiles, pd = synth.baseline_profiles(args.cfg)
    slots = 4096
    while slots < args.slot_mult * wl.E * wl.lru_capacity:
        slots *= 2
    cfg = make_config(num_endpoints=wl.E, block_bytes=wl.block_bytes, max_blocks=wl.max_blocks, lru_capacity=wl.lru_capacity,
                      max_batch=max(wl.R, 8192), max_prompt_bytes=max(wl.R, 8192) * wl.T * 4, index_slots=slots,
                      profiles=profiles, pd=pd)
    gpu = EndpointPicker(cfg)
    device_lru = args.lru_impl == "device"
    gpu.set_option("device_lru", 1 if device_lru else 0)
    if args.threads:
        gpu.set_option("lru_threads", args.threads)
    gpu.update_endpoints(wl.endpoint_states())
    cpu = None
    if not args.no_oracle:
        from oracle import epp_oracle as eo

        cpu = eo.Oracle(cfg)
        cpu.update_endpoints(wl.endpoint_states())
        cpu.index_reserve(2 * wl.E * wl.lru_capacity)

    # ---- initial state THROUGH the LRUs (filler first, then the shared group chains: the filler is the oldest)
    t0 = time.time()
    nb = wl.n_blocks
    for ops in wl.index_ops(chunk_endpoints=64):
        e, h = ops["endpoint"], ops["hash"]
        rows, eps = [], []
        for ep in np.unique(e):
            he = h[e == ep]
            seq = np.concatenate([he[wl.groups_per_endpoint * nb:], he[: wl.groups_per_endpoint * nb]])
            seq = np.concatenate([seq, np.zeros((-len(seq)) % nb, dtype=np.uint64)])
            rows.append(seq.reshape(-1, nb))
            eps.append(np.full(rows[-1].shape[0], ep, dtype=np.uint32))
        ch, ee = np.concatenate(rows), np.concatenate(eps)
        valid = (ch != 0).sum(axis=1).astype(np.uint32)
        gpu.index_add_chains(ee, ch, valid)
        if cpu is not None:
            cpu.index_add_chains(ee, ch, valid)
    gpu.index_sync()
    st0 = gpu.index_stats()
    print(f"[churn] initial state: {st0.lru_entries} LRU entries, {st0.used} keys ({time.time() - t0:.1f}s)", file=sys.stderr)

    R, P = wl.R, gpu.n_profiles
    main_p = cfg.pd_decode_profile if cfg.pd_enabled else 0
    pin_tok, pin_off, pin_h0 = PinnedBuffer(R * wl.T * 4), PinnedBuffer(8 * (R + 1)), PinnedBuffer(8 * R)
    pin_out, pin_ch = PinnedBuffer(16 * R * P), PinnedBuffer(8 * R * wl.max_blocks)
    pin_h0.array(np.uint64)[:] = np.uint64(wl.h0)
    picks_v = pin_out.array(np.uint8).view(np.dtype([("endpoint", "<u4"), ("match_blocks", "<u2"), ("n_blocks", "<u2"),
                                                      ("score", "<f8")])).reshape(R, P)
    chains_v = pin_ch.array(np.uint64).reshape(R, wl.max_blocks)
    t_pick, t_add, exact, hits = [], [], True, 0
    t_cpu_pick = t_cpu_add = 0.0
    tomb = []

    def add_step():
        ends = np.ascontiguousarray(picks_v[:, main_p]["endpoint"])
        nbl = np.ascontiguousarray(picks_v[:, main_p]["n_blocks"]).astype(np.uint32)
        if device_lru:
            gpu.index_add_chains_device(ends, 0, 0, nbl)  # the chains of the pick just made, from the handle's buffer
        else:
            gpu.index_add_chains(ends, chains_v, nbl)

    for step in range(args.steps):
        tok, offs = wl.prompts(batch=step)
        pin_tok.array(np.uint32)[:] = tok.reshape(-1)
        pin_off.array(np.uint64)[:] = offs
        t0 = time.perf_counter()
        # (the device LRU's kernels are asynchronous: the NEXT pick waits for them on the GPU, so their time
        # shows up in pick_ms; decisions/s is over the wall clock of both calls, with a final sync)
        gpu.pick_batch_raw(pin_tok.ptr, pin_off.ptr, pin_h0.ptr, R, pin_out.ptr, 0 if device_lru else pin_ch.ptr)
        t1 = time.perf_counter()
        add_step()
        if step == args.steps - 1:
            gpu.index_sync()
        t2 = time.perf_counter()
        t_pick.append(t1 - t0)
        t_add.append(t2 - t1)
        hits += int((picks_v[:, main_p]["match_blocks"] > 0).sum())
        if cpu is not None and step < args.oracle_steps:
            c0 = time.perf_counter()
            want, wch = cpu.pick_batch(tok, offs, wl.h0, want_chains=True, nthreads=os.cpu_count() or 1)
            c1 = time.perf_counter()
            exact = exact and picks_v.tobytes() == want.tobytes()
            cpu.index_add_chains(want[:, main_p]["endpoint"], wch, want[:, main_p]["n_blocks"])
            t_cpu_pick += c1 - c0
            t_cpu_add += time.perf_counter() - c1
        if step % 3 == 2 or step == args.steps - 1:
            s = gpu.index_stats()  # (synchronises the index stream: kept out of the timed calls)
            tomb.append({"step": step, "used": int(s.used), "tombstones": int(s.tombstones), "rebuilds": int(s.rebuilds)})
    gpu.index_sync()
    st = gpu.index_stats()
    # device time of the index kernels for one more step (profiled: CUDA events around every launch)
    gpu.reset_stats()
    gpu.set_profiling(True)
    gpu.pick_batch_raw(pin_tok.ptr, pin_off.ptr, pin_h0.ptr, R, pin_out.ptr, 0 if device_lru else pin_ch.ptr)
    add_step()
    gpu.index_sync()
    pst = gpu.stats()
    gpu.set_profiling(False)
    tp, ta = np.array(t_pick[1:]), np.array(t_add[1:])  # step 0 warms the worker pool / first-touch pages
    n = R * len(tp)
    out = {
        "mode": "churn: pick + indexer.Add(chain, picked endpoint) for every decision, LRUs at capacity",
        "workload": f"cfg{args.cfg}: {R} req/step x {wl.E} endpoints x {wl.T}-token prompts, lruCapacityPerServer {wl.lru_capacity}",
        "steps_timed": len(tp), "lru": args.lru_impl,
        "lru_threads": None if device_lru else (args.threads or "default (usable cores, <= 128)"),
        "decisions_per_s": n / float(tp.sum() + ta.sum()),
        "pick_ms": {"p50": 1e3 * float(np.median(tp)), "p99": 1e3 * float(np.quantile(tp, 0.99)), "max": 1e3 * float(tp.max())},
        "add_ms": {"p50": 1e3 * float(np.median(ta)), "p99": 1e3 * float(np.quantile(ta, 0.99)), "max": 1e3 * float(ta.max())},
        "lru_touches_per_s": n * wl.n_blocks / float(ta.sum()),
        "index_kernels_ms_per_step": pst.ms_index_apply, "index_kernel_launches_per_step": int(pst.n_index_apply),
        "index": {"slots": int(st.slots), "used": int(st.used), "tombstones": int(st.tombstones), "rebuilds": int(st.rebuilds),
                  "ops_applied": int(st.ops_applied), "lru_entries": int(st.lru_entries), "growth": tomb},
        "requests_with_prefix_hit": hits,
        "oracle": None if cpu is None else {
            "steps_mirrored": min(args.oracle_steps, args.steps), "bit_exact": bool(exact),
            "pick_ms_per_step": 1e3 * t_cpu_pick / max(min(args.oracle_steps, args.steps), 1),
            "add_ms_per_step": 1e3 * t_cpu_add / max(min(args.oracle_steps, args.steps), 1),
            "decisions_per_s": R * min(args.oracle_steps, args.steps) / max(t_cpu_pick + t_cpu_add, 1e-9),
            "how": "oracle: multi-threaded pick, then indexer.Add one chain at a time on one thread (std::list + unordered_map LRU)"},
    }
    print(json.dumps(out), flush=True)
    gpu.close()


if __name__ == "__main__":
    main()
