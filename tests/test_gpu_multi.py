"""GPU (-m gpu, needs >= 2 devices): endpoint-range sharded pool over the library's NCCL
communicator — one handle per GPU driven from one thread each — equals the unsharded
oracle, in both match modes.  Every rank's index is a directory of the whole pool's keys kept
exact by gossiping the owners' appear/vanish transitions, so upstream's stopping point (the
first block NO pod holds) is found locally; holes in the index, CLEARs and LRU churn make
that gossip matter.  Also: split vs replicated hashing, both pick exchanges, and BASELINE.json's
configs 4 and 5 at full size."""
import threading

import numpy as np
import pytest

from fusioninfer_b200 import EndpointPicker, synth
from fusioninfer_b200 import _abi as abi
from fusioninfer_b200.dist import shard_range
from oracle import epp_oracle as eo
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _run_sharded(world, wl, profiles, pd, mode):
    uid = EndpointPicker.comm_unique_id()
    tok, offs = wl.prompts()
    results = [None] * world
    errors = []

    def worker(rank):
        try:
            begin, count = shard_range(wl.E, rank, world)
            cfg = H.config_for(wl, profiles=profiles, pd=pd, match_mode=mode, device=rank, endpoint_begin=begin,
                               endpoint_count=count)
            p = EndpointPicker(cfg)
            p.comm_init(uid, rank, world)
            p.update_endpoints(wl.endpoint_states())
            for ops in wl.index_ops():
                p.index_apply(ops)  # the library keeps only this shard's entries
            results[rank] = p.pick_batch(tok, offs, wl.h0)
            p.close()
        except Exception as e:  # pragma: no cover
            errors.append((rank, repr(e)))

    ths = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=600)
    assert not errors, errors
    return results, tok, offs


@pytest.mark.parametrize("mode", [abi.FI_MATCH_UPSTREAM, abi.FI_MATCH_LPM])
@pytest.mark.parametrize("holes", [False, True])
def test_sharded_pool_equals_unsharded_oracle(gpu_count, mode, holes):
    if gpu_count < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if gpu_count < 4 else 4
    wl = H.small_workload(E=200, R=192, holes=holes, lru_capacity=300)
    profiles = [{"name": "default", "scorers": [(H.P, 100), (H.K, 13), (H.Q, 7)]}]
    results, tok, offs = _run_sharded(world, wl, profiles, None, mode)
    cpu = eo.Oracle(H.config_for(wl, profiles=profiles, match_mode=mode))
    cpu.update_endpoints(wl.endpoint_states())
    for ops in wl.index_ops():
        cpu.index_apply(ops)
    want = cpu.pick_batch(tok, offs, wl.h0)
    for r in range(world):  # every rank ends with the full, identical answer
        assert H.picks_equal(results[r], want), f"rank {r}\n" + H.describe_diff(results[r], want)


def test_sharded_pd_pick(gpu_count):
    if gpu_count < 2:
        pytest.skip("needs >= 2 GPUs")
    wl = H.small_workload(E=128, R=128, pd=True)
    profiles, pd = synth.baseline_profiles(5)
    pd = dict(pd, threshold=700.0)
    results, tok, offs = _run_sharded(2, wl, profiles, pd, abi.FI_MATCH_UPSTREAM)
    cpu = eo.Oracle(H.config_for(wl, profiles=profiles, pd=pd))
    cpu.update_endpoints(wl.endpoint_states())
    for ops in wl.index_ops():
        cpu.index_apply(ops)
    want = cpu.pick_batch(tok, offs, wl.h0)
    assert H.picks_equal(results[0], want), H.describe_diff(results[0], want)
    assert H.picks_equal(results[1], want)


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_sharded_exchange_many_steps(gpu_count, exchange, monkeypatch):
    """Consecutive picks of different batch sizes through both exchange paths: the peer-memory path
    double-buffers its slots by step parity, so at least three back-to-back steps must stay exact."""
    if gpu_count < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("FI_EPP_EXCHANGE", exchange)
    world = 2
    wl = H.small_workload(E=160, R=256, holes=True, lru_capacity=300, pd=True)
    profiles, pd = synth.baseline_profiles(5)
    pd = dict(pd, threshold=700.0)
    batches = []
    for b, R in enumerate([256, 64, 1, 255, 128, 256]):
        tok, offs = wl.prompts(batch=b)
        batches.append((np.ascontiguousarray(tok[:R]), offs[: R + 1].copy(), wl.h0))
    uid = EndpointPicker.comm_unique_id()
    results = [[] for _ in range(world)]
    modes = [None] * world
    errors = []

    def worker(rank):
        try:
            begin, count = shard_range(wl.E, rank, world)
            cfg = H.config_for(wl, profiles=profiles, pd=pd, device=rank, endpoint_begin=begin, endpoint_count=count)
            p = EndpointPicker(cfg)
            p.comm_init(uid, rank, world)
            modes[rank] = p.comm_exchange()
            p.update_endpoints(wl.endpoint_states())
            for ops in wl.index_ops():
                p.index_apply(ops)
            for tok, offs, h0 in batches:
                results[rank].append(p.pick_batch(tok, offs, h0))
            p.close()
        except Exception as e:  # pragma: no cover
            errors.append((rank, repr(e)))

    ths = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=600)
    assert not errors, errors
    assert modes == [exchange] * world
    cpu = eo.Oracle(H.config_for(wl, profiles=profiles, pd=pd))
    cpu.update_endpoints(wl.endpoint_states())
    for ops in wl.index_ops():
        cpu.index_apply(ops)
    for i, (tok, offs, h0) in enumerate(batches):
        want = cpu.pick_batch(tok, offs, h0)
        for r in range(world):
            assert H.picks_equal(results[r][i], want), f"step {i} rank {r}\n" + H.describe_diff(results[r][i], want)


def _threads(world, fn):
    errors = []

    def run(rank):
        try:
            fn(rank)
        except Exception:  # pragma: no cover
            import traceback

            errors.append((rank, traceback.format_exc()))

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=1800)
    assert not errors, "\n".join(f"rank {r}: {m}" for r, m in errors)


@pytest.mark.parametrize("hashing", ["split", "replicated"])
def test_sharded_churn_clears_and_add_chains(gpu_count, hashing, monkeypatch):
    """The directory under change: SET/CLEAR streams that make keys vanish from one rank while another still
    holds them (and vanish everywhere), then several steps of pick + collective fi_epp_index_add_chains with
    LRU evictions — every rank must keep returning the unsharded oracle's picks."""
    if gpu_count < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("FI_EPP_SHARD_HASH", hashing)
    world = 2 if gpu_count < 4 else 4
    wl = H.small_workload(E=40, R=160, T=768, max_blocks=48, lru_capacity=90, holes=True)
    prof = [{"name": "d", "scorers": [(H.P, 100), (H.K, 9), (H.Q, 5)]}]
    base_ops = np.concatenate(list(wl.index_ops()))
    rng = np.random.default_rng(5)
    # clears: a random third of the pairs, plus ALL pairs of a few hashes (those keys vanish from the pool)
    clr = base_ops[rng.random(len(base_ops)) < 0.33].copy()
    gone = np.unique(base_ops["hash"])[::17]
    clr_all = base_ops[np.isin(base_ops["hash"], gone)].copy()
    clears = np.concatenate([clr, clr_all])
    clears["op"] = abi.FI_OP_CLEAR
    readd = clr_all[::3].copy()  # some of the vanished keys come back on a subset of their endpoints
    steps = 4
    uid = EndpointPicker.comm_unique_id()
    results = [[] for _ in range(world)]
    cfg_full = H.config_for(wl, profiles=prof, lru_capacity=90, index_slots=1 << 17)

    # the oracle runs first: its picks decide which chains are added (every rank adds the same)
    cpu = eo.Oracle(cfg_full)
    cpu.update_endpoints(wl.endpoint_states())
    want = []
    for ops in (base_ops, clears, readd):
        cpu.index_apply(ops)
    batches = [wl.prompts(batch=b % 2) for b in range(steps)]
    for tok, offs in batches:
        w, ch = cpu.pick_batch(tok, offs, wl.h0, want_chains=True)
        want.append(w)
        cpu.index_add_chains(w[:, 0]["endpoint"], ch, w[:, 0]["n_blocks"])

    def worker(rank):
        begin, count = shard_range(wl.E, rank, world)
        cfg = H.config_for(wl, profiles=prof, lru_capacity=90, index_slots=1 << 17, device=rank, endpoint_begin=begin,
                           endpoint_count=count)
        p = EndpointPicker(cfg)
        p.comm_init(uid, rank, world)
        p.update_endpoints(wl.endpoint_states())
        for ops in (base_ops, clears, readd):
            p.index_apply(ops)  # collective: same call sequence on every rank; the library keeps its shard
        for tok, offs in batches:
            got, ch = p.pick_batch(tok, offs, wl.h0, want_chains=True)
            results[rank].append(got)
            p.index_add_chains(got[:, 0]["endpoint"], ch, got[:, 0]["n_blocks"])
        p.close()

    _threads(world, worker)
    for i in range(steps):
        for r in range(world):
            assert H.picks_equal(results[r][i], want[i]), f"step {i} rank {r}\n" + H.describe_diff(results[r][i], want[i])


@pytest.mark.parametrize("cfg_id", [4, 5])
def test_full_size_sharded_configs(gpu_count, cfg_id):
    """BASELINE.json configs[3] (65 536 req x 4 096 endpoints) and configs[4] (PD, 512 P + 512 D, kv/queue
    weighted) at FULL size over >= 2 GPUs: sampled bit-exact parity with the unsharded oracle (which sees only
    the index entries the sampled requests can touch — for them exactly equivalent), every rank identical."""
    if gpu_count < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if gpu_count >= 8 else 4 if gpu_count >= 4 else 2
    wl = synth.baseline_workload(cfg_id)
    profiles, pd = synth.baseline_profiles(cfg_id)
    if pd is not None:
        pd = dict(pd, threshold=6000.0)  # a non-trivial threshold: prefills behind a > 63 % decode hit are skipped
    slots = 4096
    while slots < 2 * wl.E * wl.lru_capacity:  # the directory holds the whole pool's keys
        slots *= 2
    tok, offs = wl.prompts()
    chunks = list(wl.index_ops(chunk_endpoints=128))
    uid = EndpointPicker.comm_unique_id()
    results = [None] * world
    chains_out = [None]

    def worker(rank):
        begin, count = shard_range(wl.E, rank, world)
        cfg = H.config_for(wl, profiles=profiles, pd=pd, index_slots=slots, max_prompt_bytes=wl.R * wl.T * 4, device=rank,
                           endpoint_begin=begin, endpoint_count=count)
        p = EndpointPicker(cfg)
        p.comm_init(uid, rank, world)
        p.update_endpoints(wl.endpoint_states())
        for ops in chunks:
            p.index_apply(ops[(ops["endpoint"] >= begin) & (ops["endpoint"] < begin + count)])
        if rank == 0:
            results[rank], chains_out[0] = p.pick_batch(tok, offs, wl.h0, want_chains=True)
        else:
            results[rank] = p.pick_batch(tok, offs, wl.h0)
        p.close()

    _threads(world, worker)
    for r in range(1, world):
        assert H.picks_equal(results[r], results[0]), f"rank {r} differs from rank 0"
    picks, chains = results[0], chains_out[0]
    mb = picks[:, -1]["match_blocks"].astype(np.int64)
    # (with cfg 5's kv / queue weights a matching endpoint does not always win: ~28 % of the picks carry a match)
    assert (picks["n_blocks"] == wl.n_blocks).all() and (mb > 0).mean() > (0.2 if pd is not None else 0.5)
    S = 768
    idx = np.linspace(0, wl.R - 1, S).astype(np.int64)
    needed = np.unique(chains[idx])
    cpu = eo.Oracle(H.config_for(wl, profiles=profiles, pd=pd))
    cpu.update_endpoints(wl.endpoint_states())
    for ops in chunks:
        cpu.index_apply(ops[np.isin(ops["hash"], needed)])
    # the oracle's tie rotation of a short prompt depends on the request's index in the call; these prompts all
    # have blocks, so sampling does not change it
    sub_offs = np.arange(S + 1, dtype=np.uint64) * np.uint64(wl.T * 4)
    want = cpu.pick_batch(np.ascontiguousarray(tok[idx]), sub_offs, wl.h0, nthreads=8)
    assert H.picks_equal(picks[idx], want), H.describe_diff(picks[idx], want)
