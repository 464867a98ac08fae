"""GPU (-m gpu, needs >= 2 devices): endpoint-range sharded pool over the library's NCCL
communicator — one handle per GPU driven from one thread each — equals the unsharded
oracle, in both match modes (the upstream mode exchanges presence masks to find the global
first miss; holes in the index make that exchange matter)."""
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
