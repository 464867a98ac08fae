"""CPU: the synthetic workload generator is deterministic and has the stated shape."""
import numpy as np

from fusioninfer_b200 import _abi as abi
from fusioninfer_b200 import synth
from tests import helpers as H


def test_splitmix_reference_values():
    # SplitMix64 with state 0 → first output 0xE220A8397B1DCDAF (published test vector)
    assert int(synth.sm64(np.uint64(0))) == 0xE220A8397B1DCDAF


def test_prompts_deterministic_and_structured():
    wl = H.small_workload(R=200, E=16)
    a, offs = wl.prompts()
    b, _ = wl.prompts()
    assert np.array_equal(a, b) and a.dtype == np.uint32 and a.shape == (200, wl.T)
    assert a.max() < synth.VOCAB
    assert list(offs[:3]) == [0, wl.T * 4, 2 * wl.T * 4]
    c, _ = wl.prompts(batch=1)
    assert not np.array_equal(a, c)
    groups, shared = wl.request_params()
    assert 0.03 < (shared == 0).mean() < 0.25  # ~10 % fully unique
    nz = shared[shared > 0]
    assert nz.min() >= wl.T // 4 and nz.max() <= 3 * wl.T // 4 and (nz % wl.block_tokens == 0).all()
    # two requests of one group share exactly their common prefix
    base = wl.group_base(groups[:1])[0]
    if shared[0]:
        assert np.array_equal(a[0, : shared[0]], base[: shared[0]])
        assert not np.array_equal(a[0, shared[0]: shared[0] + 16], base[shared[0]: shared[0] + 16])


def test_index_ops_shape_and_independent_hashes():
    wl = H.small_workload(R=8, E=12, lru_capacity=300)
    ops = np.concatenate(list(wl.index_ops()))
    assert ops.dtype == H.OP_DTYPE and (ops["op"] == abi.FI_OP_SET).all()
    per_ep = np.bincount(ops["endpoint"], minlength=wl.E)
    assert (per_ep == wl.lru_capacity).all()
    # the group chains come from python xxhash over the group's tokens
    eg = wl.endpoint_groups()
    g = int(eg[0, 0])
    want = synth.chain_py(wl.group_base(np.array([g]))[0].tobytes(), wl.block_bytes, wl.max_blocks, wl.h0)
    mine = ops[ops["endpoint"] == 0]["hash"][: len(want)]
    assert np.array_equal(mine, want)
    holes = H.small_workload(R=8, E=12, lru_capacity=300, holes=True)
    ops_h = np.concatenate(list(holes.index_ops()))
    assert 0.90 < len(ops_h) / len(ops) < 0.99


def test_endpoint_states_and_baseline_configs():
    wl = synth.baseline_workload(5, R=8)
    st = wl.endpoint_states()
    assert (st["role_mask"][: wl.E // 2] == abi.FI_ROLE_PREFILLER).all() and (st["role_mask"][wl.E // 2:] == abi.FI_ROLE_DECODER).all()
    assert ((st["kv_util"] * 1024) % 1 == 0).all() and st["queue_depth"].max() < 32
    for cfg, (R, E, T) in {1: (64, 8, 256), 2: (4096, 256, 2048), 3: (16384, 1024, 4096), 4: (65536, 4096, 4096)}.items():
        w = synth.baseline_workload(cfg)
        assert (w.R, w.E, w.T) == (R, E, T) and w.block_bytes == 64 and w.seed == synth.SEEDS[cfg]
    prof, pd = synth.baseline_profiles(5)
    assert pd == {"decode": 1, "prefill": 0, "threshold": 0.0} and len(prof) == 2


def test_config1_cpu_plumbing_case_runs_on_the_oracle():
    """BASELINE.json configs[0]: 64 reqs × 8 endpoints, 256-tok prompts — the reference-side CPU scorer."""
    from oracle import epp_oracle as eo

    wl = synth.baseline_workload(1, lru_capacity=400)
    cfg = H.config_for(wl)
    o = eo.Oracle(cfg)
    o.update_endpoints(wl.endpoint_states())
    for ops in wl.index_ops():
        o.index_apply(ops)
    tok, offs = wl.prompts()
    pk = o.pick_batch(tok, offs, wl.h0)
    assert pk.shape == (64, 1) and (pk["n_blocks"] == 16).all()
    assert (pk["match_blocks"] > 0).sum() > 20
    groups, shared = wl.request_params()
    hit = pk[:, 0]["match_blocks"] > 0
    # a hit can never exceed the shared prefix of the request
    assert (pk[:, 0]["match_blocks"][hit] <= (shared[hit] // wl.block_tokens)).all()
