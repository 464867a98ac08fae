"""CPU: the oracle (oracle/epp_oracle.cpp) against a second, independent restatement of SURVEY.md Appendix A
(tests/restate.py: python dicts/sets, python floats, the xxhash wheel) — cfg 1 of BASELINE.json, hole-y
indexes in both match modes, PD, LoRA, the LRU, and the edge cases a reader of the appendix could get wrong.
Both sides are test infrastructure; agreement of three differently-shaped implementations (these two + the
kernels, tests/test_gpu_parity.py) is what stands in for the reference's absent golden vectors."""
import numpy as np
import pytest

from fusioninfer_b200 import _abi as abi
from fusioninfer_b200 import synth
from oracle import epp_oracle as eo
from tests import helpers as H
from tests import restate

P, K, Q, L = H.P, H.K, H.Q, abi.FI_SCORER_LORA


def _both(cfg, wl=None, states=None, ops_iter=None):
    o = eo.Oracle(cfg)
    r = restate.from_config(cfg)
    st = states if states is not None else wl.endpoint_states()
    o.update_endpoints(st)
    r.update_endpoints(st)
    for ops in (ops_iter if ops_iter is not None else wl.index_ops()):
        o.index_apply(ops)
        r.apply(ops)
    return o, r


def _same(a, b):
    assert H.picks_equal(a, b), H.describe_diff(a, b)


@pytest.mark.parametrize("block_bytes", [64, 5])
def test_cfg1_oracle_equals_restatement(block_bytes):
    """BASELINE.json configs[0]: 64 requests x 8 endpoints, 256-token prompts; also with the reference's own
    blockSize: 5 (strategy.go:57)."""
    wl = synth.baseline_workload(1, lru_capacity=300)
    profiles, pd = synth.baseline_profiles(1)
    cfg = H.config_for(wl, profiles=profiles, pd=pd)
    tok, offs = wl.prompts()
    if block_bytes == 5:
        cfg.block_bytes = 5
        cfg.max_blocks = 64
        o = eo.Oracle(cfg)
        r = restate.from_config(cfg)
        st = wl.endpoint_states()
        o.update_endpoints(st)
        r.update_endpoints(st)
        # index: the 5-byte chains of the first 24 prompts on endpoints r % 8
        ch, nb = o.hash_batch(tok, offs, wl.h0)
        trip = [(int(ch[q, i]), q % 8, 1) for q in range(24) for i in range(int(nb[q]) - (q % 5))]
        ops = H.ops_array(trip)
        o.index_apply(ops)
        r.apply(ops)
    else:
        o, r = _both(cfg, wl)
    _same(o.pick_batch(tok, offs, wl.h0), r.pick(tok, offs, wl.h0))


@pytest.mark.parametrize("lpm", [False, True])
def test_holey_index_both_modes(lpm):
    wl = H.small_workload(E=24, R=96, holes=True, lru_capacity=200)
    prof = [{"name": "d", "scorers": [(P, 100), (K, 13), (Q, 7)]}]
    cfg = H.config_for(wl, profiles=prof, match_mode=abi.FI_MATCH_LPM if lpm else abi.FI_MATCH_UPSTREAM)
    o, r = _both(cfg, wl)
    tok, offs = wl.prompts()
    _same(o.pick_batch(tok, offs, wl.h0), r.pick(tok, offs, wl.h0))


@pytest.mark.parametrize("threshold", [0.0, 700.0, 1e9])
def test_pd_profiles_and_threshold(threshold):
    wl = H.small_workload(E=32, R=80, pd=True, holes=True)
    profiles, pd = synth.baseline_profiles(5)
    cfg = H.config_for(wl, profiles=profiles, pd=dict(pd, threshold=threshold))
    o, r = _both(cfg, wl)
    tok, offs = wl.prompts()
    want = r.pick(tok, offs, wl.h0)
    _same(o.pick_batch(tok, offs, wl.h0), want)
    if threshold == 1e9:
        assert (want[:, pd["prefill"]]["endpoint"] == abi.FI_NO_ENDPOINT).all()


def test_lru_add_chain_equals_restatement():
    wl = H.small_workload(E=6, R=64, lru_capacity=40)
    prof = [{"name": "d", "scorers": [(P, 100), (Q, 3)]}]
    cfg = H.config_for(wl, profiles=prof, lru_capacity=40)
    o = eo.Oracle(cfg)
    r = restate.from_config(cfg)
    st = wl.endpoint_states()
    o.update_endpoints(st)
    r.update_endpoints(st)
    for step in range(4):
        tok, offs = wl.prompts(batch=step)
        got, ch = o.pick_batch(tok, offs, wl.h0, want_chains=True)
        _same(got, r.pick(tok, offs, wl.h0))
        for q in range(wl.R):
            e, n = int(got[q, 0]["endpoint"]), int(got[q, 0]["n_blocks"])
            o.index_add_chain(e, ch[q, :n])
            r.add_chain(e, ch[q, :n])


def test_lora_affinity():
    wl = H.small_workload(E=12, R=48)
    prof = [{"name": "d", "scorers": [(L, 60), (P, 40)]}]
    cfg = H.config_for(wl, profiles=prof)
    o, r = _both(cfg, wl)
    ls = np.zeros(12, dtype=abi.lora_dtype())
    rng = np.random.default_rng(3)
    for e in range(12):
        ls[e]["endpoint"] = e
        ls[e]["max_active"] = rng.integers(0, 4)
        ls[e]["n_active"] = rng.integers(0, 3)
        ls[e]["n_waiting"] = rng.integers(0, 3)
        ls[e]["active"][:] = rng.integers(1, 6, size=8)
        ls[e]["waiting"][:] = rng.integers(1, 6, size=8)
    o.update_endpoints_lora(ls)
    r.update_lora(ls)
    tok, offs = wl.prompts()
    ad = rng.integers(0, 7, size=wl.R).astype(np.uint64)
    _same(o.pick_batch(tok, offs, wl.h0, adapters=ad), r.pick(tok, offs, wl.h0, adapters=ad))


# ---- edge vectors (VERDICT r1 "What's weak" 1) ---------------------------------------------------
def _edge_cfg(E, profiles, pd=None, M=4, **kw):
    from fusioninfer_b200 import make_config

    return make_config(num_endpoints=E, block_bytes=64, max_blocks=M, max_batch=16, profiles=profiles, pd=pd, **kw)


def test_edge_prompt_shorter_than_a_block_with_pd():
    """n = 0: no hashes, prefix score 0 everywhere, hit ratio 0 -> (1-0)·len >= threshold decides the prefill."""
    profiles, pd = synth.baseline_profiles(5)
    for thr, runs in ((0.0, True), (40.0, True), (41.0, False)):
        cfg = _edge_cfg(8, profiles, dict(pd, threshold=thr))
        st = H.states_array(8, kv=np.linspace(0, 0.7, 8), queue=np.arange(8),
                            roles=np.array([abi.FI_ROLE_PREFILLER] * 4 + [abi.FI_ROLE_DECODER] * 4, dtype=np.uint32))
        o, r = _both(cfg, states=st, ops_iter=[])
        data, offs = H.pack_prompts([bytes(40), bytes(63)])
        got = o.pick_batch(data, offs, 77)
        _same(got, r.pick(data, offs, 77))
        assert (got["n_blocks"] == 0).all() and (got["match_blocks"] == 0).all()
        assert (got[0, pd["prefill"]]["endpoint"] != abi.FI_NO_ENDPOINT) == runs
        assert got[1, pd["prefill"]]["endpoint"] != abi.FI_NO_ENDPOINT  # 63 bytes >= 41


def test_edge_queue_scorer_with_one_eligible_endpoint():
    """maxQ == minQ (a single candidate after the filter) -> queue score 1.0, not 0/0."""
    prof = [{"name": "d", "role_mask": abi.FI_ROLE_DECODER, "scorers": [(Q, 10)]}]
    cfg = _edge_cfg(4, prof)
    st = H.states_array(4, queue=np.array([9, 3, 5, 7]),
                        roles=np.array([1, 1, abi.FI_ROLE_DECODER, 1], dtype=np.uint32))
    o, r = _both(cfg, states=st, ops_iter=[])
    data, offs = H.pack_prompts([bytes(64)])
    got = o.pick_batch(data, offs, 1)
    _same(got, r.pick(data, offs, 1))
    assert got[0, 0]["endpoint"] == 2 and got[0, 0]["score"] == 10.0


def test_edge_zero_weight_scorer_first_in_profile():
    """weight 0 prefix scorer ahead of kv: matches do not move the total; match_blocks still reports them."""
    prof = [{"name": "d", "scorers": [(P, 0), (K, 10)]}]
    cfg = _edge_cfg(3, prof)
    blob = bytes(range(200))
    ch = restate.chain(blob, 64, 4, 5)
    st = H.states_array(3, kv=np.array([0.5, 0.25, 0.25]))
    ops = H.ops_array([(ch[0], 0, 1), (ch[1], 0, 1), (ch[0], 2, 1)])
    o, r = _both(cfg, states=st, ops_iter=[ops])
    data, offs = H.pack_prompts([blob])
    got = o.pick_batch(data, offs, 5)
    _same(got, r.pick(data, offs, 5))
    # endpoints 1 and 2 tie at 7.5; whichever the rotation picks, its own match count is reported
    e = int(got[0, 0]["endpoint"])
    assert e in (1, 2) and got[0, 0]["score"] == 7.5 and got[0, 0]["match_blocks"] == (1 if e == 2 else 0)
    start = restate.tie_start(3, ch[0], 5, 0, 3)
    assert e == min((1, 2), key=lambda x: (x - start) % 3)


def test_edge_truncation_exactly_at_max_blocks():
    """A prompt of M blocks + a partial one and a prompt of M + 3 blocks both hash exactly M blocks."""
    prof = [{"name": "d", "scorers": [(P, 100)]}]
    cfg = _edge_cfg(2, prof, M=4)
    a = bytes(range(256)) + bytes(10)       # 4 blocks + 10 bytes
    b = bytes(range(256)) + bytes(3 * 64)   # 7 blocks
    c = bytes(range(255))                   # 3 blocks + 63 bytes
    ch = restate.chain(a, 64, 4, 9)
    assert len(ch) == 4 and restate.chain(b, 64, 4, 9) == ch and len(restate.chain(c, 64, 4, 9)) == 3
    ops = H.ops_array([(h, 1, 1) for h in ch])
    o, r = _both(cfg, states=H.states_array(2), ops_iter=[ops])
    data, offs = H.pack_prompts([a, b, c])
    got = o.pick_batch(data, offs, 9)
    _same(got, r.pick(data, offs, 9))
    assert list(got[:, 0]["n_blocks"]) == [4, 4, 3] and list(got[:, 0]["match_blocks"]) == [4, 4, 3]
    assert (got[:, 0]["endpoint"] == 1).all() and (got[:, 0]["score"] == 100.0).all()


def test_tie_rotation_spreads_cold_requests():
    """The reference's default profile (prefix scorer only, strategy.go:51-68) with an empty index: every
    endpoint ties at 0.  The rotation must not send everything to endpoint 0."""
    wl = H.small_workload(E=16, R=128)
    cfg = H.config_for(wl)
    o, r = _both(cfg, states=wl.endpoint_states(), ops_iter=[])
    tok, offs = wl.prompts()
    got = o.pick_batch(tok, offs, wl.h0)
    _same(got, r.pick(tok, offs, wl.h0))
    used = np.unique(got[:, 0]["endpoint"])
    assert len(used) >= 10  # 128 requests over 16 endpoints: nearly all get some
    # requests that share their first block rotate alike
    groups, shared = wl.request_params()
    first = {}
    for q in range(wl.R):
        if shared[q]:
            first.setdefault(int(groups[q]), set()).add(int(got[q, 0]["endpoint"]))
    assert all(len(v) == 1 for v in first.values())


def test_chained_label_filters():
    """Several by-label filters in one profile are ANDed (upstream filter plugins chain): oracle == restatement,
    and a disjoint pair admits nothing."""
    wl = H.small_workload(E=30, R=40)
    prof = [{"name": "a", "role_mask": 1 | 4, "more_filters": [8 | 16, 32], "scorers": [(P, 60), (Q, 40)]},
            {"name": "b", "role_mask": 2, "more_filters": [4], "scorers": [(K, 1)]}]
    cfg = H.config_for(wl, profiles=prof)
    st = wl.endpoint_states()
    rng = np.random.default_rng(9)
    st["role_mask"] = (1 << rng.integers(0, 3, size=30)) | (8 << rng.integers(0, 2, size=30)) | np.where(rng.random(30) < 0.7, 32, 0)
    o, r = _both(cfg, wl, states=st)
    tok, offs = wl.prompts()
    got = o.pick_batch(tok, offs, wl.h0)
    _same(got, r.pick(tok, offs, wl.h0))
    ok = ((st["role_mask"] & 5) != 0) & ((st["role_mask"] & 24) != 0) & ((st["role_mask"] & 32) != 0)
    assert ok.any() and ok[got[:, 0]["endpoint"]].all()
    assert (got[:, 1]["endpoint"] == abi.FI_NO_ENDPOINT).all()  # bits 2 and 4 are never set together here
