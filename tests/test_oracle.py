"""CPU: pin the oracle (test infrastructure) before trusting it.

* XXH64 against the published known answers, the committed golden vectors and python
  `xxhash` 3.7.0 on random inputs.
* the hash chain (Appendix A.1) against the golden vectors.
* match / score / pick / PD / LRU semantics (Appendix A.2-A.6) on hand-built cases whose
  answers are worked out in the test.  These are "parity unpinned by reference tests":
  the reference (/root/reference/pkg/router/strategy_test.go) pins YAML substrings only.
"""
import random
import struct

import numpy as np
import pytest

from fusioninfer_b200 import _abi as abi
from fusioninfer_b200 import make_config
from oracle import epp_oracle as eo
from tests import helpers as H

P, K, Q = H.P, H.K, H.Q


def test_xxh64_known_answers():
    g = H.golden()
    for s, d in g["known_answers"].items():
        assert f"{eo.xxh64(s.encode()):016x}" == d


def test_xxh64_golden_vectors():
    for case in H.golden()["xxh64"]:
        assert f"{eo.xxh64(bytes.fromhex(case['hex'])):016x}" == case["digest"]


def test_xxh64_random_vs_python_xxhash():
    xxhash = pytest.importorskip("xxhash")
    rng = random.Random(7)
    for _ in range(400):
        n = rng.randrange(0, 300)
        data = bytes(rng.getrandbits(8) for _ in range(n))
        seed = rng.getrandbits(64)
        assert eo.xxh64(data, seed) == xxhash.xxh64_intdigest(data, seed)


def _oracle(E=4, B=64, M=16, **kw):
    cfg = make_config(num_endpoints=E, block_bytes=B, max_blocks=M, max_batch=64, **kw)
    return eo.Oracle(cfg), cfg


def test_chain_golden_vectors():
    g = H.golden()
    h0 = int(g["h0"], 16)
    for case in g["chains"]:
        o, _ = _oracle(B=case["block_bytes"], M=case["max_blocks"])
        data, offs = H.pack_prompts([bytes.fromhex(case["hex"])])
        chains, nb = o.hash_batch(data, offs, h0)
        want = [int(x, 16) for x in case["chain"]]
        assert nb[0] == len(want)
        assert list(chains[0, : nb[0]]) == want
        assert not chains[0, nb[0]:].any()


def _chain(blob, B=64, M=16, h0=1234):
    o, _ = _oracle(B=B, M=M)
    data, offs = H.pack_prompts([blob])
    c, nb = o.hash_batch(data, offs, h0)
    return [int(x) for x in c[0, : nb[0]]]


def test_match_upstream_vs_lpm_with_holes():
    # endpoint 0 holds blocks 0,1,2,3; endpoint 1 holds 0 and 2,3 (hole at 1); endpoint 2 holds 1 only
    blob = bytes(range(256)) + bytes(64)  # 5 blocks
    ch = _chain(blob)
    ops = [(ch[i], 0, 1) for i in range(4)] + [(ch[0], 1, 1), (ch[2], 1, 1), (ch[3], 1, 1), (ch[1], 2, 1)]
    data, offs = H.pack_prompts([blob])
    for mode, want in ((abi.FI_MATCH_UPSTREAM, {0: 4, 1: 3, 2: 1}), (abi.FI_MATCH_LPM, {0: 4, 1: 1, 2: 0})):
        for target in (0, 1, 2):
            # make `target` the only alive endpoint so its match length is what is reported
            o, _ = _oracle(match_mode=mode)
            alive = np.zeros(4, dtype=np.uint32)
            alive[target] = 1
            o.update_endpoints(H.states_array(4, alive=alive))
            o.index_apply(H.ops_array(ops))
            pk = o.pick_batch(data, offs, 1234)
            assert pk[0, 0]["endpoint"] == target
            assert pk[0, 0]["match_blocks"] == want[target], (mode, target)
            assert pk[0, 0]["n_blocks"] == 5


def test_upstream_stops_at_first_global_miss():
    blob = bytes(range(256))  # 4 blocks
    ch = _chain(blob)
    # nobody holds block 1: everything after it is ignored even though endpoint 0 holds 2,3
    ops = [(ch[0], 0, 1), (ch[2], 0, 1), (ch[3], 0, 1)]
    data, offs = H.pack_prompts([blob])
    o, _ = _oracle()
    o.update_endpoints(H.states_array(4))
    o.index_apply(H.ops_array(ops))
    pk = o.pick_batch(data, offs, 1234)[0, 0]
    assert (pk["endpoint"], pk["match_blocks"]) == (0, 1)
    assert pk["score"] == 100.0 * (1.0 / 4.0)


def test_tie_break_rotation_and_dead_endpoints():
    """Equal totals: the alive endpoint nearest after the request's rotation start wins (fi_epp.h "Ties";
    upstream shuffles).  The start is worked out here from the rule, not taken from the oracle."""
    from tests import restate

    o, _ = _oracle(E=6)
    st = H.states_array(6, alive=np.array([0, 0, 1, 1, 1, 0], dtype=np.uint32))
    o.update_endpoints(st)
    seen = set()
    for k in range(12):
        blob = bytes([k]) * 128
        data, offs = H.pack_prompts([blob])
        pk = o.pick_batch(data, offs, 5)[0, 0]
        start = restate.tie_start(2, _chain(blob, h0=5)[0], 5, 0, 6)
        want = min((2, 3, 4), key=lambda e: (e - start) % 6)
        assert pk["endpoint"] == want and pk["score"] == 0.0
        seen.add(int(pk["endpoint"]))
    assert len(seen) > 1  # different prompts rotate differently
    # a prompt shorter than one block rotates by (h0, request index)
    data, offs = H.pack_prompts([bytes(10), bytes(10)])
    pk = o.pick_batch(data, offs, 5)
    for r in range(2):
        start = restate.tie_start(0, 0, 5, r, 6)
        assert pk[r, 0]["endpoint"] == min((2, 3, 4), key=lambda e: (e - start) % 6)
    blob = bytes(128)
    data, offs = H.pack_prompts([blob])
    o2, _ = _oracle(E=3)
    o2.update_endpoints(H.states_array(3, alive=np.zeros(3, dtype=np.uint32)))
    pk = o2.pick_batch(data, offs, 5)[0, 0]
    assert pk["endpoint"] == abi.FI_NO_ENDPOINT and pk["match_blocks"] == 0 and pk["score"] == 0.0


def test_weighted_score_formula_and_order():
    # prefix 50, kv 25, queue 25 — total in profile order, fp64
    prof = [{"name": "w", "scorers": [(P, 50), (K, 25), (Q, 25)]}]
    blob = bytes(range(192))  # 3 blocks
    ch = _chain(blob)
    o, _ = _oracle(E=3, profiles=prof)
    kv = np.array([0.25, 0.5, 0.125])
    q = np.array([4, 0, 8], dtype=np.int32)
    o.update_endpoints(H.states_array(3, kv=kv, queue=q))
    o.index_apply(H.ops_array([(ch[0], 1, 1), (ch[1], 1, 1), (ch[0], 2, 1)]))
    data, offs = H.pack_prompts([blob])
    pk = o.pick_batch(data, offs, 1234)[0, 0]

    def total(m, kvu, qd):
        t = 0.0
        t = t + (m / 3.0) * 50.0
        t = t + (1.0 - kvu) * 25.0
        t = t + ((8 - qd) / (8 - 0)) * 25.0
        return t

    totals = [total(0, 0.25, 4), total(2, 0.5, 0), total(1, 0.125, 8)]
    best = int(np.argmax(totals))
    assert pk["endpoint"] == best and pk["score"] == totals[best] and pk["match_blocks"] == [0, 2, 1][best]


def test_queue_scorer_all_equal_is_one_and_filter_scopes_minmax():
    prof = [
        {"name": "prefill", "role_mask": abi.FI_ROLE_PREFILLER, "scorers": [(Q, 10)]},
        {"name": "decode", "role_mask": abi.FI_ROLE_DECODER, "scorers": [(Q, 10)]},
    ]
    o, _ = _oracle(E=4, profiles=prof)
    roles = np.array([2, 2, 4, 4], dtype=np.uint32)
    q = np.array([7, 7, 1, 3], dtype=np.int32)
    o.update_endpoints(H.states_array(4, queue=q, roles=roles))
    data, offs = H.pack_prompts([bytes(64)])
    pk = o.pick_batch(data, offs, 1)
    assert pk[0, 0]["endpoint"] == 0 and pk[0, 0]["score"] == 10.0  # both prefillers equal → 1.0
    assert pk[0, 1]["endpoint"] == 2 and pk[0, 1]["score"] == 10.0  # min queue among decoders


def test_pd_threshold_rule():
    prof = [
        {"name": "prefill", "role_mask": abi.FI_ROLE_PREFILLER, "scorers": [(P, 50)]},
        {"name": "decode", "role_mask": abi.FI_ROLE_DECODER, "scorers": [(P, 50)]},
    ]
    blob = bytes(range(256))  # 4 blocks, 256 bytes
    ch = _chain(blob)
    roles = np.array([2, 4], dtype=np.uint32)
    data, offs = H.pack_prompts([blob])
    # decoder (endpoint 1) holds 3 of 4 blocks → hit 0.75 → (1-0.75)*256 = 64 uncached bytes
    for thr, runs in ((0.0, True), (64.0, True), (64.5, False)):
        o, _ = _oracle(E=2, profiles=prof, pd={"decode": 1, "prefill": 0, "threshold": thr})
        o.update_endpoints(H.states_array(2, roles=roles))
        o.index_apply(H.ops_array([(ch[i], 1, 1) for i in range(3)]))
        pk = o.pick_batch(data, offs, 1234)
        assert pk[0, 1]["endpoint"] == 1 and pk[0, 1]["match_blocks"] == 3
        assert (pk[0, 0]["endpoint"] == 0) == runs
        if not runs:
            assert pk[0, 0]["endpoint"] == abi.FI_NO_ENDPOINT


def test_lru_add_chain_eviction_order():
    o, _ = _oracle(E=2, lru_capacity=4)
    o.update_endpoints(H.states_array(2))
    o.index_add_chain(0, np.array([1, 2, 3, 4], dtype=np.uint64))
    assert all(o.index_contains(0, h) for h in (1, 2, 3, 4))
    o.index_add_chain(0, np.array([1], dtype=np.uint64))      # touch 1 → 2 is now oldest
    o.index_add_chain(0, np.array([5], dtype=np.uint64))      # evicts 2
    assert not o.index_contains(0, 2) and all(o.index_contains(0, h) for h in (1, 3, 4, 5))
    o.index_add_chain(0, np.array([6, 7], dtype=np.uint64))   # evicts 3 then 4
    assert [o.index_contains(0, h) for h in (1, 3, 4, 5, 6, 7)] == [True, False, False, True, True, True]
    assert not o.index_contains(1, 1)


def test_short_and_empty_prompts():
    o, _ = _oracle()
    o.update_endpoints(H.states_array(4))
    data, offs = H.pack_prompts([b"", bytes(63), bytes(64), bytes(64 * 40)])
    pk, ch = o.pick_batch(data, offs, 9, want_chains=True)
    assert list(pk[:, 0]["n_blocks"]) == [0, 0, 1, 16]  # truncated at max_blocks
    assert not ch[0].any() and not ch[1].any()


def test_multithreaded_matches_single():
    wl = H.small_workload()
    cfg = H.config_for(wl, profiles=[{"name": "d", "scorers": [(P, 100), (K, 3), (Q, 5)]}])
    o = eo.Oracle(cfg)
    o.update_endpoints(wl.endpoint_states())
    for ops in wl.index_ops():
        o.index_apply(ops)
    tok, offs = wl.prompts()
    a = o.pick_batch(tok, offs, wl.h0, nthreads=1)
    b = o.pick_batch(tok, offs, wl.h0, nthreads=4)
    assert H.picks_equal(a, b)
    assert (a["match_blocks"] > 0).sum() > len(a) // 2  # the workload really exercises prefix hits


def _lora_states(E, rows):
    """rows: {endpoint: (max_active, [active ids], [waiting ids])}"""
    from fusioninfer_b200 import LORA_DTYPE

    st = np.zeros(len(rows), dtype=LORA_DTYPE)
    for i, (e, (mx, act, wai)) in enumerate(sorted(rows.items())):
        st[i]["endpoint"] = e
        st[i]["max_active"] = mx
        st[i]["n_active"] = len(act)
        st[i]["n_waiting"] = len(wai)
        st[i]["active"][: len(act)] = act
        st[i]["waiting"][: len(wai)] = wai
    return st


def test_lora_affinity_scorer_classes():
    """upstream lora-affinity-scorer (strategy.go:100-113): active 1.0 > room 0.8 > queued 0.6 > none 0."""
    L = abi.FI_SCORER_LORA
    prof = [{"name": "default", "scorers": [(L, 100)]}]
    o, _ = _oracle(E=5, profiles=prof)
    o.update_endpoints(H.states_array(5))
    o.update_endpoints_lora(_lora_states(5, {
        0: (2, [11, 12], [77]),      # full, 77 queued            -> 0.6 for 77, 0 otherwise
        1: (4, [11], []),            # room                       -> 1.0 for 11, 0.8 otherwise
        2: (1, [77], []),            # full, 77 active            -> 1.0 for 77
        3: (1, [12], [13]),          # full, nothing relevant     -> 0
        # endpoint 4 never listed: max_active 0 -> 0
    }))
    data, offs = H.pack_prompts([bytes(64)] * 3)
    pk = o.pick_batch(data, offs, 1, adapters=np.array([77, 11, 99], dtype=np.uint64))
    assert (pk[0, 0]["endpoint"], pk[0, 0]["score"]) == (2, 100.0)   # active beats room (0.8) and queued (0.6)
    from tests import restate

    start = restate.tie_start(1, _chain(bytes(64), h0=1)[0], 1, 1, 5)  # endpoints 0 and 1 both have 11 active: rotation
    assert (pk[1, 0]["endpoint"], pk[1, 0]["score"]) == (min((0, 1), key=lambda e: (e - start) % 5), 100.0)
    assert (pk[2, 0]["endpoint"], pk[2, 0]["score"]) == (1, 80.0)    # only endpoint 1 has room
    # without room anywhere the queued endpoint wins
    o.update_endpoints_lora(_lora_states(5, {1: (1, [11], [])}))
    pk = o.pick_batch(data[:64 + 16], offs[:2], 1, adapters=np.array([77], dtype=np.uint64))
    assert (pk[0, 0]["endpoint"], pk[0, 0]["score"]) == (2, 100.0)
    o.update_endpoints_lora(_lora_states(5, {2: (1, [5], [])}))
    pk = o.pick_batch(data[:64 + 16], offs[:2], 1, adapters=np.array([77], dtype=np.uint64))
    assert (pk[0, 0]["endpoint"], pk[0, 0]["score"]) == (0, 60.0)
