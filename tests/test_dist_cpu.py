"""CPU, gloo, world_size 2: the host-side logic of the endpoint-range sharded path.

Each rank runs the CPU oracle restricted to its shard (other endpoints not alive, only
its shard's index entries), the ranks exchange their per-request (score, endpoint)
picks exactly as the GPU path does (all-gather + max with the request's tie rotation), and the
merged result must equal the unsharded oracle.  Also covers shard_range, the
unique-id broadcast and the max-over-ranks timing reduction used by bench.py.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def merge_picks_numpy(gathered, starts, E):
    """[world, R, P] picks -> [R, P]: score desc, then the request's tie rotation (starts[r]: fi_epp.h "Ties"),
    FI_NO_ENDPOINT never wins — what merge_picks_kernel does."""
    from fusioninfer_b200 import _abi as abi

    out = gathered[0].copy()
    st = np.asarray(starts, dtype=np.int64)[:, None]
    for g in gathered[1:]:
        none_o = out["endpoint"] == abi.FI_NO_ENDPOINT
        none_g = g["endpoint"] == abi.FI_NO_ENDPOINT
        rot_o = (out["endpoint"].astype(np.int64) - st) % E
        rot_g = (g["endpoint"].astype(np.int64) - st) % E
        better = (~none_g) & (none_o | (g["score"] > out["score"]) | ((g["score"] == out["score"]) & (rot_g < rot_o)))
        out[better] = g[better]
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist

    from fusioninfer_b200 import _abi as abi
    from fusioninfer_b200 import dist as fdist
    from oracle import epp_oracle as eo
    from tests import helpers as H

    fdist.init_process_group("gloo")
    try:
        # plumbing
        payload = bytes(range(128)) if rank == 0 else None
        got = fdist.broadcast_bytes(payload, 128)
        assert got == bytes(range(128))
        assert fdist.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        assert fdist.sum_over_ranks(1.0) == float(world)

        for mode in (abi.FI_MATCH_UPSTREAM, abi.FI_MATCH_LPM):
            wl = H.small_workload(E=37, R=96)  # 37 endpoints: uneven shards
            prof = [{"name": "d", "scorers": [(H.P, 100), (H.K, 7), (H.Q, 9)]}]
            cfg = H.config_for(wl, profiles=prof, match_mode=mode)
            states = wl.endpoint_states()
            begin, count = fdist.shard_range(wl.E, rank, world)
            # this rank's view: only its shard is alive, only its shard's index entries exist.
            # The queue scorer's min/max is over the whole pool (every rank receives all states),
            # so the shard view keeps the other endpoints' queue depths but marks them ineligible
            # through a role no profile selects... simpler: run the shard with the global min/max
            # injected by two sentinel-free tricks is not possible in the oracle, so the sharded
            # check uses scorers whose per-endpoint value does not depend on the candidate set.
            prof_local = [{"name": "d", "scorers": [(H.P, 100), (H.K, 7)]}]
            cfg_l = H.config_for(wl, profiles=prof_local, match_mode=mode)
            local_states = states.copy()
            outside = (local_states["endpoint"] < begin) | (local_states["endpoint"] >= begin + count)
            local_states["flags"][outside] = 0
            o_local = eo.Oracle(cfg_l)
            o_local.update_endpoints(local_states)
            o_global = eo.Oracle(cfg_l)
            o_global.update_endpoints(states)
            for ops in wl.index_ops():
                o_global.index_apply(ops)
                mine = ops[(ops["endpoint"] >= begin) & (ops["endpoint"] < begin + count)]
                o_local.index_apply(mine)
            tok, offs = wl.prompts()
            local, chains = o_local.pick_batch(tok, offs, wl.h0, want_chains=True)
            gathered = fdist.all_gather_array(local)
            from tests import restate

            starts = [restate.tie_start(int(local[q, 0]["n_blocks"]), int(chains[q, 0]), wl.h0, q, wl.E) for q in range(wl.R)]
            merged = merge_picks_numpy(list(gathered), starts, wl.E)
            want = o_global.pick_batch(tok, offs, wl.h0)
            assert H.picks_equal(merged, want), H.describe_diff(merged, want)
            del cfg
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_pool_exactly():
    from fusioninfer_b200.dist import shard_range

    for E in (1, 7, 8, 37, 1024, 4096):
        for world in (1, 2, 3, 8):
            if world > E:
                continue
            spans = [shard_range(E, r, world) for r in range(world)]
            pos = 0
            for b, c in spans:
                assert b == pos and c >= E // world
                pos += c
            assert pos == E
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_sharded_pick_over_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
