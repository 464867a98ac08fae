"""Synthetic request / endpoint-state generator.

Shared by bench.py and the tests so that the GPU path, the CPU oracle and the
reference arm all see bit-identical inputs.  Everything derives from a
counter-based SplitMix64 (vectorised with numpy); block hashes of the shared
prefix groups are computed with the python `xxhash` package — an XXH64
implementation independent of both the kernels and the oracle — so a wrong GPU
hash can never "agree" with the index state by construction.

Workload: tokens are uint32 LE uniform in [0, 128000); G = 4·E
prefix groups; a request picks a group Zipf(s=1), shares a prefix whose length is
uniform over multiples of one block in [T/4, 3T/4], then continues with unique
tokens; 10 % of requests are fully unique.  Every endpoint holds the full chains
of 8 groups (prefix-closed) plus private filler hashes up to
lruCapacityPerServer entries (/root/reference/pkg/router/strategy.go:59).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Iterator, Tuple

import numpy as np

from . import _abi as abi

U64 = np.uint64
_GAMMA = U64(0x9E3779B97F4A7C15)
_M1 = U64(0xBF58476D1CE4E5B9)
_M2 = U64(0x94D049BB133111EB)

SEEDS = {1: 0xF0510001, 2: 0xF0510002, 3: 0xF0510003, 4: 0xF0510004, 5: 0xF0510005}
MODEL_NAME = b"synthetic/model"
VOCAB = 128000


def sm64(x):
    """SplitMix64 output function of state x (numpy uint64, wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=U64) + _GAMMA
        z = (z ^ (z >> U64(30))) * _M1
        z = (z ^ (z >> U64(27))) * _M2
        return z ^ (z >> U64(31))


def stream_key(seed: int, stream: int, sub: int = 0):
    with np.errstate(over="ignore"):
        return sm64(sm64(U64(seed & 0xFFFFFFFFFFFFFFFF) + U64(stream)) + U64(sub))


def stream_values(key, idx):
    with np.errstate(over="ignore"):
        return sm64(np.asarray(key, dtype=U64) + np.asarray(idx, dtype=U64))


def xxh64_py(data: bytes) -> int:
    import xxhash

    return xxhash.xxh64_intdigest(data)


def chain_py(block_bytes_seq: bytes, block_bytes: int, max_blocks: int, h0: int) -> np.ndarray:
    """SURVEY.md Appendix A.1 restated over python xxhash (independent of oracle and kernels)."""
    import xxhash

    n = min(len(block_bytes_seq) // block_bytes, max_blocks)
    out = np.zeros(n, dtype=np.uint64)
    prev = h0
    pack = struct.Struct("<Q").pack
    f = xxhash.xxh64_intdigest
    for i in range(n):
        prev = f(block_bytes_seq[i * block_bytes : (i + 1) * block_bytes] + pack(prev))
        out[i] = prev
    return out


@dataclass
class Workload:
    R: int  # requests per batch
    E: int  # endpoints in the pool
    T: int  # tokens per prompt
    seed: int = SEEDS[3]
    block_tokens: int = 16  # 16 uint32 tokens = 64 B blocks
    max_blocks: int = 256
    lru_capacity: int = 31250
    groups_per_endpoint: int = 8
    unique_frac: float = 0.10
    holes: bool = False  # clear 5 % of the membership bits (exercises upstream vs LPM semantics)
    pd: bool = False  # first half of the pool prefillers, second half decoders

    @property
    def block_bytes(self) -> int:
        return self.block_tokens * 4

    @property
    def G(self) -> int:
        return 4 * self.E

    @property
    def n_blocks(self) -> int:
        return min(self.T // self.block_tokens, self.max_blocks)

    @property
    def h0(self) -> int:
        return xxh64_py(MODEL_NAME)

    # ---- prompts -----------------------------------------------------------------
    def group_base(self, groups: np.ndarray) -> np.ndarray:
        """tokens [len(groups), T] of the given prefix groups"""
        keys = stream_key(self.seed, 1, 0) + np.asarray(groups, dtype=U64) * U64(0x100000001B3)
        with np.errstate(over="ignore"):
            v = stream_values(keys[:, None], np.arange(self.T, dtype=U64)[None, :])
        return (v % U64(VOCAB)).astype(np.uint32)

    def _zipf_groups(self, u: np.ndarray) -> np.ndarray:
        w = 1.0 / np.arange(1, self.G + 1, dtype=np.float64)
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        return np.minimum(np.searchsorted(cdf, u, side="right"), self.G - 1).astype(np.int64)

    def request_params(self, batch: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """(group [R], shared_tokens [R]); shared_tokens == 0 for fully unique requests"""
        key = stream_key(self.seed, 3, batch)
        r = np.arange(self.R, dtype=U64)
        ug = stream_values(key, r * U64(4)).astype(np.float64) / 2.0**64
        ul = stream_values(key, r * U64(4) + U64(1))
        uu = stream_values(key, r * U64(4) + U64(2)).astype(np.float64) / 2.0**64
        groups = self._zipf_groups(ug)
        lo = (self.T // 4) // self.block_tokens
        hi = (3 * self.T // 4) // self.block_tokens
        span = max(hi - lo + 1, 1)
        shared = (lo + (ul % U64(span)).astype(np.int64)) * self.block_tokens
        shared[uu < self.unique_frac] = 0
        return groups, shared

    def prompts(self, batch: int = 0, chunk: int = 1024):
        """-> (tokens uint32 [R, T], offsets uint64 [R+1] in bytes)"""
        groups, shared = self.request_params(batch)
        out = np.empty((self.R, self.T), dtype=np.uint32)
        pos = np.arange(self.T, dtype=np.int64)[None, :]
        idx = np.arange(self.T, dtype=U64)[None, :]
        for lo in range(0, self.R, chunk):
            hi = min(self.R, lo + chunk)
            rk = stream_key(self.seed, 2, batch) + np.arange(lo, hi, dtype=U64) * U64(0x100000001B3)
            with np.errstate(over="ignore"):
                suf = (stream_values(rk[:, None], idx) % U64(VOCAB)).astype(np.uint32)
            base = self.group_base(groups[lo:hi])
            out[lo:hi] = np.where(pos < shared[lo:hi, None], base, suf)
        offsets = np.arange(self.R + 1, dtype=np.uint64) * np.uint64(self.T * 4)
        return out, offsets

    # ---- endpoint state --------------------------------------------------------------
    def endpoint_states(self) -> np.ndarray:
        _, _, ep_dt = abi.np_dtypes()
        st = np.zeros(self.E, dtype=ep_dt)
        e = np.arange(self.E, dtype=U64)
        key = stream_key(self.seed, 6, 0)
        st["endpoint"] = np.arange(self.E, dtype=np.uint32)
        st["kv_util"] = (stream_values(key, e * U64(2)) % U64(1024)).astype(np.float64) / 1024.0
        st["queue_depth"] = (stream_values(key, e * U64(2) + U64(1)) % U64(32)).astype(np.int32)
        st["flags"] = abi.FI_ENDPOINT_ALIVE
        if self.pd:
            st["role_mask"] = np.where(np.arange(self.E) < self.E // 2, abi.FI_ROLE_PREFILLER, abi.FI_ROLE_DECODER)
        else:
            st["role_mask"] = abi.FI_ROLE_WORKER
        return st

    def endpoint_groups(self) -> np.ndarray:
        """[E, groups_per_endpoint] group ids each endpoint caches"""
        e = np.arange(self.E, dtype=U64)[:, None]
        j = np.arange(self.groups_per_endpoint, dtype=U64)[None, :]
        key = stream_key(self.seed, 4, 0)
        return (stream_values(key, e * U64(64) + j) % U64(self.G)).astype(np.int64)

    def group_chains(self, groups: np.ndarray) -> dict:
        """{group id: chain hashes [n_blocks]} via python xxhash"""
        uniq = np.unique(groups)
        out = {}
        h0 = self.h0
        step = 256
        for lo in range(0, len(uniq), step):
            ids = uniq[lo : lo + step]
            base = self.group_base(ids)
            for k, g in enumerate(ids):
                out[int(g)] = chain_py(base[k].tobytes(), self.block_bytes, self.max_blocks, h0)
        return out

    def index_ops(self, ep_lo: int = 0, ep_hi: int | None = None, chunk_endpoints: int = 64) -> Iterator[np.ndarray]:
        """Yield SET ops (OP dtype) describing the initial index state of endpoints [ep_lo, ep_hi)."""
        _, op_dt, _ = abi.np_dtypes()
        ep_hi = self.E if ep_hi is None else ep_hi
        eg = self.endpoint_groups()
        chains = self.group_chains(eg[ep_lo:ep_hi])
        nb = self.n_blocks
        n_fill = max(self.lru_capacity - self.groups_per_endpoint * nb, 0)
        hole_key = stream_key(self.seed, 7, 0)
        for lo in range(ep_lo, ep_hi, chunk_endpoints):
            hi = min(ep_hi, lo + chunk_endpoints)
            parts_h, parts_e = [], []
            for e in range(lo, hi):
                hs = [chains[int(g)] for g in eg[e]]
                fk = stream_key(self.seed, 5, e)
                hs.append(stream_values(fk, np.arange(n_fill, dtype=U64)))
                h = np.concatenate(hs)
                parts_h.append(h)
                parts_e.append(np.full(len(h), e, dtype=np.uint32))
            h = np.concatenate(parts_h)
            ee = np.concatenate(parts_e)
            if self.holes:
                with np.errstate(over="ignore"):
                    u = stream_values(hole_key, h ^ (ee.astype(U64) * U64(0x9E3779B97F4A7C15)))
                keep = (u.astype(np.float64) / 2.0**64) >= 0.05
                h, ee = h[keep], ee[keep]
            ops = np.zeros(len(h), dtype=op_dt)
            ops["hash"] = h
            ops["endpoint"] = ee
            ops["op"] = abi.FI_OP_SET
            yield ops


# the five BASELINE.json configs; cfg 1 is the CPU plumbing case
def baseline_workload(cfg: int, **over) -> Workload:
    table = {
        1: dict(R=64, E=8, T=256),
        2: dict(R=4096, E=256, T=2048),
        3: dict(R=16384, E=1024, T=4096),
        4: dict(R=65536, E=4096, T=4096),
        5: dict(R=16384, E=1024, T=4096, pd=True),
    }
    kw = dict(table[cfg])
    kw["seed"] = SEEDS[cfg]
    kw.update(over)
    return Workload(**kw)


# scoring profiles of the BASELINE configs
def baseline_profiles(cfg: int):
    P, K, Q = abi.FI_SCORER_PREFIX, abi.FI_SCORER_KV_UTIL, abi.FI_SCORER_QUEUE
    if cfg == 5:  # "KV-queue-weighted score": prefix 50 + kv 25 + queue 25 on each PD profile
        return (
            [
                {"name": "prefill", "role_mask": abi.FI_ROLE_PREFILLER, "scorers": [(P, 50), (K, 25), (Q, 25)]},
                {"name": "decode", "role_mask": abi.FI_ROLE_DECODER, "scorers": [(P, 50), (K, 25), (Q, 25)]},
            ],
            {"decode": 1, "prefill": 0, "threshold": 0.0},
        )
    # generatePrefixCacheConfig (strategy.go:51-68): prefix scorer weight 100
    return [{"name": "default", "role_mask": 0, "scorers": [(P, 100)]}], None
