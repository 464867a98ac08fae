"""ctypes wrapper of oracle/libepp_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package fusioninfer_b200
never does (tests/test_boundary.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from fusioninfer_b200 import _abi as abi  # struct layouts of include/fi_epp.h only

PICK_DTYPE, OP_DTYPE, ENDPOINT_DTYPE = abi.np_dtypes()
_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libepp_oracle.so")
_lib = None
_P = C.c_void_p


def build():
    subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.epo_xxh64.restype = C.c_uint64
    lib.epo_xxh64.argtypes = [_P, C.c_uint64, C.c_uint64]
    lib.epo_create.restype = _P
    lib.epo_create.argtypes = [C.POINTER(abi.fi_epp_config)]
    lib.epo_destroy.restype = None
    lib.epo_destroy.argtypes = [_P]
    lib.epo_endpoints_update.restype = C.c_int
    lib.epo_endpoints_update.argtypes = [_P, _P, C.c_uint32]
    lib.epo_endpoints_lora_update.restype = C.c_int
    lib.epo_endpoints_lora_update.argtypes = [_P, _P, C.c_uint32]
    lib.epo_pick_batch_lora.restype = C.c_int
    lib.epo_pick_batch_lora.argtypes = [_P, _P, _P, _P, _P, C.c_uint32, _P, _P, C.c_uint32]
    lib.epo_index_reserve.restype = C.c_int
    lib.epo_index_reserve.argtypes = [_P, C.c_uint64]
    lib.epo_index_apply.restype = C.c_int
    lib.epo_index_apply.argtypes = [_P, _P, C.c_uint64]
    lib.epo_index_add_chain.restype = C.c_int
    lib.epo_index_add_chain.argtypes = [_P, C.c_uint32, _P, C.c_uint32]
    lib.epo_index_keys.restype = C.c_uint64
    lib.epo_index_keys.argtypes = [_P]
    lib.epo_index_contains.restype = C.c_int
    lib.epo_index_contains.argtypes = [_P, C.c_uint32, C.c_uint64]
    lib.epo_hash_batch.restype = C.c_int
    lib.epo_hash_batch.argtypes = [_P, _P, _P, _P, C.c_uint32, _P, _P]
    lib.epo_pick_batch.restype = C.c_int
    lib.epo_pick_batch.argtypes = [_P, _P, _P, _P, C.c_uint32, _P, _P, C.c_uint32]
    lib.epo_pick_batch_repeat.restype = C.c_int
    lib.epo_pick_batch_repeat.argtypes = [_P, _P, _P, _P, C.c_uint32, _P, C.c_uint32, C.c_uint32]
    lib.epo_index_add_chains.restype = C.c_int
    lib.epo_index_add_chains.argtypes = [_P, _P, _P, C.c_uint32, _P, C.c_uint32]
    lib.epo_usable_cores.restype = C.c_uint32
    lib.epo_usable_cores.argtypes = []
    lib.epo_config_default.restype = C.c_int
    lib.epo_config_default.argtypes = [C.POINTER(abi.fi_epp_config)]
    _lib = lib
    return lib


def usable_cores() -> int:
    """host cores the CPU legs may use: affinity mask capped by the cgroup CPU quota"""
    return int(load().epo_usable_cores())


def default_config() -> abi.fi_epp_config:
    """generatePrefixCacheConfig defaults filled by the ORACLE library (the reference arm of bench.py must not
    load the product library)."""
    cfg = abi.fi_epp_config()
    rc = load().epo_config_default(C.byref(cfg))
    assert rc == 0, rc
    return cfg


def xxh64(data: bytes, seed: int = 0) -> int:
    return load().epo_xxh64(data, len(data), seed)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


class Oracle:
    """CPU restatement of the whole pool (never sharded): the checker for the GPU path."""

    def __init__(self, cfg: abi.fi_epp_config):
        self._lib = load()
        # the oracle always models the full pool
        self.cfg = abi.fi_epp_config.from_buffer_copy(cfg)
        self._h = self._lib.epo_create(C.byref(self.cfg))
        if not self._h:
            raise RuntimeError("epo_create failed (see stderr)")
        self.P = cfg.n_profiles
        self.M = cfg.max_blocks

    def close(self):
        if self._h:
            self._lib.epo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_endpoints(self, states):
        states = np.ascontiguousarray(states, dtype=ENDPOINT_DTYPE)
        rc = self._lib.epo_endpoints_update(self._h, _ptr(states), len(states))
        assert rc == 0, rc

    def update_endpoints_lora(self, states):
        states = np.ascontiguousarray(states, dtype=abi.lora_dtype())
        rc = self._lib.epo_endpoints_lora_update(self._h, _ptr(states), len(states))
        assert rc == 0, rc

    def index_reserve(self, keys: int):
        self._lib.epo_index_reserve(self._h, keys)

    def index_apply(self, ops):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        rc = self._lib.epo_index_apply(self._h, _ptr(ops), len(ops))
        assert rc == 0, rc

    def index_add_chain(self, endpoint: int, hashes):
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        rc = self._lib.epo_index_add_chain(self._h, endpoint, _ptr(hashes), len(hashes))
        assert rc == 0, rc

    def index_add_chains(self, endpoints, chains, nblocks):
        endpoints = np.ascontiguousarray(endpoints, dtype=np.uint32)
        nblocks = np.ascontiguousarray(nblocks, dtype=np.uint32)
        chains = np.ascontiguousarray(chains, dtype=np.uint64)
        rc = self._lib.epo_index_add_chains(self._h, _ptr(endpoints), _ptr(chains), chains.shape[1], _ptr(nblocks), len(endpoints))
        assert rc == 0, rc

    def index_contains(self, endpoint: int, h: int) -> bool:
        return bool(self._lib.epo_index_contains(self._h, endpoint, h))

    @staticmethod
    def _inputs(prompts, offsets, h0):
        prompts = np.ascontiguousarray(np.frombuffer(prompts, dtype=np.uint8) if isinstance(prompts, (bytes, bytearray)) else prompts)
        if prompts.dtype != np.uint8:
            prompts = prompts.view(np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        R = len(offsets) - 1
        h0 = np.ascontiguousarray(np.broadcast_to(np.asarray(h0, dtype=np.uint64), (R,)))
        return prompts, offsets, h0, R

    def hash_batch(self, prompts, offsets, h0):
        prompts, offsets, h0, R = self._inputs(prompts, offsets, h0)
        chains = np.zeros((R, self.M), dtype=np.uint64)
        nb = np.zeros(R, dtype=np.uint32)
        rc = self._lib.epo_hash_batch(self._h, _ptr(prompts), _ptr(offsets), _ptr(h0), R, _ptr(chains), _ptr(nb))
        assert rc == 0, rc
        return chains, nb

    def pick_batch(self, prompts, offsets, h0, want_chains=False, nthreads=1, adapters=None):
        prompts, offsets, h0, R = self._inputs(prompts, offsets, h0)
        picks = np.zeros((R, self.P), dtype=PICK_DTYPE)
        chains = np.zeros((R, self.M), dtype=np.uint64) if want_chains else None
        if adapters is not None:
            ad = np.ascontiguousarray(np.broadcast_to(np.asarray(adapters, dtype=np.uint64), (R,)))
            rc = self._lib.epo_pick_batch_lora(self._h, _ptr(prompts), _ptr(offsets), _ptr(h0), _ptr(ad), R, _ptr(picks),
                                               _ptr(chains), nthreads)
            assert rc == 0, rc
            return (picks, chains) if want_chains else picks
        rc = self._lib.epo_pick_batch(self._h, _ptr(prompts), _ptr(offsets), _ptr(h0), R, _ptr(picks), _ptr(chains), nthreads)
        assert rc == 0, rc
        return (picks, chains) if want_chains else picks

    def pick_batch_repeat(self, prompts, offsets, h0, nthreads=1, repeat=1):
        """Timing helper: each thread walks its shard `repeat` times; returns the picks."""
        prompts, offsets, h0, R = self._inputs(prompts, offsets, h0)
        picks = np.zeros((R, self.P), dtype=PICK_DTYPE)
        rc = self._lib.epo_pick_batch_repeat(self._h, _ptr(prompts), _ptr(offsets), _ptr(h0), R, _ptr(picks), nthreads, repeat)
        assert rc == 0, rc
        return picks
