"""Host-side mirror of the Endpoint Picker scheduling interface over the C ABI.

`EndpointPicker` plays the role of the upstream EPP's scheduler object for the
ONE path this repo implements: it is configured from the EndpointPickerConfig
YAML that FusionInfer's router role generates
(/root/reference/pkg/router/strategy.go:27-165), receives pod-state refreshes
and prefix-index updates, and schedules batches of requests.  Every call goes
through include/fi_epp.h into the sm_100a kernels; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi as abi

PICK_DTYPE, OP_DTYPE, ENDPOINT_DTYPE = abi.np_dtypes()
LORA_DTYPE = abi.lora_dtype()


class FiEppError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        name = abi.load().fi_epp_status_string(status).decode()
        super().__init__(f"{where}: {name} ({status}){': ' + detail if detail else ''}")


def default_config() -> abi.fi_epp_config:
    cfg = abi.fi_epp_config()
    rc = abi.load().fi_epp_config_default(C.byref(cfg))
    if rc != abi.FI_OK:
        raise FiEppError(rc, "fi_epp_config_default")
    return cfg


def config_from_yaml(yaml_text: str, base: Optional[abi.fi_epp_config] = None) -> abi.fi_epp_config:
    """Apply an EndpointPickerConfig YAML (strategy.go:52-67,126-164) onto `base`."""
    cfg = base if base is not None else default_config()
    raw = yaml_text.encode()
    err = C.create_string_buffer(512)
    rc = abi.load().fi_epp_config_from_yaml(raw, len(raw), C.byref(cfg), err, len(err))
    if rc != abi.FI_OK:
        raise FiEppError(rc, "fi_epp_config_from_yaml", err.value.decode())
    return cfg


def make_config(
    *,
    num_endpoints: int,
    block_bytes: int = 64,
    max_blocks: int = 256,
    lru_capacity: int = 0,
    max_batch: int = 1024,
    max_prompt_bytes: int = 0,
    index_slots: int = 0,
    match_mode: int = abi.FI_MATCH_UPSTREAM,
    device: int = 0,
    endpoint_begin: int = 0,
    endpoint_count: Optional[int] = None,
    profiles: Optional[Sequence[dict]] = None,
    pd: Optional[dict] = None,
    base: Optional[abi.fi_epp_config] = None,
) -> abi.fi_epp_config:
    """Build a config in code.  profiles: [{"name", "role_mask", "more_filters": [...], "scorers": [(kind, weight), ...]}].
    base: a pre-filled struct to start from instead of fi_epp_config_default() (then libfi_epp.so is not touched:
    bench.py's CPU reference arm builds its configuration without mapping the product library)."""
    cfg = base if base is not None else default_config()
    cfg.device = device
    cfg.block_bytes = block_bytes
    cfg.max_blocks = max_blocks
    cfg.lru_capacity = lru_capacity
    cfg.num_endpoints = num_endpoints
    cfg.endpoint_begin = endpoint_begin
    cfg.endpoint_count = num_endpoints - endpoint_begin if endpoint_count is None else endpoint_count
    cfg.match_mode = match_mode
    cfg.max_batch = max_batch
    cfg.max_prompt_bytes = max_prompt_bytes
    cfg.index_slots = index_slots
    if profiles is not None:
        cfg.n_profiles = len(profiles)
        for i, p in enumerate(profiles):
            prof = cfg.profiles[i]
            prof.name = p.get("name", f"p{i}").encode()
            prof.role_mask = p.get("role_mask", 0)
            more = list(p.get("more_filters", []))  # further by-label filters, ANDed with role_mask
            prof.n_more_filters = len(more)
            for j, f in enumerate(more):
                prof.more_filters[j] = f
            sc = p["scorers"]
            prof.n_scorers = len(sc)
            for j, (kind, weight) in enumerate(sc):
                prof.scorers[j].kind = kind
                prof.scorers[j].weight = weight
    if pd is not None:
        cfg.pd_enabled = 1
        cfg.pd_decode_profile = pd["decode"]
        cfg.pd_prefill_profile = pd["prefill"]
        cfg.pd_threshold = float(pd.get("threshold", 0.0))
    return cfg


def model_seed(model: bytes, salt: bytes = b"") -> int:
    out = C.c_uint64(0)
    rc = abi.load().fi_epp_model_seed(model, len(model), salt, len(salt), C.byref(out))
    if rc != abi.FI_OK:
        raise FiEppError(rc, "fi_epp_model_seed")
    return out.value


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EndpointPicker:
    """One handle = one GPU = one endpoint-range shard of the pool."""

    def __init__(self, cfg: abi.fi_epp_config):
        self._lib = abi.load()
        self.cfg = cfg
        h = C.c_void_p()
        rc = self._lib.fi_epp_create(C.byref(cfg), C.byref(h))
        if rc != abi.FI_OK:
            raise FiEppError(rc, "fi_epp_create", "see stderr")
        self._h = h
        self.n_profiles = cfg.n_profiles
        self.max_blocks = cfg.max_blocks

    # -- lifecycle ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.fi_epp_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, where: str):
        if rc != abi.FI_OK:
            raise FiEppError(rc, where, self._lib.fi_epp_last_error(self._h).decode())

    # -- pod datastore -------------------------------------------------------
    def update_endpoints(self, states: np.ndarray):
        states = np.ascontiguousarray(states, dtype=ENDPOINT_DTYPE)
        self._check(self._lib.fi_epp_endpoints_update(self._h, _ptr(states), len(states)), "fi_epp_endpoints_update")

    def update_endpoints_lora(self, states: np.ndarray):
        """Adapter residency per endpoint (LORA_DTYPE rows) for the lora-affinity-scorer."""
        states = np.ascontiguousarray(states, dtype=LORA_DTYPE)
        self._check(self._lib.fi_epp_endpoints_lora_update(self._h, _ptr(states), len(states)),
                    "fi_epp_endpoints_lora_update")

    # -- prefix index --------------------------------------------------------
    def index_apply(self, ops: np.ndarray):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        self._check(self._lib.fi_epp_index_apply(self._h, _ptr(ops), len(ops)), "fi_epp_index_apply")

    def index_add_chain(self, endpoint: int, hashes: np.ndarray):
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        self._check(
            self._lib.fi_epp_index_add_chain(self._h, endpoint, _ptr(hashes), len(hashes)), "fi_epp_index_add_chain"
        )

    def index_add_chains(self, endpoints: np.ndarray, chains: np.ndarray, nblocks: np.ndarray):
        """indexer.Add(chains[r, :nblocks[r]], endpoints[r]) for a whole batch of decisions (upstream PreRequest);
        chains: [R, pitch] u64 as returned by pick_batch(want_chains=True).  Collective on a sharded pool."""
        endpoints = np.ascontiguousarray(endpoints, dtype=np.uint32)
        nblocks = np.ascontiguousarray(nblocks, dtype=np.uint32)
        chains = np.ascontiguousarray(chains, dtype=np.uint64)
        R = len(endpoints)
        pitch = chains.shape[1] if chains.ndim == 2 else (chains.size // max(R, 1))
        self._check(self._lib.fi_epp_index_add_chains(self._h, _ptr(endpoints), _ptr(chains), pitch, _ptr(nblocks), R),
                    "fi_epp_index_add_chains")

    def index_add_chains_device(self, endpoints: np.ndarray, d_chains: int, pitch: int, nblocks: np.ndarray, stream: int = 0):
        """The same with the chains in device memory (pointer; e.g. the chains_out of pick_batch_device, written on
        `stream`): only the two small host arrays cross PCIe.  Needs the device-resident LRU (the default)."""
        endpoints = np.ascontiguousarray(endpoints, dtype=np.uint32)
        nblocks = np.ascontiguousarray(nblocks, dtype=np.uint32)
        self._check(self._lib.fi_epp_index_add_chains_device(self._h, _ptr(endpoints), C.c_void_p(d_chains), pitch, _ptr(nblocks),
                                                             len(endpoints), C.c_void_p(stream)),
                    "fi_epp_index_add_chains_device")

    def lru_dump(self, endpoint: int) -> np.ndarray:
        """Keys of `endpoint` in the device-resident LRU, least recently used first (diagnostics)."""
        cap = max(int(self.cfg.lru_capacity), 1) + 1
        out = np.zeros(cap, dtype=np.uint64)
        n = C.c_uint32(0)
        self._check(self._lib.fi_epp_lru_dump(self._h, endpoint, _ptr(out), cap, C.byref(n)), "fi_epp_lru_dump")
        return out[: n.value].copy()

    def pipeline_info(self) -> dict:
        """How pick_submit runs: partitioned GPU (three batches in flight) or not (two)."""
        out = (C.c_int32 * 3)()
        self._check(self._lib.fi_epp_pipeline_info(self._h, out), "fi_epp_pipeline_info")
        return {"partitioned": bool(out[0]), "walk_sms": int(out[1]), "main_sms": int(out[2])}

    def lru_counters(self) -> dict:
        """Totals of the device-resident LRU since create (diagnostics)."""
        out = (C.c_uint64 * 6)()
        self._check(self._lib.fi_epp_lru_counters(self._h, out), "fi_epp_lru_counters")
        return dict(zip(("sets", "clears", "doomed", "maintained", "deferred_requests", "sub_batches"), [int(x) for x in out]))

    def index_sync(self):
        self._check(self._lib.fi_epp_index_sync(self._h), "fi_epp_index_sync")

    def index_contains(self, queries: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(queries, dtype=OP_DTYPE)
        out = np.zeros(len(q), dtype=np.uint8)
        self._check(self._lib.fi_epp_index_contains(self._h, _ptr(q), len(q), _ptr(out)), "fi_epp_index_contains")
        return out

    def index_stats(self) -> abi.fi_index_stats:
        st = abi.fi_index_stats()
        self._check(self._lib.fi_epp_index_stats(self._h, C.byref(st)), "fi_epp_index_stats")
        return st

    # -- hashing / scheduling --------------------------------------------------
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
        """-> (chains [R, max_blocks] u64, nblocks [R] u32)"""
        prompts, offsets, h0, R = self._inputs(prompts, offsets, h0)
        chains = np.zeros((R, self.max_blocks), dtype=np.uint64)
        nb = np.zeros(R, dtype=np.uint32)
        self._check(
            self._lib.fi_epp_hash_batch(self._h, _ptr(prompts), _ptr(offsets), _ptr(h0), R, _ptr(chains), _ptr(nb)),
            "fi_epp_hash_batch",
        )
        return chains, nb

    def pick_batch(self, prompts, offsets, h0, want_chains: bool = False, adapters=None):
        """Schedule R requests held in host memory.  -> picks [R, n_profiles] (PICK_DTYPE)[, chains]
        adapters: optional [R] uint64 target adapter ids (lora-affinity-scorer)."""
        prompts, offsets, h0, R = self._inputs(prompts, offsets, h0)
        picks = np.zeros((R, self.n_profiles), dtype=PICK_DTYPE)
        chains = np.zeros((R, self.max_blocks), dtype=np.uint64) if want_chains else None
        if adapters is None:
            rc = self._lib.fi_epp_pick_batch(self._h, _ptr(prompts), _ptr(offsets), _ptr(h0), R, _ptr(picks), _ptr(chains))
        else:
            ad = np.ascontiguousarray(np.broadcast_to(np.asarray(adapters, dtype=np.uint64), (R,)))
            rc = self._lib.fi_epp_pick_batch_lora(self._h, _ptr(prompts), _ptr(offsets), _ptr(h0), _ptr(ad), R,
                                                  _ptr(picks), _ptr(chains))
        self._check(rc, "fi_epp_pick_batch")
        return (picks, chains) if want_chains else picks

    def pick_batch_raw(self, prompts_ptr: int, offsets_ptr: int, h0_ptr: int, R: int, out_ptr: int, chains_ptr: int = 0):
        """Host-pointer variant without numpy marshalling (pinned buffers from pinned_alloc)."""
        self._check(
            self._lib.fi_epp_pick_batch(self._h, prompts_ptr, offsets_ptr, h0_ptr, R, out_ptr, chains_ptr or None),
            "fi_epp_pick_batch",
        )

    def pick_batch_device(self, d_prompts: int, d_offsets: int, d_h0: int, R: int, total_bytes: int, d_out: int,
                          d_chains: int = 0, stream: int = 0):
        """Every buffer already resident in this handle's device memory (raw device pointers)."""
        self._check(
            self._lib.fi_epp_pick_batch_device(
                self._h, d_prompts, d_offsets, d_h0, R, total_bytes, d_out, d_chains or None, stream or None
            ),
            "fi_epp_pick_batch_device",
        )

    def pick_submit(self, d_prompts: int, d_offsets: int, d_h0: int, R: int, total_bytes: int, d_out: int, stream: int = 0):
        """Pipelined device path: enqueue a batch; its result is valid for `stream` only after pick_wait."""
        self._check(
            self._lib.fi_epp_pick_submit(self._h, d_prompts, d_offsets, d_h0, R, total_bytes, d_out, stream or None),
            "fi_epp_pick_submit",
        )

    def pick_wait(self, stream: int = 0):
        """Make `stream` wait (device-side) for every batch submitted so far."""
        self._check(self._lib.fi_epp_pick_wait(self._h, stream or None), "fi_epp_pick_wait")

    # -- multi-GPU -------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * abi.FI_EPP_UNIQUE_ID_BYTES)()
        rc = abi.load().fi_epp_comm_unique_id(buf)
        if rc != abi.FI_OK:
            raise FiEppError(rc, "fi_epp_comm_unique_id")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * abi.FI_EPP_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self._lib.fi_epp_comm_init(self._h, buf, rank, world), "fi_epp_comm_init")

    def comm_exchange(self) -> str:
        """'none' (one rank), 'peer' (in-kernel stores over NVLink peer memory) or 'nccl' (all-gathers)."""
        rc = self._lib.fi_epp_comm_exchange(self._h)
        if rc < 0:
            raise FiEppError(rc, "fi_epp_comm_exchange")
        return {0: "none", 1: "peer", 2: "nccl"}[rc]

    def set_option(self, name: str, value: int):
        """Runtime knobs of include/fi_epp.h (fi_epp_set_option): exchange, shard_hash, feed_slices, lru_threads."""
        self._check(self._lib.fi_epp_set_option(self._h, name.encode(), int(value)), f"fi_epp_set_option({name})")

    # -- stats -----------------------------------------------------------------
    def set_profiling(self, on: bool):
        self._check(self._lib.fi_epp_set_profiling(self._h, 1 if on else 0), "fi_epp_set_profiling")

    def stats(self) -> abi.fi_epp_stats:
        st = abi.fi_epp_stats()
        self._check(self._lib.fi_epp_get_stats(self._h, C.byref(st)), "fi_epp_get_stats")
        return st

    def reset_stats(self):
        self._check(self._lib.fi_epp_reset_stats(self._h), "fi_epp_reset_stats")


class PinnedBuffer:
    """Page-locked host memory from the library (so H2D copies are truly asynchronous)."""

    def __init__(self, nbytes: int):
        self._lib = abi.load()
        self.nbytes = nbytes
        self.ptr = self._lib.fi_epp_pinned_alloc(nbytes)
        if not self.ptr:
            raise MemoryError(f"fi_epp_pinned_alloc({nbytes}) failed")

    def array(self, dtype, count: Optional[int] = None) -> np.ndarray:
        dt = np.dtype(dtype)
        n = self.nbytes // dt.itemsize if count is None else count
        buf = (C.c_uint8 * (n * dt.itemsize)).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dt, count=n)

    def free(self):
        if self.ptr:
            self._lib.fi_epp_pinned_free(self.ptr)
            self.ptr = None
