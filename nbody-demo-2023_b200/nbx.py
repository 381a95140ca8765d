"""ctypes mirror of include/nbx.h -- same names, same argument meaning, same error
behaviour (non-zero return -> NbxError carrying nbx_last_error()).  No torch types:
host arrays are numpy float32, device state lives inside the opaque context.

There is no fallback: if libnbx.so is missing or no B200 is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NBX_LIB selects another build of the same ABI (the bounds-checked libnbx_debug.so, the tuning
# tools' libnbx_ablation.so); a bare file name is looked up next to this file.
LIB_PATH = os.environ.get("NBX_LIB") or os.path.join(_HERE, "libnbx.so")
if not os.path.isabs(LIB_PATH) and not os.path.exists(LIB_PATH):
    LIB_PATH = os.path.join(_HERE, LIB_PATH)

ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_STATE, ERR_NODEVICE, ERR_PEER, ERR_DEBUG = 1, 2, 3, 4, 5, 6, 7
EXCHANGE_NCCL = 0
EXCHANGE_P2P = 1
EXCHANGE_NCCL_OVERLAP = 2
UNIQUE_ID_BYTES = 128
P2P_BLOB_BYTES = 256

# every symbol include/nbx.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "nbx_abi_version", "nbx_last_error", "nbx_device_count", "nbx_create", "nbx_destroy",
    "nbx_set_option", "nbx_get_info", "nbx_plan", "nbx_trace_read", "nbx_variant_count", "nbx_variant_name", "nbx_upload",
    "nbx_download", "nbx_upload_sharded", "nbx_upload_group", "nbx_download_shard", "nbx_run", "nbx_accelerations", "nbx_simulate", "nbx_comm_unique_id",
    "nbx_comm_init", "nbx_comm_init_all", "nbx_run_group", "nbx_p2p_export", "nbx_p2p_attach", "nbx_p2p_attach_group",
    "nbx_ic_uniform", "nbx_ic_plummer", "nbx_gflop_per_step", "nbx_host_alloc", "nbx_host_free",
]


class NbxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nbx error {code}: {msg}")
        self.code = code


class Info(C.Structure):
    _fields_ = [(k, C.c_int) for k in (
        "abi_version", "device", "sm_count", "sm_clock_khz", "n", "n_pad", "rank", "world",
        "i_begin", "i_count", "threads", "bodies_per_thread", "tile_bodies", "stages",
        "i_tiles", "whole_tiles", "j_splits", "ctas_per_sm", "use_graph", "exchange", "variant")] + [
        ("kernel_launches", C.c_longlong), ("aux_launches", C.c_longlong),
        ("last_run_seconds", C.c_double), ("kernel_seconds_total", C.c_double),
        ("device_error", C.c_int), ("peer_timeout_ms", C.c_int), ("multicast", C.c_int), ("reserved", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_lib = None


def lib() -> C.CDLL:
    """Load libnbx.so (built in-tree by the package Makefile).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not built -- run __graft_entry__.build() or `make -C {_HERE}`")
        L = C.CDLL(LIB_PATH)
        L.nbx_last_error.restype = C.c_char_p
        L.nbx_variant_name.restype = C.c_char_p
        L.nbx_variant_name.argtypes = [C.c_int]
        L.nbx_gflop_per_step.restype = C.c_double
        L.nbx_gflop_per_step.argtypes = [C.c_int]
        L.nbx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_float, C.c_float, C.c_float]
        L.nbx_destroy.argtypes = [C.c_void_p]
        L.nbx_destroy.restype = None
        L.nbx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
        L.nbx_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.nbx_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.POINTER(Info)]
        L.nbx_upload.argtypes = [C.c_void_p] + [_f32p] * 7
        L.nbx_trace_read.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.nbx_download.argtypes = [C.c_void_p] + [_f32p] * 6
        L.nbx_upload_sharded.argtypes = [C.c_void_p] + [_f32p] * 7
        L.nbx_upload_group.argtypes = [C.POINTER(C.c_void_p), C.c_int] + [_f32p] * 7
        L.nbx_download_shard.argtypes = [C.c_void_p] + [_f32p] * 6
        L.nbx_run.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p]
        L.nbx_accelerations.argtypes = [C.c_void_p] + [_f32p] * 3
        L.nbx_simulate.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float] + [_f32p] * 7 + [_f64p, _f64p]
        L.nbx_comm_unique_id.argtypes = [C.c_void_p]
        L.nbx_comm_init.argtypes = [C.c_void_p, C.c_void_p]
        L.nbx_comm_init_all.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.nbx_run_group.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, _f64p, _f64p]
        L.nbx_p2p_export.argtypes = [C.c_void_p, C.c_void_p]
        L.nbx_p2p_attach.argtypes = [C.c_void_p, C.c_void_p]
        L.nbx_p2p_attach_group.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.nbx_ic_uniform.argtypes = [C.c_int] + [_f32p] * 7
        L.nbx_ic_uniform.restype = None
        L.nbx_ic_plummer.argtypes = [C.c_int] + [_f32p] * 7
        L.nbx_ic_plummer.restype = None
        L.nbx_device_count.argtypes = [C.POINTER(C.c_int)]
        L.nbx_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.nbx_host_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise NbxError(rc, lib().nbx_last_error().decode("utf-8", "replace"))


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "need contiguous float32"
    return a.ctypes.data_as(_f32p)


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().nbx_device_count(C.byref(n)))
    return n.value


def variant_names():
    L = lib()
    return [L.nbx_variant_name(i).decode() for i in range(L.nbx_variant_count())]


def plan(n: int, rank: int = 0, world: int = 1, sm_count: int = 148, exchange: int = EXCHANGE_NCCL,
         variant: int = -1, j_splits: int = 0) -> dict:
    """nbx_plan: the launch plan for a shard, computed on the host (no GPU needed)."""
    i = Info()
    _check(lib().nbx_plan(n, rank, world, sm_count, exchange, variant, j_splits, C.byref(i)))
    return i.as_dict()


def gflop_per_step(n: int) -> float:
    return float(lib().nbx_gflop_per_step(n))


def ic(n: int, kind: str = "uniform"):
    """Initial conditions (px,py,pz,vx,vy,vz,mass), each float32[n]."""
    arrs = [np.zeros(n, dtype=np.float32) for _ in range(7)]
    fn = {"uniform": lib().nbx_ic_uniform, "plummer": lib().nbx_ic_plummer}[kind]
    fn(n, *[_p(a) for a in arrs])
    return arrs


def pinned_empty(n: int) -> np.ndarray:
    """float32[n] in page-locked host memory (nbx_host_alloc); never freed explicitly."""
    ptr = C.c_void_p()
    _check(lib().nbx_host_alloc(C.byref(ptr), n * 4))
    buf = (C.c_float * n).from_address(ptr.value)
    return np.frombuffer(buf, dtype=np.float32)


class Context:
    """One GPU's share of a simulation: nbx_create ... nbx_destroy."""

    def __init__(self, n: int, device: int = 0, rank: int = 0, world: int = 1,
                 dt: float = 0.1, G: float = 6.67259e-11, eps2: float = 1e-3):
        self._h = C.c_void_p()
        _check(lib().nbx_create(C.byref(self._h), n, device, rank, world, dt, G, eps2))
        self.n = n

    def close(self):
        if self._h:
            lib().nbx_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key: str, value: int):
        _check(lib().nbx_set_option(self._h, key.encode(), int(value)))

    def info(self) -> dict:
        i = Info()
        _check(lib().nbx_get_info(self._h, C.byref(i)))
        return i.as_dict()

    def upload(self, px, py, pz, vx, vy, vz, mass):
        _check(lib().nbx_upload(self._h, *[_p(a) for a in (px, py, pz, vx, vy, vz, mass)]))

    def download(self, px, py, pz, vx, vy, vz):
        _check(lib().nbx_download(self._h, *[_p(a) for a in (px, py, pz, vx, vy, vz)]))

    def upload_sharded(self, px, py, pz, vx, vy, vz, mass):
        """Collective: each rank copies only its own shard over PCIe; NCCL all-gathers the packed records."""
        _check(lib().nbx_upload_sharded(self._h, *[_p(a) for a in (px, py, pz, vx, vy, vz, mass)]))

    def download_shard(self, px, py, pz, vx, vy, vz):
        """Writes only this context's [i_begin, i_begin+i_count) of the six arrays."""
        _check(lib().nbx_download_shard(self._h, *[_p(a) for a in (px, py, pz, vx, vy, vz)]))

    def state(self):
        arrs = [np.zeros(self.n, dtype=np.float32) for _ in range(6)]
        self.download(*arrs)
        return arrs

    def run(self, nsteps: int):
        """-> (kenergy float64[nsteps], device seconds)"""
        ke = np.zeros(max(nsteps, 1), dtype=np.float64)
        secs = C.c_double(0.0)
        _check(lib().nbx_run(self._h, nsteps, ke.ctypes.data_as(_f64p), C.byref(secs)))
        return ke[:nsteps], secs.value

    def trace(self):
        """Trace build only: uint64[steps, ctas, 6] of per-CTA %globaltimer stamps (see nbx_trace_read)."""
        steps, ctas = C.c_int(0), C.c_int(0)
        rc = lib().nbx_trace_read(self._h, None, 0, C.byref(steps), C.byref(ctas))
        if rc == ERR_STATE:
            _check(rc)
        buf = np.zeros(max(1, steps.value * ctas.value * 6), dtype=np.uint64)
        _check(lib().nbx_trace_read(self._h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), buf.size, C.byref(steps), C.byref(ctas)))
        return buf.reshape(steps.value, ctas.value, 6)

    def accelerations(self):
        cnt = self.info()["i_count"]
        a = [np.zeros(cnt, dtype=np.float32) for _ in range(3)]
        _check(lib().nbx_accelerations(self._h, *[_p(x) for x in a]))
        return np.stack(a, axis=1)

    # multi-GPU plumbing
    def comm_init(self, unique_id: bytes):
        assert len(unique_id) == UNIQUE_ID_BYTES
        buf = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES)
        _check(lib().nbx_comm_init(self._h, buf))

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(P2P_BLOB_BYTES)
        _check(lib().nbx_p2p_export(self._h, buf))
        return buf.raw

    def p2p_attach(self, blobs: bytes):
        buf = C.create_string_buffer(blobs, len(blobs))
        _check(lib().nbx_p2p_attach(self._h, buf))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(lib().nbx_comm_unique_id(buf))
    return buf.raw


def comm_init_all(ctxs):
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    _check(lib().nbx_comm_init_all(arr, len(ctxs)))


def p2p_attach_group(ctxs):
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    _check(lib().nbx_p2p_attach_group(arr, len(ctxs)))


def upload_group(ctxs, px, py, pz, vx, vy, vz, mass):
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    _check(lib().nbx_upload_group(arr, len(ctxs), *[_p(a) for a in (px, py, pz, vx, vy, vz, mass)]))


def run_group(ctxs, nsteps: int):
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    ke = np.zeros(max(nsteps, 1), dtype=np.float64)
    secs = C.c_double(0.0)
    _check(lib().nbx_run_group(arr, len(ctxs), nsteps, ke.ctypes.data_as(_f64p), C.byref(secs)))
    return ke[:nsteps], secs.value


def simulate(nsteps, px, py, pz, vx, vy, vz, mass, dt=0.1, G=6.67259e-11, eps2=1e-3):
    """nbx_simulate: host buffers in, host buffers out (updated in place)."""
    n = px.shape[0]
    ke = np.zeros(max(nsteps, 1), dtype=np.float64)
    secs = C.c_double(0.0)
    _check(lib().nbx_simulate(n, nsteps, dt, G, eps2, *[_p(a) for a in (px, py, pz, vx, vy, vz, mass)],
                              ke.ctypes.data_as(_f64p), C.byref(secs)))
    return ke[:nsteps], secs.value


def shard_of(n: int, rank: int, world: int):
    """(i_begin, i_count, n_pad) -- the same arithmetic nbx_create uses (host-side logic,
    testable without a GPU): n padded to a multiple of 8*world, equal contiguous shards."""
    grain = 8 * world
    n_pad = (n + grain - 1) // grain * grain
    cnt = n_pad // world
    return rank * cnt, cnt, n_pad
