"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes access to the CPU restatement (oracle/nbody_oracle.c -> liboracle.so) and
to the compiled, unmodified reference (oracle/_ref/, built by oracle/Makefile from
/root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path
(nbody-demo-2023_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(with_ref: bool = True) -> None:
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    targets = ["all"] + (["ref"] if with_ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(with_ref=False)
        L = C.CDLL(LIB_PATH)
        L.oracle_ic_uniform.argtypes = [C.c_int] + [_f32p] * 7
        L.oracle_ic_uniform.restype = None
        for name in ("oracle_run_ver2", "oracle_run_ver0", "oracle_run_ver7"):
            fn = getattr(L, name)
            fn.argtypes = [C.c_int] + [_f32p] * 7 + [C.c_float, C.c_int, _f32p]
            fn.restype = C.c_int
        L.oracle_acc_fp64.argtypes = [C.c_int] + [_f32p] * 4 + [C.c_int, _i32p] + [_f64p] * 3
        L.oracle_acc_fp64.restype = None
        L.oracle_acc_f32.argtypes = [C.c_int] + [_f32p] * 4 + [C.c_int, _i32p] + [_f32p] * 3
        L.oracle_acc_f32.restype = None
        L.oracle_run_fp64.argtypes = [C.c_int] + [_f64p] * 6 + [_f32p, C.c_double, C.c_int, _f64p]
        L.oracle_run_fp64.restype = C.c_int
        L.oracle_kenergy_fp64.argtypes = [C.c_int] + [_f32p] * 4
        L.oracle_kenergy_fp64.restype = C.c_double
        L.oracle_gflop_per_step.argtypes = [C.c_int]
        L.oracle_gflop_per_step.restype = C.c_double
        L.oracle_num_threads.restype = C.c_int
        _lib = L
    return _lib


class State:
    """Host SoA particle state: px py pz vx vy vz mass (float32[n] each)."""

    FIELDS = ("px", "py", "pz", "vx", "vy", "vz", "mass")

    def __init__(self, n: int):
        self.n = n
        for f in self.FIELDS:
            setattr(self, f, np.zeros(n, dtype=np.float32))

    def arrays(self):
        return [getattr(self, f) for f in self.FIELDS]

    def copy(self) -> "State":
        s = State(self.n)
        for f in self.FIELDS:
            setattr(s, f, getattr(self, f).copy())
        return s

    def pos(self) -> np.ndarray:
        return np.stack([self.px, self.py, self.pz], axis=1)

    def vel(self) -> np.ndarray:
        return np.stack([self.vx, self.vy, self.vz], axis=1)


def ic_uniform(n: int) -> State:
    s = State(n)
    lib().oracle_ic_uniform(n, *s.arrays())
    return s


def run(state: State, nsteps: int, dt: float = 0.1, variant: str = "ver2") -> np.ndarray:
    """Advance `state` in place by nsteps; returns per-step kenergy (float32[nsteps])."""
    ke = np.zeros(max(nsteps, 1), dtype=np.float32)
    fn = getattr(lib(), "oracle_run_" + variant)
    rc = fn(state.n, *state.arrays(), np.float32(dt), nsteps, ke)
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return ke[:nsteps]


def run_fp64(state: State, nsteps: int, dt: float = float(np.float32(0.1))):
    """fp64 "truth" run from `state` (float inputs widened): returns (pos float64[n,3], vel float64[n,3],
    kenergy float64[nsteps]).  dt defaults to the float value 0.1f the reference uses."""
    p = [np.ascontiguousarray(getattr(state, f), dtype=np.float64) for f in ("px", "py", "pz", "vx", "vy", "vz")]
    ke = np.zeros(max(nsteps, 1), dtype=np.float64)
    if lib().oracle_run_fp64(state.n, *p, np.ascontiguousarray(state.mass, dtype=np.float32), float(dt), nsteps, ke) != 0:
        raise MemoryError("oracle allocation failed")
    return np.stack(p[:3], axis=1), np.stack(p[3:], axis=1), ke[:nsteps]


def acc_fp64(state: State, sel: np.ndarray) -> np.ndarray:
    sel = np.ascontiguousarray(sel, dtype=np.int32)
    out = [np.zeros(sel.size, dtype=np.float64) for _ in range(3)]
    lib().oracle_acc_fp64(state.n, state.px, state.py, state.pz, state.mass, sel.size, sel, *out)
    return np.stack(out, axis=1)


def acc_f32(state: State, sel: np.ndarray) -> np.ndarray:
    """Reference-order float accelerations of the selected bodies (what ver2 holds in acc[])."""
    sel = np.ascontiguousarray(sel, dtype=np.int32)
    out = [np.zeros(sel.size, dtype=np.float32) for _ in range(3)]
    lib().oracle_acc_f32(state.n, state.px, state.py, state.pz, state.mass, sel.size, sel, *out)
    return np.stack(out, axis=1)


def kenergy_fp64(state: State) -> float:
    return float(lib().oracle_kenergy_fp64(state.n, state.vx, state.vy, state.vz, state.mass))


def gflop_per_step(n: int) -> float:
    return float(lib().oracle_gflop_per_step(n))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ----------------------------------------------------------------------------
#  compiled reference (oracle/_ref)
# ----------------------------------------------------------------------------
def ref_available(ver: str = "ver2") -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ref_dump_" + ver))


def read_dump(path: str):
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:4] == b"NBXD", "bad dump magic"
    n, steps = struct.unpack_from("<ii", raw, 4)
    (ke,) = struct.unpack_from("<f", raw, 12)
    (secs,) = struct.unpack_from("<d", raw, 16)
    body = np.frombuffer(raw, dtype=np.float32, offset=24, count=7 * n).reshape(7, n)
    s = State(n)
    for k, f in enumerate(State.FIELDS):
        setattr(s, f, body[k].copy())
    return s, np.float32(ke), secs, steps


def ref_run(ver: str, n: int, nsteps: int, threads: int | None = None):
    """Run the compiled reference `ver` for nsteps; returns (State, kenergy_last, loop_seconds)."""
    exe = os.path.join(REF_DIR, "ref_dump_" + ver)
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    env.setdefault("OMP_PROC_BIND", "close")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "dump.bin")
        subprocess.run([exe, str(n), str(nsteps), out], check=True, env=env,
                       stdout=subprocess.DEVNULL)
        s, ke, secs, _ = read_dump(out)
    return s, ke, secs


def write_dump(path: str, state: "State", steps: int = 0, ke: float = 0.0, secs: float = 0.0) -> None:
    """Write `state` in the NBXD layout (what NBODY_DUMP / ref_dump_verN / ref_state_verN exchange)."""
    with open(path, "wb") as f:
        f.write(b"NBXD" + struct.pack("<iifd", state.n, steps, ke, secs))
        for fld in State.FIELDS:
            f.write(np.ascontiguousarray(getattr(state, fld), dtype=np.float32).tobytes())


def ref_state_available(ver: str = "ver8") -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ref_state_" + ver))


def ref_state_run(ver: str, state: "State", nsteps: int, threads: int | None = None):
    """Run the compiled reference `ver` for nsteps FROM `state` (any initial conditions, e.g. Plummer)
    through ref_state_verN; returns (State after nsteps, kenergy float32[nsteps], loop_seconds)."""
    exe = os.path.join(REF_DIR, "ref_state_" + ver)
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    env.setdefault("OMP_PROC_BIND", "close")
    with tempfile.TemporaryDirectory() as td:
        fin, fout, fke = (os.path.join(td, x) for x in ("in.nbxd", "out.nbxd", "ke.txt"))
        write_dump(fin, state)
        subprocess.run([exe, fin, str(nsteps), fout, fke], check=True, env=env)
        s, _, secs, _ = read_dump(fout)
        ke = np.loadtxt(fke, dtype=np.float64, ndmin=1).astype(np.float32) if nsteps > 0 else np.zeros(0, np.float32)
    return s, ke, secs


def ref_cuda_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ver5_all_cuda", "nbody.x"))


def ref_cuda_rate(n: int, steps: int = 150):
    """Run the reference's own CUDA backend (ver5_all/programming_models/cuda/Compute.cu, unmodified,
    rebuilt for sm_100a) through its CLI and return (G pairs/s, kenergy column) from its own table:
    windows after the first (the first window carries CUDA context creation).  Its timer brackets
    the per-step H2D + kernel + D2H + host Euler/energy loop (cuda/Compute.cu:150-194)."""
    import re
    exe = os.path.join(REF_DIR, "ver5_all_cuda", "nbody.x")
    out = subprocess.run([exe, str(n), str(steps), "gpu"], check=True, capture_output=True, text=True).stdout
    rows = [l.split() for l in out.splitlines() if re.match(r"^ \d+\s", l)]
    if len(rows) < 2:
        raise RuntimeError("reference CUDA backend printed fewer than 2 table rows")
    secs = sum(float(r[3]) for r in rows[1:])
    nsteps = 50 * (len(rows) - 1)
    return float(n) * float(n) * nsteps / secs / 1e9, [r[2] for r in rows]
