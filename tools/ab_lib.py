#!/usr/bin/env python
"""A/B two builds of libnbx.so in one process on one box (interleaved, median).
    python tools/ab_lib.py N steps reps libA.so libB.so"""
import ctypes as C, importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("nbody-demo-2023_b200")
n, steps, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
libs = sys.argv[4:]
arrs = pkg.nbx.ic(n)
f32p = C.POINTER(C.c_float); f64p = C.POINTER(C.c_double)
H = []
for path in libs:
    L = C.CDLL(os.path.abspath(path))
    L.nbx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    L.nbx_upload.argtypes = [C.c_void_p] + [f32p] * 7
    L.nbx_run.argtypes = [C.c_void_p, C.c_int, f64p, f64p]
    L.nbx_last_error.restype = C.c_char_p
    h = C.c_void_p()
    assert L.nbx_create(C.byref(h), n, 0, 0, 1, 0.1, 6.67259e-11, 1e-3) == 0, L.nbx_last_error()
    assert L.nbx_upload(h, *[a.ctypes.data_as(f32p) for a in arrs]) == 0
    s = C.c_double()
    assert L.nbx_run(h, 2, None, C.byref(s)) == 0, L.nbx_last_error()
    H.append((path, L, h))
res = {p: [] for p, _, _ in H}
for r in range(reps):
    for p, L, h in H:
        s = C.c_double()
        assert L.nbx_run(h, steps, None, C.byref(s)) == 0
        res[p].append(s.value / steps)
for p, v in res.items():
    med = float(np.median(v))
    print(f"{os.path.basename(p):24s} N={n} med {med*1e3:10.4f} ms  {float(n)*n/med/1e9:8.1f} Gpairs/s  {float(n)*n/med/1e9*20e-3/74.45*100:5.1f}% peak", flush=True)
