#!/usr/bin/env python
"""Per-pair rounding of two formulations of the force loop, emulated in numpy (CPU, no GPU needed).

  current   d = r_j - r_i;  r2 = d.d + eps;  s = Gm_j * rsqrt(r2)^3;            a += s * d       (12 packed FP32 instr)
  q-scaled  q_j = Gm_j^(-1/2);  d' = fma(-r_i, q_j, q_j r_j);  r2' = d'.d' + q_j^2 eps;
            u = rsqrt(r2')^3 = Gm^(3/2) (r2)^(-3/2);             a += u * d'  ( = Gm r^-3 d)     (11 packed FP32 instr)

The q-scaled form drops one multiply per pair but rounds q_j r_j once per body, which acts like a half-ulp
perturbation of r_j: for a near pair that is an error of 2^-24 |r_j| / |d| relative to d.  This script measures
what that does to the force on sampled bodies at full N (sums in double, so that only the per-pair arithmetic
differs).   python tests/qscale_emulation.py [N] [uniform|plummer] [samples]

Outcome (round 2): accuracy is no obstacle (N >= 65 536: both forms 1e-8 from fp64 per body; N = 2000: 2e-7
instead of 2e-8).  With the two j-bodies of a record in the packed lanes the kernel built from it was 3-4 % SLOWER
than the 12-instruction pair (profiles/r02_qscaled_pair_experiment.{patch,log}): its three subtract-FMAs read two
64-bit registers plus a 32-bit one, three registers from one bank.  With the two i-bodies of a record in the lanes
the j data are scalars, the subtract-FMA reads a pair and two scalars of opposite register parity, and the kernel
is 2.9 % FASTER at N = 1 M: that form is the default from 65 536 bodies on (nbx_kernels.cuh, the QS block).
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import importlib  # noqa: E402

f32, f64 = np.float32, np.float64
G, EPS = f32(6.67259e-11), f32(1e-3)


def fma(a, b, c):
    return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)   # product exact in double


def rsqrt(x):
    return (1.0 / np.sqrt(x.astype(f64))).astype(f32)


def errors(n, ic="uniform", ns=256):
    """Relative force error against fp64 of the two formulations on `ns` sampled bodies: (current, q-scaled)."""
    from oracle import oracle as O
    if ic == "uniform":
        s = O.ic_uniform(n)
        x, y, z, m = s.px, s.py, s.pz, s.mass
    else:
        nbx = importlib.import_module("nbody-demo-2023_b200").nbx
        a = nbx.ic(n, "plummer")
        x, y, z, m = a[0], a[1], a[2], a[6]
    gm = (G * m).astype(f32)
    gm_c = np.maximum(gm, f32(1e-30))
    q = (1.0 / np.sqrt(gm_c.astype(f64))).astype(f32)
    qx, qy, qz = (q * x).astype(f32), (q * y).astype(f32), (q * z).astype(f32)
    q2e = ((q * q).astype(f32) * EPS).astype(f32)
    sel = np.random.default_rng(5).choice(n, ns, replace=False)
    err_cur, err_new = [], []
    X, Y, Z, GM = x.astype(f64), y.astype(f64), z.astype(f64), gm.astype(f64)
    for i in sel:
        xi, yi, zi = x[i], y[i], z[i]
        # fp64 truth from the fp32 inputs
        dX, dY, dZ = X - f64(xi), Y - f64(yi), Z - f64(zi)
        w = GM * (dX * dX + dY * dY + dZ * dZ + f64(EPS)) ** -1.5
        t = np.array([np.sum(w * dX), np.sum(w * dY), np.sum(w * dZ)])
        # current formulation
        dx, dy, dz = x - xi, y - yi, z - zi
        r2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, np.full(n, EPS))))
        ri = rsqrt(r2)
        sc = ((ri * ri).astype(f32) * ri).astype(f32)
        sc = (sc * gm).astype(f32)
        c = np.array([np.sum(sc.astype(f64) * dx), np.sum(sc.astype(f64) * dy), np.sum(sc.astype(f64) * dz)])
        # q-scaled formulation
        nxi, nyi, nzi = np.full(n, -xi), np.full(n, -yi), np.full(n, -zi)
        ex, ey, ez = fma(nxi, q, qx), fma(nyi, q, qy), fma(nzi, q, qz)
        r2q = fma(ez, ez, fma(ey, ey, fma(ex, ex, q2e)))
        rq = rsqrt(r2q)
        u = ((rq * rq).astype(f32) * rq).astype(f32)
        nw = np.array([np.sum(u.astype(f64) * ex), np.sum(u.astype(f64) * ey), np.sum(u.astype(f64) * ez)])
        tn = np.linalg.norm(t)
        err_cur.append(np.linalg.norm(c - t) / tn)
        err_new.append(np.linalg.norm(nw - t) / tn)
    return np.array(err_cur), np.array(err_new)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    ic = sys.argv[2] if len(sys.argv) > 2 else "uniform"
    ns = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    err_cur, err_new = errors(n, ic, ns)
    for name, e in (("current", err_cur), ("q-scaled", err_new)):
        print(f"N={n} {ic:8s} {name:9s} force error vs fp64 over {ns} bodies: median {np.median(e):.2e}  p90 {np.quantile(e, .9):.2e}  max {e.max():.2e}")


if __name__ == "__main__":
    main()
