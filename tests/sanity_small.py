#!/usr/bin/env python
"""Ragged-size / split / graph / shard matrix of the step kernel against the oracle, for whichever
build of the library NBX_LIB selects.  tests/test_gpu_debug_build.py runs it with the bounds-checked
libnbx_debug.so (every store index and ticket value is checked on the device; compute-sanitizer is
closed on the GPU pool).  Exit code 0 = all cases agree with the oracle and no device check fired.

    NBX_LIB=libnbx_debug.so python tests/sanity_small.py
"""
import importlib
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
nbx = importlib.import_module("nbody-demo-2023_b200").nbx
from oracle import oracle as O  # noqa: E402  (test infrastructure)


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-300))


def main():
    print("library:", nbx.LIB_PATH, flush=True)
    cases = 0
    names = nbx.variant_names()
    for n in (1, 2, 7, 8, 9, 63, 100, 257, 1000, 2049, 5000, 20011):
        s = O.ic_uniform(n)
        ke_o = O.run(s, 4, variant="ver2" if n <= 5000 else "ver7")
        arrs = nbx.ic(n)
        for opts in ({}, {"j_splits": 1}, {"j_splits": 3}, {"j_splits": 61, "graph": 1}, {"graph": 0, "pdl": 1},
                     {"accurate": 1}, {"variant": names.index("r4_t128_u2")}, {"variant": names.index("r4_t512_u2")},
                     {"variant": names.index("r4_t64_u2"), "j_splits": 5},
                     {"variant": names.index("r4_t256_u4_stage_f2_qi")},
                     {"variant": names.index("r4_t256_u4_stage_f2_qi"), "j_splits": 3, "graph": 1}):
            with nbx.Context(n) as c:
                for k, v in opts.items():
                    c.set_option(k, v)
                c.upload(*arrs)
                ke, _ = c.run(4)
                out = c.state()
                acc = c.accelerations()
                assert c.info()["device_error"] == 0
            assert np.max(np.abs(ke - ke_o) / ke_o) < 1e-4, (n, opts, ke, ke_o)
            assert rel(np.stack(out[:3], axis=1), s.pos()) < 1e-4, (n, opts)
            assert acc.shape[0] >= n
            cases += 1
    # i-shards of a bigger world on one GPU (accelerations only: no peer is needed for those)
    n = 8200
    arrs = nbx.ic(n)
    with nbx.Context(n) as c:
        c.set_option("j_splits", 4)
        c.upload(*arrs)
        full = c.accelerations()
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            with nbx.Context(n, rank=r, world=world) as c:
                c.set_option("j_splits", 4)
                c.upload(*arrs)
                parts.append(c.accelerations())
                assert c.info()["device_error"] == 0
        got = np.concatenate(parts)[:n]
        assert rel(got, full[:n]) < 1e-6, world
        cases += 1
    # a many-wave grid with whole + split tiles, PDL on (the default large-N plan)
    n = 460000
    arrs = nbx.ic(n)
    with nbx.Context(n) as c:
        c.upload(*arrs)
        ke, _ = c.run(2)
        info = c.info()
        assert info["whole_tiles"] > 0 and info["j_splits"] > 1 and info["device_error"] == 0
        st = O.State(n)
        st.vx, st.vy, st.vz = c.state()[3:6]
        st.mass = arrs[6]
        assert abs(ke[-1] - O.kenergy_fp64(st)) / ke[-1] < 1e-6
    cases += 1
    print(f"sanity_small: {cases} cases OK", flush=True)


if __name__ == "__main__":
    main()
