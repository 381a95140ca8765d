#!/usr/bin/env python
"""Small end-to-end exercise of every code path of the step kernel (for compute-sanitizer):
both default shapes, j-split combine, graph replay, accelerations, odd sizes."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nbx = importlib.import_module("nbody-demo-2023_b200").nbx
for n, opts in [(1000, {}), (2049, {"j_splits": 5}), (9000, {"j_splits": 3, "graph": 1}), (12288, {"graph": 0}), (7, {})]:
    arrs = nbx.ic(n)
    with nbx.Context(n) as c:
        for k, v in opts.items():
            c.set_option(k, v)
        c.upload(*arrs)
        acc = c.accelerations()
        ke, _ = c.run(5)
        st = c.state()
        assert np.all(np.isfinite(ke)) and np.all(np.isfinite(st[0])) and np.all(np.isfinite(acc))
        print(n, opts, c.info()["variant"], c.info()["j_splits"], ke[-1])
print("ok")
