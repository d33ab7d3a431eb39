#!/usr/bin/env python
"""Development tool: one timing of the per-node Gaussian interpolation (Q2, n^3 elements) and of the Taylor-Hood batched
interpolation with whatever library DFX_LIB points to — run alternately with two builds for an A/B in one gpurun call."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import dune_functions_b200 as dfx  # noqa: E402


def timeit(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


tag = sys.argv[1]
for n in (512, 256):
    b = dfx.makeBasis(dfx.YaspGrid([n, n, n]), dfx.lagrange(2))
    x = b.newVector(0.0)
    med, best = timeit(lambda: dfx.interpolate(b, x, dfx.gaussian(1.0)))
    print(json.dumps({"probe": "Q2 per-node Gaussian interpolation", "lib": tag, "n": n, "ms_median": med, "ms_min": best,
                      "checksum": float(x.sum().item())}), flush=True)
    del b, x
    torch.cuda.empty_cache()
