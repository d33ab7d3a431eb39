#!/usr/bin/env python
"""Development tool: one timing of the Q2 evaluation at 3^3 Gauss points (n^3 elements) with whatever library DFX_LIB
points to — run alternately with two builds for an A/B in one gpurun call."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dune_functions_b200 as dfx
def timeit(fn, reps=12, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
tag = sys.argv[1]
for n in (512, 256):
    b = dfx.makeBasis(dfx.YaspGrid([n, n, n]), dfx.lagrange(2))
    x = b.newVector(0.0)
    dfx.interpolate(b, x, dfx.gaussian(1.0))
    pts, _ = dfx.tensor_gauss(3, 3)
    v = torch.empty((b.numElements(), 27, 1), dtype=torch.float64, device="cuda")
    f = dfx.makeDiscreteGlobalBasisFunction(b, x)
    ms = timeit(lambda: f.evaluate(pts, out_values=v))
    print(json.dumps({"probe": "Q2 evaluate at 3^3", "lib": tag, "n": n, "ms": ms, "checksum": float(v[::977].sum().item())}), flush=True)
    del b, x, v, f
    torch.cuda.empty_cache()
