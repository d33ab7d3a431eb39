#!/usr/bin/env python
"""Development tool: per-node Gaussian interpolation (Q2, n^3) for several rows-per-block settings with whatever
library DFX_LIB points to (A/B of two builds in one gpurun call)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dune_functions_b200 as dfx
from dune_functions_b200 import _capi as capi
L = capi.lib()
def timeit(fn, reps=20, warm=4):
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
    out = {"lib": tag, "n": n}
    for rows in (-1, 8, 16, 32):
        L.dfx_set_kernel_path(capi.OP_TUNE_ROWS, rows)
        out[f"rows_{rows}"] = round(timeit(lambda: dfx.interpolate(b, x, dfx.gaussian(1.0))), 4)
    L.dfx_set_kernel_path(capi.OP_TUNE_ROWS, -1)
    out["checksum"] = float(x.sum().item())
    print(json.dumps(out), flush=True)
    del b, x
    torch.cuda.empty_cache()
