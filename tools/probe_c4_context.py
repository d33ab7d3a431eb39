#!/usr/bin/env python
"""Development probe: DG4 192^3 values + gradients timed back to back and inside the bench's alternation with the
per-node interpolation of the quadratic, for both coefficient-load variants of k_eval_pair."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dune_functions_b200 as dfx
from dune_functions_b200 import _capi as capi
L = capi.lib()
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
b = dfx.makeBasis(dfx.YaspGrid([192, 192, 192]), dfx.lagrangeDG(4))
ne = b.numElements()
x = b.newVector(0.0).uniform_(-1, 1)
y = b.newVector(0.0)
quad = dfx.quadratic(0.0, [0, 0, 0], [42, 7, 0, 13, 0, 5])
pts, _ = dfx.tensor_gauss(3, 5)
v = torch.empty((ne, 125, 1), dtype=torch.float64, device="cuda")
j = torch.empty((ne, 125, 1, 3), dtype=torch.float64, device="cuda")
f = dfx.makeDiscreteGlobalBasisFunction(b, x)
nbj = b.dimension() * 8 + v.numel() * 8 + j.numel() * 8
def ev():
    f.evaluate(pts, jacobians=True, out_values=v, out_jacobians=j)
for rep in range(3):
    for mode, name in ((2, "per warp"), (1, "per block")):
        L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_PAIR, mode)
        for ctx in ("back to back", "alternating with interpolate"):
            ts = []
            for it in range(12):
                if ctx != "back to back":
                    dfx.interpolate(b, y, quad)
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ev(); e.record(); torch.cuda.synchronize()
                if it >= 3:
                    ts.append(a.elapsed_time(e))
            ms = float(np.median(ts))
            print(json.dumps({"probe": "DG4 192^3 values+gradients", "bulk_load": name, "context": ctx, "ms": ms, "frac": nbj / ms / 1e6 / PEAK}), flush=True)
L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_PAIR, -1)
