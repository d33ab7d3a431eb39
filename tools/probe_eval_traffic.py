#!/usr/bin/env python
"""Four launches of the Q2 512^3 evaluation (strip schedule x evict-first stores) for an ncu DRAM-traffic capture."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_functions_b200 as dfx  # noqa: E402
from dune_functions_b200 import _capi as capi  # noqa: E402
L = capi.lib()
b = dfx.makeBasis(dfx.YaspGrid([512, 512, 512]), dfx.lagrange(2))
x = b.newVector(0.0)
dfx.interpolate(b, x, dfx.gaussian(1.0), separable=True)
pts, _ = dfx.tensor_gauss(3, 3)
v = torch.empty((b.numElements(), 27, 1), dtype=torch.float64, device="cuda")
f = dfx.makeDiscreteGlobalBasisFunction(b, x)
for swz in (0, 1):
    for hint in (0, 1):
        L.dfx_set_kernel_path(capi.OP_TUNE_SWIZZLE, swz)
        L.dfx_set_kernel_path(capi.OP_TUNE_STREAM_HINT, hint)
        f.evaluate(pts, out_values=v)
        print("launched strip_schedule", swz, "evict_first", hint)
torch.cuda.synchronize()
