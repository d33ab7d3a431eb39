#!/usr/bin/env python
"""One pass over the kernels of the BASELINE configs (each launched twice) for an ncu capture:
    python tools/profile_r02.py            # plain run first (must exit 0)
    ncu --set full --clock-control none --import-source on -k regex:"k_eval|k_interp|k_indices" -o gpurun_out/r02_prof python tools/profile_r02.py
Development tool; numbers printed under ncu are never bench values."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dune_functions_b200 as dfx  # noqa: E402
from dune_functions_b200 import _capi as capi  # noqa: E402

L = capi.lib()
which = set(sys.argv[1:]) or {"c5w", "c2", "c3", "c4"}
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
REPS = int(os.environ.get("DFX_PROFILE_REPS", "1"))

if "c5w" in which:
    b = dfx.makeBasis(dfx.YaspGrid([512, 512, 512]), dfx.lagrange(2))
    x = b.newVector(0.0)
    pts, _ = dfx.tensor_gauss(3, 3)
    v = torch.empty((b.numElements(), 27, 1), dtype=torch.float64, device="cuda")
    f = dfx.makeDiscreteGlobalBasisFunction(b, x)
    for _ in range(REPS):
        dfx.interpolate(b, x, dfx.gaussian(1.0))
        f.evaluate(pts, out_values=v)
    torch.cuda.synchronize()
    del b, x, v, f
    torch.cuda.empty_cache()
if "c2" in which:
    b = dfx.makeBasis(dfx.YaspGrid([256, 256, 256]), dfx.lagrange(2))
    x = b.newVector(0.0)
    pts, _ = dfx.tensor_gauss(3, 3)
    v = torch.empty((b.numElements(), 27, 1), dtype=torch.float64, device="cuda")
    f = dfx.makeDiscreteGlobalBasisFunction(b, x)
    for _ in range(REPS):
        dfx.interpolate(b, x, dfx.gaussian(1.0))
        f.evaluate(pts, out_values=v)
    torch.cuda.synchronize()
    del b, x, v, f
    torch.cuda.empty_cache()
if "c3" in which:
    b = dfx.makeBasis(dfx.YaspGrid([128, 128, 128]), dfx.taylorHood())
    ne, nl = b.numElements(), b.maxNodeSize()
    x = b.newVector(0.0)
    idx = torch.empty((ne, nl), dtype=torch.int32, device="cuda")
    mi = torch.empty((ne, nl, 2), dtype=torch.int32, device="cuda")
    pts, _ = dfx.tensor_gauss(3, 3)
    v = torch.empty((ne, 27, 4), dtype=torch.float64, device="cuda")
    f = dfx.makeDiscreteGlobalBasisFunction(b, x)
    parts = [(dfx.subspaceBasis(b, 0), dfx.identity(3)), (dfx.subspaceBasis(b, 1), dfx.gaussian(1.0))]
    for _ in range(REPS):
        capi.check(L.dfx_indices_flat(b._h, 0, ne, C.c_void_p(idx.data_ptr()), st()))
        capi.check(L.dfx_indices_multi(b._h, 0, ne, C.c_void_p(mi.data_ptr()), 2, st()))
        dfx.interpolateBatch(b, x, parts)
        f.evaluate(pts, out_values=v)
    torch.cuda.synchronize()
    del b, x, idx, mi, v, f
    torch.cuda.empty_cache()
if "c4" in which:
    b = dfx.makeBasis(dfx.YaspGrid([192, 192, 192]), dfx.lagrangeDG(4))
    ne = b.numElements()
    x = b.newVector(0.0)
    quad = dfx.quadratic(0.0, [0, 0, 0], [42, 7, 0, 13, 0, 5])
    pts, _ = dfx.tensor_gauss(3, 5)
    v = torch.empty((ne, 125, 1), dtype=torch.float64, device="cuda")
    j = torch.empty((ne, 125, 1, 3), dtype=torch.float64, device="cuda")
    f = dfx.makeDiscreteGlobalBasisFunction(b, x)
    for _ in range(REPS):
        dfx.interpolate(b, x, quad)
        f.evaluate(pts, out_values=v)
        f.evaluate(pts, jacobians=True, out_values=v, out_jacobians=j)
    torch.cuda.synchronize()
if "q3" in which:
    b = dfx.makeBasis(dfx.YaspGrid([128, 128, 128]), dfx.lagrange(3))
    x = b.newVector(0.0)
    pts, _ = dfx.tensor_gauss(3, 4)
    v = torch.empty((b.numElements(), 64, 1), dtype=torch.float64, device="cuda")
    f = dfx.makeDiscreteGlobalBasisFunction(b, x)
    for _ in range(REPS):
        dfx.interpolate(b, x, dfx.gaussian(1.0))
        f.evaluate(pts, out_values=v)
    torch.cuda.synchronize()
    del b, x, v, f
    torch.cuda.empty_cache()
print("profile pass done, launches:", L.dfx_kernel_launch_count())
