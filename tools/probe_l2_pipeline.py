#!/usr/bin/env python
"""Development probe: does it pay to run interpolate and evaluate slab by slab, so that the evaluation reads the
coefficients the interpolation has just written from the L2 (126 MB) instead of from DRAM?  Q2 on 512^3 elements,
z-slabs of 512 / nchunks element layers handled one after the other on one stream (the multi-GPU decomposition run
sequentially on one GPU: every slab has its own handle and coefficient vector), against the monolithic step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import dune_functions_b200 as dfx  # noqa: E402


def timeit(fn, reps=8, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
grid = dfx.YaspGrid([n, n, n])
pts, _ = dfx.tensor_gauss(3, 3)
f = dfx.gaussian(1.0)
b = dfx.makeBasis(grid, dfx.lagrange(2))
x = b.newVector(0.0)
v = torch.empty((b.numElements(), 27, 1), dtype=torch.float64, device="cuda")
u = dfx.makeDiscreteGlobalBasisFunction(b, x)


def mono():
    dfx.interpolate(b, x, f)
    u.evaluate(pts, out_values=v)


ms0 = timeit(mono)
print(json.dumps({"probe": "monolithic step", "n": n, "ms": ms0}), flush=True)
ref = v[:: 4099].clone()
for nchunks in (16, 32, 64, 128, 256):
    if n % nchunks:
        continue
    parts = []
    layer = n * n
    for k in range(nchunks):
        bk = dfx.makeBasis(grid.slab(k, nchunks), dfx.lagrange(2))
        xk = bk.newVector(0.0)
        uk = dfx.makeDiscreteGlobalBasisFunction(bk, xk)
        e0 = k * (n // nchunks) * layer
        parts.append((bk, xk, uk, v[e0:e0 + bk.numElements()]))

    def piped():
        for bk, xk, uk, vk in parts:
            dfx.interpolate(bk, xk, f)
            uk.evaluate(pts, out_values=vk)

    v.zero_()
    ms = timeit(piped)
    same = bool(torch.equal(v[:: 4099], ref))
    print(json.dumps({"probe": "slab by slab", "n": n, "chunks": nchunks, "layers_per_chunk": n // nchunks,
                      "coefficients_MB_per_chunk": parts[0][1].numel() * 8 / 1e6, "ms": ms, "speedup": ms0 / ms,
                      "same_values": same}), flush=True)
    del parts
    torch.cuda.empty_cache()
