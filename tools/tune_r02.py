#!/usr/bin/env python
"""Round-2 tuning probes (development tool): rows per block of the lean per-node interpolation kernel,
DG4 per-node interpolation, and the accuracy of the two DG4 gradient forms against the analytic gradient."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import dune_functions_b200 as dfx  # noqa: E402
from dune_functions_b200 import _capi as capi  # noqa: E402

L = capi.lib()
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    what = set(sys.argv[1:]) or {"rows", "dg4", "jac"}
    if "rows" in what:
        for n in (256, 512):
            b = dfx.makeBasis(dfx.YaspGrid([n, n, n]), dfx.lagrange(2))
            x = b.newVector(0.0)
            nd = b.dimension()
            for rows in (1, 2, 4, 8, 16, 32):
                L.dfx_set_kernel_path(capi.OP_TUNE_ROWS, rows)
                ms = timeit(lambda: dfx.interpolate(b, x, dfx.gaussian(1.0)))
                print(json.dumps({"probe": "lean rows", "n": n, "rows": rows, "ms": ms, "frac": nd * 8 / ms / 1e6 / PEAK}), flush=True)
            L.dfx_set_kernel_path(capi.OP_TUNE_ROWS, -1)
            ms = timeit(lambda: dfx.interpolate(b, x, dfx.quadratic(0.5, [1, 2, 3], [1, 2, 3, 4, 5, 6])))
            print(json.dumps({"probe": "quadratic per node", "n": n, "ms": ms, "frac": nd * 8 / ms / 1e6 / PEAK}), flush=True)
            ms = timeit(lambda: dfx.interpolate(b, x, dfx.gaussian(1.0), separable=True))
            print(json.dumps({"probe": "separable tables", "n": n, "ms": ms, "frac": nd * 8 / ms / 1e6 / PEAK}), flush=True)
            del b, x
            torch.cuda.empty_cache()
    if "swz" in what:
        b = dfx.makeBasis(dfx.YaspGrid([512, 512, 512]), dfx.lagrange(2))
        x = b.newVector(0.0)
        dfx.interpolate(b, x, dfx.gaussian(1.0))
        pts, _ = dfx.tensor_gauss(3, 3)
        v = torch.empty((b.numElements(), 27, 1), dtype=torch.float64, device="cuda")
        f = dfx.makeDiscreteGlobalBasisFunction(b, x)
        nb = b.dimension() * 8 + v.numel() * 8
        for rep in range(2):
            for swz in (0, 1):
                for hint in (0, 1):
                    L.dfx_set_kernel_path(capi.OP_TUNE_SWIZZLE, swz)
                    L.dfx_set_kernel_path(capi.OP_TUNE_STREAM_HINT, hint)
                    ms = timeit(lambda: f.evaluate(pts, out_values=v), reps=10)
                    print(json.dumps({"probe": "eval 512^3", "strip_schedule": swz, "evict_first_stores": hint, "ms": ms, "frac": nb / ms / 1e6 / PEAK}), flush=True)
        L.dfx_set_kernel_path(capi.OP_TUNE_SWIZZLE, -1)
        L.dfx_set_kernel_path(capi.OP_TUNE_STREAM_HINT, -1)
        del b, x, v, f
        torch.cuda.empty_cache()
    if "evalblock" in what:
        for n in (256, 512):
            b = dfx.makeBasis(dfx.YaspGrid([n, n, n]), dfx.lagrange(2))
            x = b.newVector(0.0)
            dfx.interpolate(b, x, dfx.gaussian(1.0))
            pts, _ = dfx.tensor_gauss(3, 3)
            v = torch.empty((b.numElements(), 27, 1), dtype=torch.float64, device="cuda")
            f = dfx.makeDiscreteGlobalBasisFunction(b, x)
            nb = b.dimension() * 8 + v.numel() * 8
            for rep in range(3):
                for epb in (64, 32):
                    L.dfx_set_kernel_path(capi.OP_TUNE_EVAL_BLOCK, epb)
                    ms = timeit(lambda: f.evaluate(pts, out_values=v), reps=10)
                    print(json.dumps({"probe": "eval Q2 values", "n": n, "elements_per_block": epb, "ms": ms, "frac": nb / ms / 1e6 / PEAK}), flush=True)
            L.dfx_set_kernel_path(capi.OP_TUNE_EVAL_BLOCK, -1)
            del b, x, v, f
            torch.cuda.empty_cache()
    if "c3" in what:
        b = dfx.makeBasis(dfx.YaspGrid([128, 128, 128]), dfx.taylorHood())
        ne, nl = b.numElements(), b.maxNodeSize()
        x = b.newVector(0.0)
        nd = b.dimension()
        parts = [(dfx.subspaceBasis(b, 0), dfx.identity(3)), (dfx.subspaceBasis(b, 1), dfx.gaussian(1.0))]
        for rows in (-1, 2, 4, 6, 8, 12, 16, 32):
            L.dfx_set_kernel_path(capi.OP_TUNE_ROWS, rows)
            ms = timeit(lambda: dfx.interpolateBatch(b, x, parts))
            print(json.dumps({"probe": "TH batch interpolate", "rows": rows, "ms": ms, "frac": nd * 8 / ms / 1e6 / PEAK}), flush=True)
        L.dfx_set_kernel_path(capi.OP_TUNE_ROWS, -1)
        idx = torch.empty((ne, nl), dtype=torch.int32, device="cuda")
        mi = torch.empty((ne, nl, 2), dtype=torch.int32, device="cuda")
        st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
        for path, label in ((1, "tile + bulk store"), (2, "direct stores")):
            L.dfx_set_kernel_path(capi.OP_INDICES, path)
            ms = timeit(lambda: capi.check(L.dfx_indices_flat(b._h, 0, ne, C.c_void_p(idx.data_ptr()), st())))
            print(json.dumps({"probe": "TH indices flat " + label, "ms": ms, "frac": idx.numel() * 4 / ms / 1e6 / PEAK}), flush=True)
            ms = timeit(lambda: capi.check(L.dfx_indices_multi(b._h, 0, ne, C.c_void_p(mi.data_ptr()), 2, st())))
            print(json.dumps({"probe": "TH indices 2 digits " + label, "ms": ms, "frac": mi.numel() * 4 / ms / 1e6 / PEAK}), flush=True)
        L.dfx_set_kernel_path(capi.OP_INDICES, -1)
        del b, x, idx, mi
        torch.cuda.empty_cache()
        b = dfx.makeBasis(dfx.YaspGrid([256, 256, 256]), dfx.lagrange(2))
        ne = b.numElements()
        idx = torch.empty((ne, 27), dtype=torch.int32, device="cuda")
        for path, label in ((1, "tile + bulk store"), (2, "direct stores")):
            L.dfx_set_kernel_path(capi.OP_INDICES, path)
            ms = timeit(lambda: capi.check(L.dfx_indices_flat(b._h, 0, ne, C.c_void_p(idx.data_ptr()), st())))
            print(json.dumps({"probe": "Q2 256^3 indices flat " + label, "ms": ms, "frac": idx.numel() * 4 / ms / 1e6 / PEAK}), flush=True)
        L.dfx_set_kernel_path(capi.OP_INDICES, -1)
        del b, idx
        torch.cuda.empty_cache()
    if "c3eval" in what:
        b = dfx.makeBasis(dfx.YaspGrid([128, 128, 128]), dfx.taylorHood())
        ne = b.numElements()
        x = b.newVector(0.0).uniform_(-1, 1)
        pts, _ = dfx.tensor_gauss(3, 3)
        v = torch.empty((ne, 27, 4), dtype=torch.float64, device="cuda")
        f = dfx.makeDiscreteGlobalBasisFunction(b, x)
        nb = b.dimension() * 8 + v.numel() * 8
        ref = None
        for rep in range(3):
            for mode in (0, 1):
                L.dfx_set_kernel_path(capi.OP_TUNE_SPLIT_MAP, mode)
                ms = timeit(lambda: f.evaluate(pts, out_values=v), reps=10)
                if ref is None:
                    ref = v.clone()
                print(json.dumps({"probe": "Taylor-Hood 128^3 whole-tree evaluate", "warps_per_leaf_group": mode, "ms": ms,
                                  "frac": nb / ms / 1e6 / PEAK, "same_values": bool(torch.equal(v, ref))}), flush=True)
        L.dfx_set_kernel_path(capi.OP_TUNE_SPLIT_MAP, -1)
        del b, x, v, f
        torch.cuda.empty_cache()
    if "c4prefetch" in what:
        b = dfx.makeBasis(dfx.YaspGrid([192, 192, 192]), dfx.lagrangeDG(4))
        x = b.newVector(0.0).uniform_(-1, 1)
        ne = b.numElements()
        pts, _ = dfx.tensor_gauss(3, 5)
        v = torch.empty((ne, 125, 1), dtype=torch.float64, device="cuda")
        j = torch.empty((ne, 125, 1, 3), dtype=torch.float64, device="cuda")
        f = dfx.makeDiscreteGlobalBasisFunction(b, x)
        nbj = b.dimension() * 8 + v.numel() * 8 + j.numel() * 8
        nbv = b.dimension() * 8 + v.numel() * 8
        for rep in range(3):
            for ahead in (0, 2, 3, 4):
                L.dfx_set_kernel_path(capi.OP_TUNE_PAIR_PREFETCH, ahead)
                ms = timeit(lambda: f.evaluate(pts, jacobians=True, out_values=v, out_jacobians=j), reps=7)
                print(json.dumps({"probe": "DG4 192^3 values+gradients", "l2_prefetch_iterations_ahead": ahead, "ms": ms,
                                  "frac": nbj / ms / 1e6 / PEAK}), flush=True)
                ms = timeit(lambda: f.evaluate(pts, out_values=v), reps=7)
                print(json.dumps({"probe": "DG4 192^3 values", "l2_prefetch_iterations_ahead": ahead, "ms": ms,
                                  "frac": nbv / ms / 1e6 / PEAK}), flush=True)
        L.dfx_set_kernel_path(capi.OP_TUNE_PAIR_PREFETCH, -1)
        del b, x, v, j, f
        torch.cuda.empty_cache()
    if "c4eval" in what:
        b = dfx.makeBasis(dfx.YaspGrid([192, 192, 192]), dfx.lagrangeDG(4))
        x = b.newVector(0.0).uniform_(-1, 1)
        ne = b.numElements()
        pts, _ = dfx.tensor_gauss(3, 5)
        v = torch.empty((ne, 125, 1), dtype=torch.float64, device="cuda")
        j = torch.empty((ne, 125, 1, 3), dtype=torch.float64, device="cuda")
        f = dfx.makeDiscreteGlobalBasisFunction(b, x)
        nbv = b.dimension() * 8 + v.numel() * 8
        nbj = nbv + j.numel() * 8
        for rep in range(3):
            # pair: 0 = k_eval_smem (one element per warp at a time), 2 / 1 = k_eval_pair, 8 warps per block, one bulk load per warp / per block
            for pair, tma in ((0, 1), (2, 1), (1, 1)):
                L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_PAIR, pair)
                L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_TMA, tma)
                ms = timeit(lambda: f.evaluate(pts, out_values=v), reps=5)
                print(json.dumps({"probe": "DG4 192^3 values", "pair_kernel_warps": pair, "tma_pair_loads": tma, "ms": ms, "frac": nbv / ms / 1e6 / PEAK}), flush=True)
                ms = timeit(lambda: f.evaluate(pts, jacobians=True, out_values=v, out_jacobians=j), reps=5)
                print(json.dumps({"probe": "DG4 192^3 values+gradients", "pair_kernel_warps": pair, "tma_pair_loads": tma, "ms": ms, "frac": nbj / ms / 1e6 / PEAK}), flush=True)
        L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_TMA, -1)
        L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_PAIR, -1)
        del b, x, v, j, f
        torch.cuda.empty_cache()
    if "dg4" in what or "jac" in what:
        b = dfx.makeBasis(dfx.YaspGrid([192, 192, 192]), dfx.lagrangeDG(4))
        x = b.newVector(0.0)
        nd = b.dimension()
        quad = dfx.quadratic(0.0, [0, 0, 0], [42, 7, 0, 13, 0, 5])
        if "dg4" in what:
            for name, f in (("quadratic", quad), ("gaussian", dfx.gaussian(1.0))):
                ms = timeit(lambda: dfx.interpolate(b, x, f), reps=5)
                print(json.dumps({"probe": "dg4 per node " + name, "ms": ms, "frac": nd * 8 / ms / 1e6 / PEAK}), flush=True)
        if "jac" in what:
            import orc
            dfx.interpolate(b, x, quad)
            pts, _ = dfx.tensor_gauss(3, 5)
            tp = torch.from_numpy(pts).cuda()
            o = orc.OracleBasis(b.grid, dfx.lagrangeDG(4))
            u = dfx.makeDiscreteGlobalBasisFunction(b, x)
            ne = b.numElements()
            for e0 in (0, ne // 2 // 128 * 128, ne - 4096):
                e1 = e0 + 4096
                e = torch.arange(e0, e1, device="cuda")
                ec = torch.stack([e % 192, (e // 192) % 192, e // (192 * 192)], 1).double()
                P = (ec[:, None, :] + tp[None]) / 192
                X, Y, Z = P[..., 0], P[..., 1], P[..., 2]
                exact = torch.stack([84 * X + 7 * Y, 26 * Y + 7 * X, 10 * Z], -1).cpu().numpy()
                idx = b.flatIndices(e0, e1, dtype=torch.int64)
                h = np.zeros(o.dimension())
                h[idx.cpu().numpy()] = x[idx].cpu().numpy()
                _, jo = o.evaluate(h, pts, jacobians=True, e0=e0, e1=e1, nthreads=8)
                scale = max(1.0, np.abs(exact).max())
                row = {"probe": "dg4 jac accuracy", "e0": e0, "oracle_vs_exact": float(np.abs(jo[:, :, 0, :] - exact).max() / scale)}
                for grad, label in ((-1, "collocation"), (1, "differentiated_tables")):
                    L.dfx_set_kernel_path(capi.OP_SMEM_GRADIENT, grad)
                    _, j = u.evaluate(pts, jacobians=True, elem_begin=e0, elem_end=e1)
                    j = j.cpu().numpy()[:, :, 0, :]
                    row[label + "_vs_exact"] = float(np.abs(j - exact).max() / scale)
                    row[label + "_vs_oracle"] = float(np.abs(j - jo[:, :, 0, :]).max() / scale)
                L.dfx_set_kernel_path(capi.OP_SMEM_GRADIENT, -1)
                print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
