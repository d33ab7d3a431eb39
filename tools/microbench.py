#!/usr/bin/env python
"""Per-kernel timings (CUDA events, after warm-up) of the hot-path kernels on the BASELINE
configs; prints one JSON line per measurement.  Development tool, not the bench contract.

    python tools/microbench.py [--only c2,c3] [--reps 10]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import dune_functions_b200 as dfx  # noqa: E402
from dune_functions_b200 import _capi as capi  # noqa: E402

PEAK = 6543.4
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def report(name, cfg, ms, best, nbytes, extra=None):
    gbs = nbytes / (ms * 1e-3) / 1e9
    line = {"kernel": name, "config": cfg, "ms_median": round(ms, 4), "ms_min": round(best, 4),
            "alg_GB": round(nbytes / 1e9, 4), "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 4)}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    L = capi.lib()
    torch.cuda.set_device(0)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731

    def want(c):
        return only is None or c in only

    # torch fill / copy as bandwidth references
    if want("ref"):
        n = 135005697
        a = torch.empty(n, dtype=torch.float64, device="cuda")
        b2 = torch.empty(n, dtype=torch.float64, device="cuda")
        ms, best = timeit(lambda: a.fill_(1.0), args.reps)
        report("torch.fill_", "135M doubles", ms, best, n * 8)
        ms, best = timeit(lambda: b2.copy_(a), args.reps)
        report("torch.copy_", "135M doubles", ms, best, 2 * n * 8)
        del a, b2

    for cfg, n, k in (("c2", [256, 256, 256], 2), ("c5", [512, 512, 512], 2), ("q1_256", [256, 256, 256], 1),
                      ("q3_128", [128, 128, 128], 3)):
        if not want(cfg):
            continue
        basis = dfx.makeBasis(dfx.YaspGrid(n), dfx.lagrange(k))
        x = basis.newVector(0.0)
        nd = basis.dimension()
        for sep, label in ((True, "separable tables (opt-in)"), (False, "f at every node")):
            ms, best = timeit(lambda: dfx.interpolate(basis, x, dfx.gaussian(1.0), separable=sep), args.reps)
            report("interpolate exp(-|x|^2) [" + label + "]", cfg, ms, best, nd * 8, {"dofs_per_s": nd / ms * 1e3})
        ms, best = timeit(lambda: dfx.interpolate(basis, x, dfx.quadratic(0.5, [1, 2, 3], [1, 2, 3, 4, 5, 6])), args.reps)
        report("interpolate quadratic [f at every node]", cfg, ms, best, nd * 8, {"dofs_per_s": nd / ms * 1e3})
        m = 3 if k <= 2 else k + 1
        pts, _ = dfx.tensor_gauss(3, m)
        ne = basis.numElements()
        f = dfx.makeDiscreteGlobalBasisFunction(basis, x)
        v = torch.empty((ne, len(pts), 1), dtype=torch.float64, device="cuda")
        ms, best = timeit(lambda: f.evaluate(pts, out_values=v), args.reps)
        report(f"evaluate values {m}^3", cfg, ms, best, nd * 8 + v.numel() * 8,
               {"path": L.dfx_evaluate_path(basis._h, 0, pts.ctypes.data_as(C.POINTER(C.c_double)), len(pts), 0)})
        if cfg in ("c2", "q1_256"):
            j = torch.empty((ne, len(pts), 1, 3), dtype=torch.float64, device="cuda")
            ms, best = timeit(lambda: f.evaluate(pts, jacobians=True, out_values=v, out_jacobians=j), args.reps)
            report(f"evaluate values+jacobians {m}^3", cfg, ms, best, nd * 8 + v.numel() * 8 * 4)
            del j
        if cfg == "c2":
            # interpolate(basis, y, makeDiscreteGlobalBasisFunction(basis, x)): evaluate at the nodes + owner scatter
            y2 = basis.newVector(0.0)
            ms, best = timeit(lambda: dfx.interpolate(basis, y2, f), max(2, args.reps // 2), warm=1)
            report("interpolate a DiscreteGlobalBasisFunction (Q2 -> Q2)", cfg, ms, best, 2 * nd * 8 + 2 * ne * 27 * 8,
                   {"dofs_per_s": nd / ms * 1e3})
            del y2
            idx = torch.empty((ne, 27), dtype=torch.int32, device="cuda")
            ms, best = timeit(lambda: capi.check(L.dfx_indices_flat(basis._h, 0, ne, C.c_void_p(idx.data_ptr()), st())), args.reps)
            report("indices flat u32", cfg, ms, best, idx.numel() * 4)
            del idx
        del basis, x, v, f
        torch.cuda.empty_cache()

    # Q3 with permuted face DOFs (random vertex ids) runs on the general kernels;
    # compare with the "q3_128" lines above (same basis, identity orientations, fast paths)
    if want("q3_128_oriented"):
        n = [128, 128, 128]
        nv = 129 ** 3
        ids = torch.randperm(nv, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        basis = dfx.makeBasis(dfx.YaspGrid(n), dfx.lagrange(3))
        basis.setVertexIds(ids)
        assert basis.hasPermutedFaceDOFs()
        nd, ne = basis.dimension(), basis.numElements()
        x = basis.newVector(0.0)
        ms, best = timeit(lambda: dfx.interpolate(basis, x, dfx.gaussian(1.0)), args.reps)
        report("interpolate exp(-|x|^2) [permuted face DOFs: run sweep + orientation bytes]", "q3_128_oriented", ms, best, nd * 8,
               {"dofs_per_s": nd / ms * 1e3})
        pts, _ = dfx.tensor_gauss(3, 4)
        f = dfx.makeDiscreteGlobalBasisFunction(basis, x)
        v = torch.empty((ne, len(pts), 1), dtype=torch.float64, device="cuda")
        ms, best = timeit(lambda: f.evaluate(pts, out_values=v), args.reps)
        report("evaluate values 4^3 [permuted face DOFs: k_eval_q3 + orientation bytes]", "q3_128_oriented", ms, best, nd * 8 + v.numel() * 8)
        idx = torch.empty((ne, 64), dtype=torch.int32, device="cuda")
        ms, best = timeit(lambda: capi.check(L.dfx_indices_flat(basis._h, 0, ne, C.c_void_p(idx.data_ptr()), st())), args.reps)
        report("indices flat u32 [permuted face DOFs: walking kernel + orientation bytes]", "q3_128_oriented", ms, best, idx.numel() * 4)
        del basis, x, v, f, idx, ids
        torch.cuda.empty_cache()

    if want("c3"):
        n = [128, 128, 128]
        basis = dfx.makeBasis(dfx.YaspGrid(n), dfx.taylorHood())
        ne, nl = basis.numElements(), basis.maxNodeSize()
        idx = torch.empty((ne, nl), dtype=torch.int32, device="cuda")
        ms, best = timeit(lambda: capi.check(L.dfx_indices_flat(basis._h, 0, ne, C.c_void_p(idx.data_ptr()), st())), args.reps)
        report("indices flat u32 (Taylor-Hood 89/elem)", "c3", ms, best, idx.numel() * 4, {"indices_per_s": idx.numel() / ms * 1e3})
        mi = torch.empty((ne, nl, 2), dtype=torch.int32, device="cuda")
        ms, best = timeit(lambda: capi.check(L.dfx_indices_multi(basis._h, 0, ne, C.c_void_p(mi.data_ptr()), 2, st())), args.reps)
        report("indices multi u32x2 (Taylor-Hood)", "c3", ms, best, mi.numel() * 4)
        del idx, mi
        x = basis.newVector(0.0)
        vel, pre = dfx.subspaceBasis(basis, 0), dfx.subspaceBasis(basis, 1)

        def interp():
            dfx.interpolate(vel, x, dfx.identity(3))
            dfx.interpolate(pre, x, dfx.gaussian(1.0))
        ms, best = timeit(interp, args.reps)
        report("interpolate velocity x + pressure exp [two calls]", "c3", ms, best, basis.dimension() * 8)
        ms, best = timeit(lambda: dfx.interpolateBatch(basis, x, [(vel, dfx.identity(3)), (pre, dfx.gaussian(1.0))]), args.reps)
        report("interpolate velocity x + pressure exp [dfx_interpolate_batch, one launch]", "c3", ms, best, basis.dimension() * 8)
        fw = dfx.makeDiscreteGlobalBasisFunction(basis, x)
        v4 = torch.empty((ne, 27, 4), dtype=torch.float64, device="cuda")
        pts3, _ = dfx.tensor_gauss(3, 3)
        ms, best = timeit(lambda: fw.evaluate(pts3, out_values=v4), args.reps)
        report("evaluate whole tree (3 x Q2 + Q1) values 3^3 [one launch, whole points]", "c3", ms, best, basis.dimension() * 8 + v4.numel() * 8)
        del v4, fw
        pts, _ = dfx.tensor_gauss(3, 3)
        fv = dfx.makeDiscreteGlobalBasisFunction(vel, x)
        v = torch.empty((ne, 27, 3), dtype=torch.float64, device="cuda")
        ms, best = timeit(lambda: fv.evaluate(pts, out_values=v), args.reps)
        report("evaluate velocity values 3^3", "c3", ms, best, 3 * 257 ** 3 * 8 + v.numel() * 8)
        del basis, x, v
        torch.cuda.empty_cache()

    if want("c4"):
        n = [192, 192, 192]
        basis = dfx.makeBasis(dfx.YaspGrid(n), dfx.lagrangeDG(4))
        x = basis.newVector(0.0)
        nd, ne = basis.dimension(), basis.numElements()
        for sep, label in ((True, "separable tables (opt-in)"), (False, "f at every node")):
            ms, best = timeit(lambda: dfx.interpolate(basis, x, dfx.gaussian(1.0), separable=sep), args.reps)
            report("interpolate DG4 exp(-|x|^2) [" + label + "]", "c4", ms, best, nd * 8)
        x.uniform_(-1, 1)
        pts, _ = dfx.tensor_gauss(3, 5)
        f = dfx.makeDiscreteGlobalBasisFunction(basis, x)
        v = torch.empty((ne, 125, 1), dtype=torch.float64, device="cuda")
        j = torch.empty((ne, 125, 1, 3), dtype=torch.float64, device="cuda")
        path = L.dfx_evaluate_path(basis._h, 0, pts.ctypes.data_as(C.POINTER(C.c_double)), 125, 1)
        ms, best = timeit(lambda: f.evaluate(pts, out_values=v), max(3, args.reps // 3), warm=1)
        report("evaluate DG4 values 5^3", "c4", ms, best, nd * 8 + v.numel() * 8, {"path": path})
        ms, best = timeit(lambda: f.evaluate(pts, jacobians=True, out_values=v, out_jacobians=j), max(3, args.reps // 3), warm=1)
        report("evaluate DG4 values+jacobians 5^3", "c4", ms, best, nd * 8 + v.numel() * 8 * 4, {"path": path})
        # the same contraction done densely on the FP64 tensor cores (path 3) and by the generic kernel
        L.dfx_set_kernel_path(capi.OP_EVALUATE, 3)
        ms, best = timeit(lambda: f.evaluate(pts, jacobians=True, out_values=v, out_jacobians=j), 3, warm=1)
        flops = 2.0 * 125 * 125 * 4 * ne
        report("evaluate DG4 values+jacobians 5^3 [dense DMMA]", "c4", ms, best, nd * 8 + v.numel() * 8 * 4,
               {"path": 3, "TFLOPs_fp64": flops / ms / 1e9})
        ms, best = timeit(lambda: f.evaluate(pts, out_values=v), 3, warm=1)
        report("evaluate DG4 values 5^3 [dense DMMA]", "c4", ms, best, nd * 8 + v.numel() * 8,
               {"path": 3, "TFLOPs_fp64": 2.0 * 125 * 125 * ne / ms / 1e9})
        L.dfx_set_kernel_path(capi.OP_EVALUATE, -1)

    # the element-loop callers (examples/poisson-pq2.cc): pattern, Laplace assembly, SpMV, boundary mask
    if want("ops"):
        for cfg, n, k in (("ops_q2_96", [96, 96, 96], 2), ("ops_q1_256", [256, 256, 256], 1)):
            basis = dfx.makeBasis(dfx.YaspGrid(n), dfx.lagrange(k))
            nd = basis.dimension()
            pat = dfx.getOccupationPattern(basis)
            nnz = pat[1].numel()
            ms, best = timeit(lambda: dfx.getOccupationPattern(basis), max(2, args.reps // 3), warm=1)
            report("occupation pattern (CSR, sorted)", cfg, ms, best, nnz * 4 + (nd + 1) * 8 * 2, {"nnz": nnz, "rows": nd})
            A, rhs = dfx.assembleLaplaceMatrix(basis, dfx.gaussian(1.0), pattern=pat)
            ms, best = timeit(lambda: dfx.assembleLaplaceMatrix(basis, None, pattern=pat), max(2, args.reps // 3), warm=1)
            report("Laplace matrix values (owner-computes rows)", cfg, ms, best, nnz * 12 + (nd + 1) * 8, {"nnz_per_s": nnz / ms * 1e3})
            ms, best = timeit(lambda: dfx.assembleLaplaceMatrix(basis, dfx.gaussian(1.0), pattern=pat), max(2, args.reps // 3), warm=1)
            report("Laplace matrix + load vector", cfg, ms, best, nnz * 12 + (nd + 1) * 8 + nd * 8)
            x = basis.newVector(0.0).uniform_(-1, 1)
            y = torch.empty_like(x)
            ms, best = timeit(lambda: A.mv(x, y), args.reps)
            report("SpMV y = A x (CSR, warp per row)", cfg, ms, best, nnz * 12 + (nd + 1) * 8 + 2 * nd * 8,
                   {"GFLOPs": 2.0 * nnz / ms / 1e6})
            op = dfx.LaplaceOperator(basis)
            ne_ = basis.numElements()
            nl_ = basis.maxNodeSize()
            ms, best = timeit(lambda: op.mv(x, y), args.reps)
            report("matrix-free Laplace y = A x", cfg, ms, best, 2 * nd * 8 + 2 * ne_ * nl_ * 8,
                   {"dofs_per_s": nd / ms * 1e3, "alg_bytes_note": "x + y + element-local scratch written and read once"})
            del op
            mask = torch.zeros(nd, dtype=torch.uint8, device="cuda")
            ms, best = timeit(lambda: dfx.boundaryDOFs(basis, mask), args.reps)
            report("boundary DOF mask", cfg, ms, best, nd, {"dofs_per_s": nd / ms * 1e3})
            del basis, A, rhs, x, y, pat, mask
            torch.cuda.empty_cache()
        # matrix-free only: the size of BASELINE config 2 (no assembled matrix would fit)
        basis = dfx.makeBasis(dfx.YaspGrid([256, 256, 256]), dfx.lagrange(2))
        nd, ne_, nl_ = basis.dimension(), basis.numElements(), basis.maxNodeSize()
        x = basis.newVector(0.0).uniform_(-1, 1)
        y = torch.empty_like(x)
        op = dfx.LaplaceOperator(basis)
        ms, best = timeit(lambda: op.mv(x, y), args.reps)
        report("matrix-free Laplace y = A x", "ops_q2_256", ms, best, 2 * nd * 8 + 2 * ne_ * nl_ * 8, {"dofs_per_s": nd / ms * 1e3})
        dn = dfx.boundaryDOFs(basis)
        opd = dfx.LaplaceOperator(basis, dn)
        ms, best = timeit(lambda: opd.mv(x, y), args.reps)
        report("matrix-free Laplace with Dirichlet rows", "ops_q2_256", ms, best, 2 * nd * 8 + 2 * ne_ * nl_ * 8 + 2 * nd)
        del basis, x, y, op, opd, dn
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
