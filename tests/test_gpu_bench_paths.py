"""Oracle parity ON THE CODE PATHS THE BENCHMARK RUNS ("What's weak"): the small-grid
suite (tests/cases.py: <= 100 elements) always takes the partial-tile branch of the evaluation
kernels, never the `c0 += blockDim.x` loop of the row interpolation kernel, and nothing close to the
index ranges of the BASELINE sizes.  Here the CPU oracle is run on

  * whole grids with several full 128-element blocks and rows longer than a thread block, and
  * element sub-ranges (first / middle / last) of the BASELINE grids themselves — 256^3 and 512^3
    Q2, 192^3 DG4, 128^3 Taylor-Hood — with seeded random coefficients, values and Jacobians,

and compared with the CUDA path through the C ABI: indices bit-exact, values within 1e-12.  The
oracle ranks multi-indices in closed form (PreBasis::lower) and takes element ranges, so a slice of a
512^3 grid costs seconds."""
import ctypes as C

import numpy as np
import pytest

import dune_functions_b200 as dfx
from dune_functions_b200 import _capi as capi, tensor_gauss

import orc

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b):
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def eval_path(basis, pts, jac, node=0):
    p = np.ascontiguousarray(pts)
    return capi.lib().dfx_evaluate_path(basis._h, node, p.ctypes.data_as(C.POINTER(C.c_double)), len(p), 1 if jac else 0)


def ranges_of(ne, count):
    count = min(count, ne)
    assert count % 128 == 0
    return sorted({0, ((ne - count) // 2) // 128 * 128, ne - count}), count


def check_evaluate_ranges(space, ob, pts, count, jac, onode=0, seed=3, nthreads=8):
    """random coefficients; oracle contracts the same numbers on [e0, e0 + count).  `space`: the basis or a
    sub-space basis of it, `onode` the oracle's tree node of the same subtree"""
    import torch
    basis = space.root if isinstance(space, dfx.SubspaceBasis) else space
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    x = basis.newVector(0.0).uniform_(-1.0, 1.0, generator=g)
    f = dfx.makeDiscreteGlobalBasisFunction(space, x)
    starts, count = ranges_of(basis.numElements(), count)
    for e0 in starts:
        e1 = e0 + count
        idx = basis.flatIndices(e0, e1, dtype=torch.int64)
        idx_o = ob.flatIndices(e0, e1).astype(np.int64)
        assert np.array_equal(idx.cpu().numpy(), idx_o), f"index table differs on [{e0}, {e1})"
        h = np.zeros(ob.dimension())
        h[idx_o] = x[idx].cpu().numpy()
        got = f.evaluate(pts, values=True, jacobians=jac, elem_begin=e0, elem_end=e1)
        ref = ob.evaluate(h, pts, node=onode, values=True, jacobians=jac, e0=e0, e1=e1, nthreads=nthreads)
        if jac:
            assert rel(got[0].cpu().numpy(), ref[0]) <= TOL, (e0, "values")
            assert rel(got[1].cpu().numpy(), ref[1]) <= TOL, (e0, "jacobians")
        else:
            assert rel(got.cpu().numpy(), ref) <= TOL, (e0, "values")
        del got
    del x


def check_interpolate_ranges(basis, ob, fns, count, separable=False, nthreads=8):
    """fns: [(gpu node, oracle node, function)]"""
    import torch
    x = basis.newVector(float("nan"))
    L = capi.lib()
    entry = L.dfx_interpolate_separable if separable else L.dfx_interpolate
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for node, _, fn in fns:
        capi.check(entry(basis._h, node, C.byref(fn), C.c_void_p(x.data_ptr()), None, st))
    assert not bool(torch.isnan(x).any())
    starts, count = ranges_of(basis.numElements(), count)
    for e0 in starts:
        e1 = e0 + count
        idx = basis.flatIndices(e0, e1, dtype=torch.int64)
        idx_o = ob.flatIndices(e0, e1).astype(np.int64)
        c = np.zeros(ob.dimension())
        for _, onode, fn in fns:
            ob.interpolate(fn, coeff=c, node=onode, e0=e0, e1=e1, nthreads=nthreads)
        assert rel(x[idx].cpu().numpy(), c[idx_o]) <= TOL, (e0, separable)
    del x


# ---- whole grids: several full blocks, rows longer than a thread block -------------------------------
@pytest.mark.parametrize("n", [[16, 8, 8], [300, 3, 2], [130, 2, 5]])
@pytest.mark.parametrize("k", [1, 2])
def test_full_blocks_and_long_rows_against_oracle(n, k):
    import torch
    grid = dfx.YaspGrid(n, upper=[1.0, 0.75, 1.25])
    tree = dfx.lagrange(k)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    assert b.numElements() >= 4 * 128
    assert np.array_equal(b.flatIndices(dtype=torch.int64).cpu().numpy(), o.flatIndices().astype(np.int64))
    for f in (dfx.gaussian(1.0), dfx.quadratic(0.5, [1, -2, 3], [1, 0.5, -1, 2, 0.25, -0.5])):
        ref = o.interpolate(f)
        for sep in (False, True):
            x = dfx.interpolate(b, b.newVector(-3.0), f, separable=sep)
            assert rel(x.cpu().numpy(), ref) <= TOL, (n, k, f.id, sep)
    rng = np.random.default_rng(11)
    xr = rng.uniform(-1, 1, o.dimension())
    xt = torch.from_numpy(xr).cuda()
    for m in (2, 3):
        pts, _ = tensor_gauss(3, m)
        assert eval_path(b, pts, False) == 1 and eval_path(b, pts, True) == 1
        v, j = dfx.makeDiscreteGlobalBasisFunction(b, xt).evaluate(pts, jacobians=True)     # 64-element blocks, bulk
        vo = dfx.makeDiscreteGlobalBasisFunction(b, xt).evaluate(pts)                        # 32-element blocks (default), bulk
        for epb in (64, 128):                                                                # the other two tile sizes
            capi.lib().dfx_set_kernel_path(capi.OP_TUNE_EVAL_BLOCK, epb)
            try:
                ve = dfx.makeDiscreteGlobalBasisFunction(b, xt).evaluate(pts)
            finally:
                capi.lib().dfx_set_kernel_path(capi.OP_TUNE_EVAL_BLOCK, -1)
            assert torch.equal(ve, vo), (n, k, m, epb)
        vr, jr = o.evaluate(xr, pts, jacobians=True, nthreads=4)
        assert rel(v.cpu().numpy(), vr) <= TOL and rel(vo.cpu().numpy(), vr) <= TOL and rel(j.cpu().numpy(), jr) <= TOL
        # a sub-range that starts inside a block and ends inside another: partial tiles at both ends
        e0, e1 = 77, b.numElements() - 51
        vs = dfx.makeDiscreteGlobalBasisFunction(b, xt).evaluate(pts, elem_begin=e0, elem_end=e1)
        assert rel(vs.cpu().numpy(), vr[e0:e1]) <= TOL


# ---- BASELINE grids, element sub-ranges -------------------------------------------------------------
@pytest.mark.parametrize("n", [256, 512])
def test_q2_baseline_grids_subranges(n):
    """configs[1] / configs[4]: k_eval_reg<3,2,3,*,*> on full 128-element blocks (TMA bulk store branch:
    range multiple of 128, fresh 512-byte aligned output) and k_interp_rows with N_x > blockDim"""
    grid = dfx.YaspGrid([n, n, n])
    tree = dfx.lagrange(2)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    assert b.dimension() == (2 * n + 1) ** 3 == o.dimension()
    pts, _ = tensor_gauss(3, 3)
    assert eval_path(b, pts, False) == 1 and eval_path(b, pts, True) == 1
    check_evaluate_ranges(b, o, pts, 1 << 17, jac=False)
    check_evaluate_ranges(b, o, pts, 1 << 15, jac=True, seed=4)
    for sep in (False, True):
        check_interpolate_ranges(b, o, [(0, 0, dfx.gaussian(1.0))], 1 << 16, separable=sep)


@pytest.mark.parametrize("n", [[70, 3, 2], [16, 8, 8], [5, 4, 3]])
@pytest.mark.parametrize("last_fastest", [False, True])
def test_q3_register_kernel_against_oracle(n, last_fastest):
    """continuous order 3 at 4^3 points, values only: k_eval_q3 (thread per element, 64 coefficients in registers,
    staged copy-out) — whole grids with full and partial 64-element blocks, sub-ranges with odd starts, both point
    orders, a sub-space of a power tree (position stride 2) and the shared-memory kernel on the same numbers"""
    import torch
    L = capi.lib()
    grid = dfx.YaspGrid(n, upper=[1.0, 0.75, 1.25])
    pts, _ = tensor_gauss(3, 4, last_fastest)
    rng = np.random.default_rng(5)
    for tree, node, onode in ((dfx.lagrange(3), 0, 0), (dfx.power(dfx.lagrange(3), 2, dfx.flatInterleaved()), None, None)):
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        xr = rng.uniform(-1, 1, o.dimension())
        xt = torch.from_numpy(xr).cuda()
        ne = b.numElements()
        if node is None:
            space, onode = dfx.subspaceBasis(b, 1), b.child(0, 1)
        else:
            space = b
        f = dfx.makeDiscreteGlobalBasisFunction(space, xt)
        ref = o.evaluate(xr, pts, node=onode, nthreads=4)
        try:
            for q3 in (1, 0):
                L.dfx_set_kernel_path(capi.OP_TUNE_EVAL_Q3, q3)
                assert rel(f.evaluate(pts).cpu().numpy(), ref) <= TOL, (n, q3)
                for e0, e1 in ((1, ne), (3, ne - 2), (ne // 2, ne // 2 + 1)):
                    assert rel(f.evaluate(pts, elem_begin=e0, elem_end=e1).cpu().numpy(), ref[e0:e1]) <= TOL, (n, q3, e0)
        finally:
            L.dfx_set_kernel_path(capi.OP_TUNE_EVAL_Q3, -1)


def test_dg4_192_subranges():
    """configs[3]: k_eval_smem<5,5,0,TMA> (values) and k_eval_pair<5,5,...> (values + collocation gradients; the
    8 pairs of a block leave as one block tile per array by TMA bulk stores); k_interp_cell"""
    grid = dfx.YaspGrid([192, 192, 192])
    tree = dfx.lagrangeDG(4)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    pts, _ = tensor_gauss(3, 5)
    assert eval_path(b, pts, False) == 2 and eval_path(b, pts, True) == 2
    check_evaluate_ranges(b, o, pts, 4096, jac=False)
    check_evaluate_ranges(b, o, pts, 2048 + 128, jac=True, seed=5)
    quad = dfx.quadratic(0.0, [0, 0, 0], [42, 7, 0, 13, 0, 5])
    check_interpolate_ranges(b, o, [(0, 0, quad)], 8192)
    check_interpolate_ranges(b, o, [(0, 0, dfx.gaussian(1.0))], 8192)
    check_interpolate_ranges(b, o, [(0, 0, dfx.gaussian(1.0))], 8192, separable=True)


def test_taylor_hood_128_subranges():
    """configs[2]: flat and two-digit index tables, velocity + pressure interpolation, whole-tree and
    per-subspace evaluation on slices of the 128^3 grid"""
    import torch
    grid = dfx.YaspGrid([128, 128, 128])
    tree = dfx.taylorHood()
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    assert b.dimension() == o.dimension() == 3 * 257 ** 3 + 129 ** 3
    starts, count = ranges_of(b.numElements(), 8192)
    for e0 in starts:
        e1 = e0 + count
        dig, lens = o.indices(e0, e1)
        assert np.array_equal(b.indices(e0, e1).cpu().numpy().astype(np.uint64), dig)
        assert np.array_equal(b.flatIndices(e0, e1, dtype=torch.int64).cpu().numpy(), o.flatIndices(e0, e1).astype(np.int64))
        assert (lens == 2).all()
    vel, pre = b.child(0, 0), b.child(0, 1)
    fns = [(vel, 1, dfx.identity(3)), (pre, 5, dfx.gaussian(1.0))]
    check_interpolate_ranges(b, o, fns, 8192)
    check_interpolate_ranges(b, o, fns, 8192, separable=True)
    pts, _ = tensor_gauss(3, 3)
    check_evaluate_ranges(b, o, pts, 4096, jac=False)                            # whole tree: Q2 x 3 + Q1, warps per leaf group
    check_evaluate_ranges(b, o, pts, 2048, jac=True, seed=6)
    capi.lib().dfx_set_kernel_path(capi.OP_TUNE_SPLIT_MAP, 0)                    # every warp runs the Q2 and the Q1 branch
    try:
        check_evaluate_ranges(b, o, pts, 4096, jac=False)
        check_evaluate_ranges(b, o, pts, 2048, jac=True, seed=6)
    finally:
        capi.lib().dfx_set_kernel_path(capi.OP_TUNE_SPLIT_MAP, -1)
    check_evaluate_ranges(dfx.subspaceBasis(b, 0), o, pts, 4096, jac=False, onode=1, seed=7)  # velocity sub-space
    check_evaluate_ranges(dfx.subspaceBasis(b, 1), o, pts, 4096, jac=False, onode=5, seed=8)  # pressure sub-space


# ---- batched interpolation (dfx_interpolate_batch) and the thread-per-(element, leaf) evaluation --------------
@pytest.mark.parametrize("n", [[5, 4, 3], [40, 9, 7], [300, 3, 2]])
def test_interpolate_batch_equals_the_sequence_of_calls(n):
    """one fused launch for two lean lattice sweeps (interleaved velocity + scalar pressure), the plain sequence
    for everything else — bit-identical to calling interpolate part by part, and equal to the oracle"""
    from cases import th_variant
    from dune_functions_b200 import (blockedLexicographic as BL, flatLexicographic as FL, flatInterleaved as FI,
                                     blockedInterleaved as BI)
    grid = dfx.YaspGrid(n, upper=[1.0, 0.5, 2.0])
    trees = {"taylorHood": dfx.taylorHood()}
    for on, o_ in (("BL", BL), ("FL", FL)):
        for inn, i_ in (("FI", FI), ("BI", BI), ("BL", BL), ("FL", FL)):
            trees[f"TH_{on}({inn})"] = th_variant(o_, i_, 3)
    for name, tree in trees.items():
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        vel, pre = dfx.subspaceBasis(b, 0), dfx.subspaceBasis(b, 1)
        for fv, fp in ((dfx.identity(3), dfx.gaussian(1.0)), (dfx.gaussian_vec(3, 0.5), dfx.quadratic(0.5, [1, -2, 3], [1, 0.5, -1, 2, 0.25, -0.5]))):
            x = dfx.interpolateBatch(b, b.newVector(-7.0), [(vel, fv), (pre, fp)])
            y = b.newVector(-7.0)
            dfx.interpolate(vel, y, fv)
            dfx.interpolate(pre, y, fp)
            assert bool((x == y).all()), name
            ref = np.full(o.dimension(), -7.0)
            o.interpolate(fv, coeff=ref, node=1)
            o.interpolate(fp, coeff=ref, node=o.numTreeNodes() - 1)
            assert rel(x.cpu().numpy(), ref) <= TOL, name
        # overlapping parts keep their order: the whole tree first, then the pressure again
        x = dfx.interpolateBatch(b, b.newVector(0.0), [(b, dfx.constant(1.0, 2.0, 3.0, 4.0)), (pre, dfx.constant(9.0))])
        ref = np.zeros(o.dimension())
        o.interpolate(dfx.constant(1.0, 2.0, 3.0, 4.0), coeff=ref)
        o.interpolate(dfx.constant(9.0), coeff=ref, node=o.numTreeNodes() - 1)
        assert np.array_equal(x.cpu().numpy(), ref), name


@pytest.mark.parametrize("n", [[16, 8, 8], [33, 5, 4]])
def test_multi_leaf_evaluation_whole_points(n):
    """power / composite subtrees go through the thread-per-(element, leaf) kernel: full blocks (bulk store) and
    ragged ends, values and Jacobians, every Taylor-Hood merging variant and plain power nodes"""
    import torch
    from cases import trees as case_trees
    grid = dfx.YaspGrid(n, upper=[1.0, 0.75, 1.25])
    rng = np.random.default_rng(21)
    for name, tree in case_trees(3).items():
        if not (name.startswith("TH_") or name in ("taylorHood", "power3_q2_BI", "power2_q2_BL", "nested_pow_of_comp")):
            continue
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        xr = rng.uniform(-1, 1, o.dimension())
        xt = torch.from_numpy(xr).cuda()
        for m in (2, 3):
            pts, _ = tensor_gauss(3, m)
            v, j = dfx.makeDiscreteGlobalBasisFunction(b, xt).evaluate(pts, jacobians=True)
            vr, jr = o.evaluate(xr, pts, jacobians=True, nthreads=4)
            assert rel(v.cpu().numpy(), vr) <= TOL and rel(j.cpu().numpy(), jr) <= TOL, (name, m)
            e0, e1 = 3, b.numElements() - 5
            vs = dfx.makeDiscreteGlobalBasisFunction(b, xt).evaluate(pts, elem_begin=e0, elem_end=e1)
            assert rel(vs.cpu().numpy(), vr[e0:e1]) <= TOL, (name, m)


@pytest.mark.parametrize("n", [[128, 70, 3], [256, 65, 2]])
def test_strip_schedule_of_the_evaluation_blocks(n):
    """k_eval_reg walks big grids strip by strip (32 element rows through all z-layers) so that z-neighbours
    share coefficients out of the L2; forced here on grids the oracle runs whole: full strips, a lower last strip,
    values (128-element blocks) and Jacobians (64-element blocks), whole range and a range of whole layers"""
    import torch
    L = capi.lib()
    grid = dfx.YaspGrid(n, upper=[1.0, 0.75, 1.25])
    tree = dfx.lagrange(2)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    xr = np.random.default_rng(5).uniform(-1, 1, o.dimension())
    xt = torch.from_numpy(xr).cuda()
    pts, _ = tensor_gauss(3, 3)
    vr, jr = o.evaluate(xr, pts, jacobians=True, nthreads=8)
    layer = n[0] * n[1]
    try:
        for mode in (2, 0):
            L.dfx_set_kernel_path(capi.OP_TUNE_SWIZZLE, mode)
            f = dfx.makeDiscreteGlobalBasisFunction(b, xt)
            v = f.evaluate(pts)
            assert rel(v.cpu().numpy(), vr) <= TOL, mode
            v, j = f.evaluate(pts, jacobians=True)
            assert rel(v.cpu().numpy(), vr) <= TOL and rel(j.cpu().numpy(), jr) <= TOL, mode
            if n[2] > 2:
                v = f.evaluate(pts, elem_begin=layer, elem_end=3 * layer)
                assert rel(v.cpu().numpy(), vr[layer:3 * layer]) <= TOL, mode
    finally:
        L.dfx_set_kernel_path(capi.OP_TUNE_SWIZZLE, -1)
