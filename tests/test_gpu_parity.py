"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
inputs.  Bar: multi-indices / flat positions bit-exact; coefficients and evaluations
<= 1e-12 relative (|a-b| / max(1, max|ref|) per array)."""
import ctypes as C

import numpy as np
import pytest

import orc
from cases import GRIDS, trees, th_variant
import dune_functions_b200 as dfx
from dune_functions_b200 import _capi as capi
from dune_functions_b200 import (YaspGrid, lagrange, lagrangeDG, power, composite, taylorHood, tensor_gauss,
                                 flatLexicographic as FL, flatInterleaved as FI,
                                 blockedLexicographic as BL, blockedInterleaved as BI)

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, ref):
    return np.abs(a - ref).max() / max(1.0, np.abs(ref).max())


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    yield
    capi.lib().dfx_set_kernel_path(capi.OP_EVALUATE, -1)


def all_cases():
    for gname, grid in GRIDS.items():
        for name, tree in trees(grid.dim).items():
            if "TH_" in name and grid.dim != 3:
                continue
            yield pytest.param(grid, tree, id=f"{gname}-{name}")


# ---- bind + index -----------------------------------------------------------------------
@pytest.mark.parametrize("grid,tree", list(all_cases()))
def test_indices_bit_exact(grid, tree):
    import torch
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    dig, lens = o.indices()
    got = b.indices().cpu().numpy().astype(np.uint64)
    assert (got == dig).all()
    got64 = b.indices(dtype=torch.int64).cpu().numpy().astype(np.uint64)
    assert (got64 == dig).all()
    assert (b.localIndexSizes() == lens[0]).all()
    flat = b.flatIndices().cpu().numpy().astype(np.uint64)
    assert (flat == o.flatIndices()).all()
    # an element sub-range
    ne = b.numElements()
    if ne > 2:
        sub = b.flatIndices(1, ne - 1).cpu().numpy().astype(np.uint64)
        assert (sub == flat[1:ne - 1]).all()


def test_indices_host_entry_point():
    grid, tree = GRIDS["3d_3x4x5"], taylorHood()
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    out = np.zeros((b.numElements(), b.maxNodeSize()), dtype=np.uint32)
    capi.check(capi.lib().dfx_indices_flat_host(b._h, 0, b.numElements(), out.ctypes.data))
    assert (out.astype(np.uint64) == o.flatIndices()).all()


# ---- interpolate ----------------------------------------------------------------------------
FUNCS = {
    "gaussian": dfx.gaussian(1.0),
    "coord0": dfx.coord(0),
    "quadratic": dfx.quadratic(0.5, [1, -2, 3], [42, 7, 1, 13, -3, 5]),
    "sinprod": dfx.sinprod(3.0),
    "const": dfx.constant(1.0),
}


@pytest.mark.parametrize("grid,tree", list(all_cases()))
@pytest.mark.parametrize("fname", ["gaussian", "quadratic"])
def test_interpolate_scalar_function(grid, tree, fname):
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    f = FUNCS[fname]
    x = b.newVector(-5.0)
    dfx.interpolate(b, x, f)
    ref = o.interpolate(f, coeff=np.full(o.dimension(), -5.0))
    assert rel(x.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_2x1x3"])
@pytest.mark.parametrize("k", [0, 1, 2, 3, 5])
def test_interpolate_coordinate_is_bit_exact(gname, k):
    """x[0] involves no transcendental: the GPU reproduces the last writer's coordinate bit for bit"""
    grid = GRIDS[gname]
    for tree in (lagrange(k), lagrangeDG(k)):
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        for f in (dfx.coord(0), dfx.coord(grid.dim - 1)):
            x = b.newVector(0.0)
            dfx.interpolate(b, x, f)
            assert (x.cpu().numpy() == o.interpolate(f)).all()


def test_interpolate_vector_valued_and_subspaces():
    """examples/interpolation.cc:71-96: Taylor-Hood FL(FL), velocity subspace <- f(x) = x,
    pressure subspace <- exp(-|x|^2)"""
    grid = GRIDS["3d_3x4x5"]
    for tree in (th_variant(FL, FL, 3), th_variant(BL, BI, 3), taylorHood(), th_variant(FL, BI, 3)):
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        x = b.newVector(3.0)
        ref = np.full(o.dimension(), 3.0)
        vel = dfx.subspaceBasis(b, 0)
        pre = dfx.subspaceBasis(b, 1)
        dfx.interpolate(vel, x, dfx.identity(3))
        o.interpolate(dfx.identity(3), coeff=ref, node=vel.node)
        assert (x.cpu().numpy() == ref).all()
        dfx.interpolate(pre, x, dfx.gaussian(1.0))
        o.interpolate(dfx.gaussian(1.0), coeff=ref, node=pre.node)
        assert rel(x.cpu().numpy(), ref) <= TOL
        # single velocity component: subspaceBasis(basis, 0, 1)
        comp = dfx.subspaceBasis(b, 0, 1)
        dfx.interpolate(comp, x, dfx.sinprod(2.0))
        o.interpolate(dfx.sinprod(2.0), coeff=ref, node=comp.node)
        assert rel(x.cpu().numpy(), ref) <= TOL
        # whole tree with a 4-component range
        dfx.interpolate(b, x, dfx.gaussian_vec(4, 0.5))
        o.interpolate(dfx.gaussian_vec(4, 0.5), coeff=ref)
        assert rel(x.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("gname", ["2d_3x2", "3d_2x1x3", "3d_3x4x5"])
@pytest.mark.parametrize("path", [-1, 1, 0])
def test_interpolate_vector_valued_separable_tables(kernel_path, gname, path):
    """vector-valued registry functions whose range entries are products of per-direction factors
    (x -> x, x -> x exp(-a|x|^2), constants) go through one table set per range entry; every kernel
    family gives the oracle's answer (bit-exact for x -> x and constants)"""
    grid = GRIDS[gname]
    d = grid.dim
    kernel_path(capi.OP_INTERPOLATE, path)
    for tname in ("power3_q2_BI", "power2_q1_FI", "power2_q2_BL", "power3_dg1_FL", "nested_FL_FI_BL"):
        tree = trees(d)[tname]
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        nl = b.numLeaves()
        for f, exact in ((dfx.identity(nl), True), (dfx.constant(*[1.5 * c - 2 for c in range(nl)]), True),
                         (dfx.gaussian_vec(nl, 0.7), False)):
            x = b.newVector(-9.0)
            dfx.interpolate(b, x, f)
            ref = o.interpolate(f, coeff=np.full(o.dimension(), -9.0))
            if exact:
                assert np.array_equal(x.cpu().numpy(), ref), (tname, f.id)
            else:
                assert rel(x.cpu().numpy(), ref) <= TOL, (tname, f.id)


def test_interpolate_masked():
    import torch
    grid = GRIDS["2d_10x10"]
    tree = power(lagrange(2), 2, BI())
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    rng = np.random.default_rng(1)
    mask = (rng.uniform(size=b.dimension()) < 0.3).astype(np.uint8)
    x = b.newVector(7.0)
    dfx.interpolate(b, x, dfx.identity(2), mask=torch.from_numpy(mask).cuda())
    ref = o.interpolate(dfx.identity(2), coeff=np.full(o.dimension(), 7.0), mask=mask)
    got = x.cpu().numpy()
    assert (got == ref).all()
    assert (got[mask == 0] == 7.0).all()


def test_interpolate_wrong_range_raises():
    b = dfx.makeBasis(GRIDS["2d_3x2"], power(lagrange(1), 2, FI()))
    with pytest.raises(capi.RangeError):
        dfx.interpolate(b, b.newVector(0.0), dfx.identity(3))
    import torch
    with pytest.raises(capi.RangeError):
        dfx.interpolate(b, torch.zeros(3, dtype=torch.float64, device="cuda"), dfx.identity(2))


def test_interpolation_points_and_from_values():
    """form (2) of f: node positions out, caller-evaluated values in"""
    grid = GRIDS["3d_2x1x3"]
    tree = composite(power(lagrange(2), 3, BI()), lagrangeDG(1))
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    xyz, leaf = dfx.interpolationPoints(b)
    xyz, leaf = xyz.cpu().numpy(), leaf.cpu().numpy()
    # f(x) = (x0, x1, x2, exp(-|x|^2)) evaluated by the caller
    vals = np.where(leaf < 3, xyz[np.arange(len(leaf)), np.minimum(leaf, 2)], np.exp(-(xyz ** 2).sum(1)))
    import torch
    x = b.newVector(0.0)
    dfx.interpolateFromValues(b, x, torch.from_numpy(vals).cuda())
    ref = np.zeros(o.dimension())
    o.interpolate(dfx.identity(3), coeff=ref, node=1)
    o.interpolate(dfx.gaussian(1.0), coeff=ref, node=b.child(0, 1))
    assert rel(x.cpu().numpy(), ref) <= TOL


def test_interpolate_host_entry_point():
    grid, tree = GRIDS["3d_3x4x5"], lagrange(2)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    out = np.zeros(b.dimension())
    f = dfx.gaussian(1.0)
    capi.check(capi.lib().dfx_interpolate_host(b._h, 0, C.byref(f), out.ctypes.data, None))
    assert rel(out, o.interpolate(f)) <= TOL
    # masked: untouched entries survive the round trip through the device
    mask = (np.arange(b.dimension()) % 3 == 0).astype(np.uint8)
    out2 = np.full(b.dimension(), 4.0)
    capi.check(capi.lib().dfx_interpolate_host(b._h, 0, C.byref(f), out2.ctypes.data, mask.ctypes.data))
    ref2 = o.interpolate(f, coeff=np.full(b.dimension(), 4.0), mask=mask)
    assert rel(out2, ref2) <= TOL and (out2[mask == 0] == 4.0).all()


# ---- DiscreteGlobalBasisFunction evaluation ---------------------------------------------------
def random_coeff(n, seed=12345):
    return np.random.default_rng(seed).uniform(-1, 1, n)


def eval_both(grid, tree, pts, node=0, jac=True, path=-1):
    import torch
    capi.lib().dfx_set_kernel_path(capi.OP_EVALUATE, path)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    c = random_coeff(b.dimension())
    f = dfx.makeDiscreteGlobalBasisFunction(b if node == 0 else _Sub(b, node), torch.from_numpy(c).cuda())
    v, j = f.evaluate(pts, jacobians=True) if jac else (f.evaluate(pts), None)
    vr, jr = o.evaluate(c, pts, node=node, jacobians=True) if jac else (o.evaluate(c, pts, node=node), None)
    capi.lib().dfx_set_kernel_path(capi.OP_EVALUATE, -1)
    return v.cpu().numpy(), (j.cpu().numpy() if jac else None), vr, jr


class _Sub(dfx.SubspaceBasis):
    def __init__(self, root, node):
        self.root, self.node, self.path = root, node, ()


@pytest.mark.parametrize("grid,tree", list(all_cases()))
def test_evaluate_generic_random_points(grid, tree):
    """arbitrary (non tensor-product) points -> generic kernel"""
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 1, (5, grid.dim))
    v, j, vr, jr = eval_both(grid, tree, pts, path=0)
    assert rel(v, vr) <= TOL
    assert rel(j, jr) <= TOL


@pytest.mark.parametrize("grid,tree", list(all_cases()))
@pytest.mark.parametrize("m,last_fastest", [(3, True), (2, False)])
def test_evaluate_tensor_gauss_points(grid, tree, m, last_fastest):
    """tensor Gauss points -> whichever path the planner picks"""
    pts, _ = tensor_gauss(grid.dim, m, last_fastest)
    v, j, vr, jr = eval_both(grid, tree, pts)
    assert rel(v, vr) <= TOL
    assert rel(j, jr) <= TOL


def test_evaluate_subspace_and_element_range():
    import torch
    grid = GRIDS["3d_3x4x5"]
    tree = taylorHood()
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    c = random_coeff(b.dimension())
    ct = torch.from_numpy(c).cuda()
    pts, _ = tensor_gauss(3, 3)
    for path in ((0,), (1,), (0, 2)):
        sb = dfx.subspaceBasis(b, *path)
        f = dfx.makeDiscreteGlobalBasisFunction(sb, ct)
        v = f.evaluate(pts, elem_begin=7, elem_end=41).cpu().numpy()
        vr = o.evaluate(c, pts, node=sb.node, e0=7, e1=41)
        assert v.shape == vr.shape and rel(v, vr) <= TOL


def test_evaluate_host_entry_point():
    grid, tree = GRIDS["3d_3x4x5"], lagrange(2)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    c = random_coeff(b.dimension())
    pts, _ = tensor_gauss(3, 3)
    ne = b.numElements()
    v = np.zeros((ne, 27, 1))
    j = np.zeros((ne, 27, 1, 3))
    capi.check(capi.lib().dfx_evaluate_host(b._h, 0, c.ctypes.data, pts.ctypes.data_as(C.POINTER(C.c_double)), 27,
                                            0, ne, v.ctypes.data, j.ctypes.data))
    vr, jr = o.evaluate(c, pts, jacobians=True)
    assert rel(v, vr) <= TOL and rel(j, jr) <= TOL


@pytest.mark.parametrize("gname,tname", [("3d_3x4x5", "lagrange2"), ("3d_3x4x5", "taylorHood"), ("3d_3x4x5", "TH_FL(FI)"),
                                         ("3d_3x4x5", "TH_BL(BI)"), ("3d_3x4x5", "lagrangeDG2"), ("3d_2x1x3", "lagrange3"),
                                         ("3d_3x4x5", "nested_FL_FI_BL"), ("2d_10x10", "power3_q2_BI"),
                                         ("2d_10x10", "lagrange1"), ("1d_5", "lagrange4")])
@pytest.mark.parametrize("band_kib", [1, 7, 64])
def test_evaluate_host_pipeline_bands(gname, tname, band_kib):
    """many small bands: every band must find the coefficients it gathers on the device
    (contiguous and interleaved layouts, sub-ranges that start and end inside a layer)"""
    grid = GRIDS[gname]
    tree = trees(grid.dim)[tname]
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    c = random_coeff(b.dimension())
    d = grid.dim
    pts, _ = tensor_gauss(d, 2)
    nq, nc, ne = len(pts), b.numLeaves(), b.numElements()
    L = capi.lib()
    L.dfx_set_kernel_path(capi.OP_HOST_BAND_KIB, band_kib)
    try:
        for e0, e1 in ((0, ne), (1, ne - 1), (ne // 2, ne // 2 + 1)):
            v = np.full((e1 - e0, nq, nc), np.nan)
            j = np.full((e1 - e0, nq, nc, d), np.nan)
            capi.check(L.dfx_evaluate_host(b._h, 0, c.ctypes.data, pts.ctypes.data_as(C.POINTER(C.c_double)), nq,
                                           e0, e1, v.ctypes.data, j.ctypes.data))
            vr, jr = o.evaluate(c, pts, jacobians=True, e0=e0, e1=e1)
            assert rel(v, vr) <= TOL and rel(j, jr) <= TOL
    finally:
        L.dfx_set_kernel_path(capi.OP_HOST_BAND_KIB, -1)


# ---- interpolate(basis, coeff, gridFunction) -----------------------------------------------------
@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_3x4x5", "3d_2x1x3"])
def test_interpolation_consistency_of_discrete_functions(gname):
    """checkInterpolationConsistency, discreteglobalbasisfunctiontest.cc:49-77: interpolating
    makeDiscreteGlobalBasisFunction(basis, x) into the same basis returns x (the reference asks
    for 1e-10; a nodal basis reproduces its own coefficients to rounding)"""
    import torch
    grid = GRIDS[gname]
    for tname, tree in trees(grid.dim).items():
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        x = random_coeff(b.dimension())
        xt = torch.from_numpy(x).cuda()
        y = b.newVector(float("nan"))
        dfx.interpolate(b, y, dfx.makeDiscreteGlobalBasisFunction(b, xt))
        got = y.cpu().numpy()
        assert np.abs(got - x).max() <= 1e-10, tname
        assert rel(got, o.interpolateGridFunction(o, x)) <= TOL, tname


@pytest.mark.parametrize("src,dst", [("lagrange2", "lagrange1"), ("lagrange1", "lagrange2"), ("lagrange3", "lagrangeDG2"),
                                     ("lagrangeDG1", "lagrange2"), ("lagrange2", "lagrange0"), ("lagrange4", "lagrange3"),
                                     ("power3_q2_BI", "power3_dg1_FL"), ("lagrange2", "power3_q2_BI")])
def test_interpolate_grid_function_between_bases(src, dst):
    """a discrete function of one basis interpolated into another basis on the same grid (restriction,
    injection, DG <-> continuous: last element in iteration order wins on shared nodes), small bands"""
    import torch
    grid = GRIDS["3d_3x4x5"]
    ts, td = trees(3)[src], trees(3)[dst]
    bs, bd = dfx.makeBasis(grid, ts), dfx.makeBasis(grid, td)
    os_, od = orc.OracleBasis(grid, ts), orc.OracleBasis(grid, td)
    xs = random_coeff(bs.dimension())
    f = dfx.makeDiscreteGlobalBasisFunction(bs, torch.from_numpy(xs).cuda())
    y = bd.newVector(-7.0)
    dfx.interpolate(bd, y, f)
    exp = od.interpolateGridFunction(os_, xs, coeff=np.full(od.dimension(), -7.0))
    assert rel(y.cpu().numpy(), exp) <= TOL
    # element loop in bands of one z-layer (the smallest scratch the entry point accepts) and a mask
    L = capi.lib()
    layer = 3 * 4
    nl = max(od.nodeInfo(0)["size"], 1)
    ncomp = os_.numLeaves()
    scratch = torch.empty(layer * bd.maxNodeSize() * ncomp, dtype=torch.float64, device="cuda")
    mask = (np.arange(bd.dimension()) % 3 != 1).astype(np.uint8)
    y2 = bd.newVector(-7.0)
    capi.check(L.dfx_interpolate_dgbf(bd._h, 0, bs._h, 0, C.c_void_p(f.coeff.data_ptr()), C.c_void_p(y2.data_ptr()),
                                      C.c_void_p(torch.from_numpy(mask).cuda().data_ptr()), C.c_void_p(scratch.data_ptr()),
                                      scratch.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    exp2 = od.interpolateGridFunction(os_, xs, coeff=np.full(od.dimension(), -7.0), mask=mask)
    assert rel(y2.cpu().numpy(), exp2) <= TOL
    del nl
    # host entry point
    yh = np.full(bd.dimension(), -7.0)
    capi.check(L.dfx_interpolate_dgbf_host(bd._h, 0, bs._h, 0, xs.ctypes.data, yh.ctypes.data, None))
    assert rel(yh, exp) <= TOL


def test_interpolate_grid_function_errors():
    import torch
    b3 = dfx.makeBasis(GRIDS["3d_3x4x5"], lagrange(1))
    other = dfx.makeBasis(GRIDS["3d_2x1x3"], lagrange(1))
    f = dfx.makeDiscreteGlobalBasisFunction(other, other.newVector(1.0))
    with pytest.raises(capi.RangeError):
        dfx.interpolate(b3, b3.newVector(0.0), f)                       # different grids
    th = dfx.makeBasis(GRIDS["3d_3x4x5"], taylorHood())
    fv = dfx.makeDiscreteGlobalBasisFunction(dfx.subspaceBasis(th, 0), th.newVector(1.0))
    with pytest.raises(capi.RangeError):
        dfx.interpolate(dfx.makeBasis(GRIDS["3d_3x4x5"], power(lagrange(1), 2, BI())), torch.zeros(2 * 4 * 5 * 6, dtype=torch.float64, device="cuda"), fv)


# ---- evaluation at world coordinates ----------------------------------------------------------------
@pytest.mark.parametrize("grid,tree", list(all_cases()))
def test_evaluate_at_world_coordinates(grid, tree):
    """DiscreteGlobalBasisFunction::operator()(x) (discreteglobalbasisfunction.hh:402-410): random points,
    points on element faces / vertices / the domain boundary (first element of the search wins, which
    matters for DG), points outside (element -1, NaN)"""
    import torch
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    d = grid.dim
    lo = np.array(grid.lower if grid.lower is not None else [0.0] * d)
    up = np.array(grid.upper if grid.upper is not None else [1.0] * d)
    rng = np.random.default_rng(17)
    pts = [lo + (up - lo) * rng.uniform(0, 1, (40, d))]
    # lattice points of the grid (faces, edges, vertices incl. the boundary) and just beside them
    nodes = np.stack(np.meshgrid(*[lo[j] + (up[j] - lo[j]) * np.arange(grid.n[j] + 1) / grid.n[j] for j in range(d)],
                                 indexing="ij"), -1).reshape(-1, d)
    pts += [nodes, nodes + 1e-15, nodes - 1e-15]
    pts.append(np.array([up + 0.25 * (up - lo), lo - 1e-3 * (up - lo)]))        # outside
    x = np.concatenate(pts)
    c = random_coeff(b.dimension())
    f = dfx.makeDiscreteGlobalBasisFunction(b, torch.from_numpy(c).cuda())
    v, j, el = f.evaluateGlobal(x, jacobians=True, return_elements=True)
    vr, jr, er = o.evaluateGlobal(c, x, jacobians=True)
    assert np.array_equal(el.cpu().numpy().astype(np.int64), er)
    assert (er[-2:] == -1).all() and (er[:40] >= 0).all()
    inside = er >= 0
    assert np.isnan(v.cpu().numpy()[~inside]).all() and np.isnan(j.cpu().numpy()[~inside]).all()
    assert rel(v.cpu().numpy()[inside], vr[inside]) <= TOL
    assert rel(j.cpu().numpy()[inside], jr[inside]) <= 1e-11      # sums of k^2-sized terms for the higher orders


def test_interpolation_consistency_through_world_coordinates():
    """second variant of checkInterpolationConsistency (discreteglobalbasisfunctiontest.cc:58-76): a
    lambda forces interpolate to use the GLOBAL evaluation of the discrete function; here the node
    positions come from dfx_interpolation_points and the values from dfx_evaluate_global"""
    import torch
    for gname, tname in (("2d_10x10", "lagrange1"), ("2d_3x2", "power2_q1_FI"), ("3d_3x4x5", "lagrange2"),
                         ("3d_2x1x3", "taylorHood"), ("3d_3x4x5", "lagrange3")):
        grid = GRIDS[gname]
        tree = trees(grid.dim)[tname]
        b = dfx.makeBasis(grid, tree)
        x = torch.from_numpy(random_coeff(b.dimension())).cuda()
        f = dfx.makeDiscreteGlobalBasisFunction(b, x)
        xyz, leaf = dfx.interpolationPoints(b)
        vals = f.evaluateGlobal(xyz)                              # [n, n_leaves]
        nodal = vals.gather(1, leaf.long().unsqueeze(1)).squeeze(1).contiguous()
        y = b.newVector(float("nan"))
        dfx.interpolateFromValues(b, y, nodal)
        assert float((y - x).abs().max()) <= 1e-10, (gname, tname)


def test_evaluate_at_world_coordinates_on_a_slab():
    import torch
    whole = YaspGrid([3, 2, 6], upper=[1.0, 0.5, 2.0])
    mine = whole.slab(1, 3)
    b = dfx.makeBasis(mine, lagrange(2))
    o = orc.OracleBasis(mine, lagrange(2))
    c = random_coeff(b.dimension())
    x = np.array([[0.3, 0.2, 0.9], [0.3, 0.2, 0.1], [0.5, 0.25, 2.0 / 3], [0.5, 0.25, 4.0 / 3], [0.1, 0.1, 1.9]])
    f = dfx.makeDiscreteGlobalBasisFunction(b, torch.from_numpy(c).cuda())
    v, el = f.evaluateGlobal(x, return_elements=True)
    vr, _, er = o.evaluateGlobal(c, x)
    assert np.array_equal(el.cpu().numpy().astype(np.int64), er) and er[1] == -1 and er[4] == -1 and er[0] >= 0
    ok = er >= 0
    assert rel(v.cpu().numpy()[ok], vr[ok]) <= TOL


def test_python_front_door_local_view():
    """the reference's Python interface (dune/python/functions/globalbasis.hh:141-252): len(basis),
    localView().bind/index/len, basis.interpolate, basis.asFunction"""
    import torch
    grid, tree = GRIDS["3d_3x4x5"], trees(3)["TH_BL(BI)"]
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    assert len(b) == o.dimension()
    dig, lens = o.indices()
    lv = b.localView()
    assert len(lv) == lv.size() == lv.maxSize() == o.maxNodeSize() and not lv.bound()
    for e in (0, 7, 59):
        lv.bind(e)
        for i in (0, 26, 80, 88):
            assert lv.index(i) == [int(x) for x in dig[e, i, :lens[e, i]]]
    lv.bind((2, 3, 4))
    assert lv.element() == 2 + 3 * (3 + 4 * 4)
    lv.unbind()
    with pytest.raises(capi.DfxError):
        lv.index(0)
    with pytest.raises(capi.RangeError):
        lv.bind(60)
    q2 = dfx.makeBasis(grid, lagrange(2))
    x = q2.newVector(0.0)
    q2.interpolate(x, dfx.gaussian(1.0))
    u = q2.asFunction(x)
    v = u.evaluate(np.array([[0.5, 0.5, 0.5]]))
    assert v.shape == (60, 1, 1) and bool(torch.isfinite(v).all())


def test_peer_memory_halo_link_between_two_slabs():
    """dfx_halo_link_* / dfx_halo_push / dfx_halo_pull with both slabs on one GPU (raw-pointer
    attachment instead of CUDA IPC): the upper slab's bottom plane lands in the lower slab's top
    plane, step after step, through both receive buffers, without any host synchronisation."""
    import torch
    whole = YaspGrid([3, 2, 6], upper=[1.0, 0.5, 2.0])
    L = capi.lib()
    for tree in (lagrange(2), trees(3)["TH_FL(FI)"], lagrange(3)):
        b0, b1 = dfx.makeBasis(whole.slab(0, 2), tree), dfx.makeBasis(whole.slab(1, 2), tree)
        l0, l1 = C.c_void_p(), C.c_void_p()
        capi.check(L.dfx_halo_link_create(b0._h, C.byref(l0)))
        capi.check(L.dfx_halo_link_create(b1._h, C.byref(l1)))
        capi.check(L.dfx_halo_link_attach(l1, 0, None, L.dfx_halo_link_block(l0)))     # slab 1 pushes into slab 0
        capi.check(L.dfx_halo_link_attach(l0, 1, None, L.dfx_halo_link_block(l1)))     # slab 0 acknowledges into slab 1
        n = int(L.dfx_halo_size(b0._h))
        assert n == int(L.dfx_halo_size(b1._h)) and n > 0
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        c0 = torch.full((b0.dimension(),), float("nan"), dtype=torch.float64, device="cuda")
        top, bottom = torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
        for step in range(1, 6):
            c1 = torch.from_numpy(random_coeff(b1.dimension(), seed=step)).cuda()
            capi.check(L.dfx_halo_push(b1._h, l1, C.c_void_p(c1.data_ptr()), step, st))
            capi.check(L.dfx_halo_pull(b0._h, l0, C.c_void_p(c0.data_ptr()), step, st))
            capi.check(L.dfx_halo_pack(b0._h, 1, C.c_void_p(c0.data_ptr()), C.c_void_p(top.data_ptr()), st))
            capi.check(L.dfx_halo_pack(b1._h, 0, C.c_void_p(c1.data_ptr()), C.c_void_p(bottom.data_ptr()), st))
            assert torch.equal(top, bottom), step
        assert L.dfx_halo_link_error(l0) == 0 and L.dfx_halo_link_error(l1) == 0
        # everything but the top plane of slab 0 is untouched
        assert int(torch.isnan(c0).sum()) == b0.dimension() - n
        L.dfx_halo_link_destroy(l0)
        L.dfx_halo_link_destroy(l1)


@pytest.mark.parametrize("tree,n", [(lagrange(17), [2, 1, 1]), (lagrange(17), [2, 3]), (lagrange(11), [1, 2, 2]),
                                    (lagrangeDG(12), [2, 1, 2]), (lagrangeDG(12), [3])])
def test_maximum_orders_indices_and_nodes(tree, n):
    """the largest orders the reference admits (LagrangeFaceDOFPermutation::maxOrder3d = 17,
    lagrangebasis.hh:579, :795-796): index table bit-exact, interpolated coordinates bit-exact"""
    grid = YaspGrid(n)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    assert b.dimension() == o.dimension()
    assert np.array_equal(b.indices().cpu().numpy()[:, :, 0].astype(np.uint64), o.indices()[0][:, :, 0])
    for j in range(len(n)):
        x = b.newVector(0.0)
        dfx.interpolate(b, x, dfx.coord(j))
        assert np.array_equal(x.cpu().numpy(), o.interpolate(dfx.coord(j)))
    got = dfx.boundaryDOFs(b).cpu().numpy()
    assert np.array_equal(got, o.boundaryDOFs()[0])


def test_empty_and_ragged_element_ranges():
    """zero-length ranges are no-ops, sub-ranges that start and end inside a row give the same
    entries as the full table; out-of-range requests raise"""
    import torch
    grid, tree = GRIDS["3d_3x4x5"], trees(3)["taylorHood"]
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    ne = b.numElements()
    full = b.flatIndices().cpu().numpy()
    assert b.flatIndices(7, 7).shape[0] == 0 and b.indices(ne, ne).shape[0] == 0
    for e0, e1 in ((1, 2), (5, 23), (ne - 1, ne)):
        assert np.array_equal(b.flatIndices(e0, e1).cpu().numpy(), full[e0:e1])
        assert np.array_equal(b.indices(e0, e1).cpu().numpy().astype(np.uint64), o.indices(e0, e1)[0])
    c = torch.from_numpy(random_coeff(b.dimension())).cuda()
    f = dfx.makeDiscreteGlobalBasisFunction(b, c)
    pts, _ = tensor_gauss(3, 3)
    assert f.evaluate(pts, elem_begin=4, elem_end=4).shape[0] == 0
    v = f.evaluate(pts, elem_begin=5, elem_end=23).cpu().numpy()
    assert rel(v, o.evaluate(c.cpu().numpy(), pts, e0=5, e1=23)) <= TOL
    with pytest.raises(capi.RangeError):
        b.flatIndices(0, ne + 1)
    with pytest.raises(capi.RangeError):
        f.evaluate(pts, elem_begin=3, elem_end=2)


# ---- the reference's known answers, now through the GPU path ------------------------------------
@pytest.mark.parametrize("tree", [lagrange(0), lagrange(1), lagrange(2), lagrange(3), lagrange(4),
                                  lagrangeDG(1), lagrangeDG(2), lagrangeDG(3)])
@pytest.mark.parametrize("n", [[7], [10, 10], [4, 3, 5]])
def test_integral_of_interpolated_x0_is_half(tree, n):
    """gridviewfunctionspacebasistest.cc:115-183"""
    b = dfx.makeBasis(YaspGrid(n), tree)
    x = dfx.interpolate(b, b.newVector(0.0), dfx.coord(0))
    pts, w = tensor_gauss(len(n), 3)
    v = dfx.makeDiscreteGlobalBasisFunction(b, x).evaluate(pts).cpu().numpy()
    integral = np.einsum("eqc,q->", v, w) / b.numElements()
    assert abs(integral - 0.5) < 1e-10


def test_round_trip_basis_function_interpolation():
    """discreteglobalbasisfunctiontest.cc:49-77: interpolate(basis, y, makeDGBF(basis, x)) == x,
    done here by evaluating at the Lagrange nodes and scattering with the flat index table"""
    import torch
    grid = GRIDS["3d_2x1x3"]
    tree = power(lagrange(2), 2, BI())
    b = dfx.makeBasis(grid, tree)
    x = torch.from_numpy(random_coeff(b.dimension())).cuda()
    nodes = np.array([[(i % 3) / 2, ((i // 3) % 3) / 2, (i // 9) / 2] for i in range(27)])
    v = dfx.makeDiscreteGlobalBasisFunction(b, x).evaluate(nodes)          # [ne, 27, 2]
    flat = b.flatIndices().long()                                           # [ne, 54] = [comp0 | comp1]
    y = torch.zeros_like(x)
    y[flat[:, :27].reshape(-1)] = v[:, :, 0].reshape(-1)
    y[flat[:, 27:].reshape(-1)] = v[:, :, 1].reshape(-1)
    assert (y - x).abs().max().item() < 1e-10


def test_derivative_of_q3_interpolant():
    """discreteglobalbasisfunctionderivativetest.cc:80-158 on its own 10x42 grid"""
    grid = YaspGrid([10, 42])
    b = dfx.makeBasis(grid, lagrange(3))
    f = dfx.quadratic(0.0, [0, 0, 0], [42, 7, 0, 13, 0, 0])
    x = dfx.interpolate(b, b.newVector(0.0), f)
    pts, w = tensor_gauss(2, 3)
    v, j = dfx.makeDiscreteGlobalBasisFunction(b, x).evaluate(pts, jacobians=True)
    v, j = v.cpu().numpy(), j.cpu().numpy()
    e = np.arange(420)
    X = ((e % 10)[:, None] + pts[None, :, 0]) / 10
    Y = ((e // 10)[:, None] + pts[None, :, 1]) / 42
    l2 = np.sqrt(np.einsum("eq,q->", (j[:, :, 0, 0] - (84 * X + 7 * Y)) ** 2 + (j[:, :, 0, 1] - (26 * Y + 7 * X)) ** 2, w) / 420)
    assert l2 <= 6.6e-11
    assert np.abs(v[:, :, 0] - (42 * X * X + 13 * Y * Y + 7 * X * Y)).max() < 1e-10


# ---- every kernel family gives the same answer ---------------------------------------------------
@pytest.fixture
def kernel_path():
    def set_path(op, path):
        capi.lib().dfx_set_kernel_path(op, path)
    yield set_path
    for op in (capi.OP_EVALUATE, capi.OP_INTERPOLATE, capi.OP_INDICES, capi.OP_SMEM_GRADIENT):
        capi.lib().dfx_set_kernel_path(op, -1)


@pytest.mark.parametrize("path", [0, 1, 2])
@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_3x4x5"])
def test_indices_kernel_families(kernel_path, gname, path):
    """0 generic per-entry kernel, 1 walking kernel through a shared-memory tile (bulk stores), 2 walking kernel
    with direct stores"""
    import torch
    kernel_path(capi.OP_INDICES, path)
    grid = GRIDS[gname]
    for name, tree in trees(grid.dim).items():
        if "TH_" in name and grid.dim != 3:
            continue
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        assert (b.indices().cpu().numpy().astype(np.uint64) == o.indices()[0]).all(), name
        assert (b.flatIndices(dtype=torch.int64).cpu().numpy().astype(np.uint64) == o.flatIndices()).all(), name


@pytest.mark.parametrize("path", [0, 1, 2])
@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_3x4x5", "3d_2x1x3"])
def test_interpolate_kernel_families(kernel_path, gname, path):
    """generic per-DOF kernel, structured per-node kernel, separable-table kernel"""
    kernel_path(capi.OP_INTERPOLATE, path)
    grid = GRIDS[gname]
    for name, tree in trees(grid.dim).items():
        if "TH_" in name and grid.dim != 3:
            continue
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        for fname, f in FUNCS.items():
            x = b.newVector(-5.0)
            dfx.interpolate(b, x, f)
            ref = o.interpolate(f, coeff=np.full(o.dimension(), -5.0))
            assert rel(x.cpu().numpy(), ref) <= TOL, (name, fname)
            if fname in ("coord0", "const", "quadratic"):
                assert (x.cpu().numpy() == ref).all(), (name, fname)


def test_partitioned_interpolate_matches_global_run(kernel_path):
    """a z-slab handle reproduces the serial run on the whole grid bit for bit, including the
    interface plane whose last writer lives in the neighbouring slab (include/dfx.h, dfx_grid_desc)"""
    whole = YaspGrid([4, 3, 8], upper=[1.0, 0.7, 1.3])
    for tree in (lagrange(2), lagrange(3), lagrange(1), lagrangeDG(2)):
        og = orc.OracleBasis(whole, tree)
        ref = og.interpolate(dfx.coord(2))
        gflat = og.flatIndices().reshape(8, 12, -1)
        for path in (0, 1, 2):
            kernel_path(capi.OP_INTERPOLATE, path)
            for r in range(4):
                slab = whole.slab(r, 4)
                b = dfx.makeBasis(slab, tree)
                x = dfx.interpolate(b, b.newVector(0.0), dfx.coord(2)).cpu().numpy()
                lflat = b.flatIndices().cpu().numpy().reshape(2, 12, -1)
                assert (x[lflat] == ref[gflat[2 * r:2 * r + 2].astype(np.int64)]).all()


def test_halo_pack_unpack_and_owned_ranges():
    """one-plane halo between z-slabs: pack the lower plane of slab r+1, unpack into the top
    plane of slab r; owned ranges tile the global vector exactly once"""
    import torch
    whole = YaspGrid([3, 2, 6])
    L = capi.lib()
    for tree in (lagrange(2), power(lagrange(2), 3, BI()), composite(power(lagrange(2), 3, FL()), lagrange(1), FL()),
                 lagrange(3), lagrangeDG(1)):
        og = orc.OracleBasis(whole, tree)
        gx = random_coeff(og.dimension())
        gflat = og.flatIndices().reshape(6, 6, -1).astype(np.int64)
        slabs = [dfx.makeBasis(whole.slab(r, 3), tree) for r in range(3)]
        vecs = []
        for r, b in enumerate(slabs):
            lflat = b.flatIndices().cpu().numpy().reshape(2, 6, -1).astype(np.int64)
            x = np.full(b.dimension(), np.nan)
            x[lflat] = gx[gflat[2 * r:2 * r + 2]]
            vecs.append(torch.from_numpy(x).cuda())
        n = int(L.dfx_halo_size(slabs[0]._h))
        for r in range(2):
            top_before = vecs[r].clone()
            buf = torch.zeros(max(n, 1), dtype=torch.float64, device="cuda")
            capi.check(L.dfx_halo_pack(slabs[r + 1]._h, 0, C.c_void_p(vecs[r + 1].data_ptr()), C.c_void_p(buf.data_ptr()), None))
            vecs[r].mul_(1.0)
            # wipe the top plane, then restore it from the neighbour's buffer
            scratch = vecs[r].clone()
            capi.check(L.dfx_halo_unpack(slabs[r]._h, 1, C.c_void_p(torch.full_like(buf, -3.0).data_ptr()), C.c_void_p(scratch.data_ptr()), None))
            assert ((scratch == -3.0).sum().item() == n)
            capi.check(L.dfx_halo_unpack(slabs[r]._h, 1, C.c_void_p(buf.data_ptr()), C.c_void_p(scratch.data_ptr()), None))
            assert torch.equal(scratch, top_before)
        if all(True for _ in [0]) and tree.kind != capi.NODE_POWER:
            # owned ranges: gather the slabs into the global vector
            out = np.full(og.dimension(), np.nan)
            for r, b in enumerate(slabs):
                lb, cnt, gb = (C.c_uint64 * 64)(), (C.c_uint64 * 64)(), (C.c_uint64 * 64)()
                nr = C.c_int32()
                capi.check(L.dfx_owned_ranges(b._h, lb, cnt, gb, 64, C.byref(nr)))
                xl = vecs[r].cpu().numpy()
                for i in range(nr.value):
                    assert np.isnan(out[gb[i]:gb[i] + cnt[i]]).all()
                    out[gb[i]:gb[i] + cnt[i]] = xl[lb[i]:lb[i] + cnt[i]]
            assert (out == gx).all()


@pytest.mark.parametrize("k", [2, 3, 4, 5])
@pytest.mark.parametrize("last_fastest", [True, False])
def test_evaluate_shared_memory_tma_pair_loads(kernel_path, k, last_fastest):
    """element-local DOFs: the coefficients of element pairs arrive by TMA bulk loads (two buffers per warp,
    mbarrier completion); both point orders (the permuted one reads the natural layout with a stride), values and —
    forced — gradients, several pairs per warp, ranges that start on an even / odd element (the odd one is not
    16-byte aligned and takes the register path)"""
    import torch
    L = capi.lib()
    grid = YaspGrid([6, 5, 4], upper=[1.0, 2.0, 0.5])
    tree = lagrangeDG(k)
    b = dfx.makeBasis(grid, tree)
    o = orc.OracleBasis(grid, tree)
    c = random_coeff(b.dimension())
    f = dfx.makeDiscreteGlobalBasisFunction(b, torch.from_numpy(c).cuda())
    ne = b.numElements()
    kernel_path(capi.OP_EVALUATE, 2)
    try:
        for m in (k + 1, k + 2):
            pts, _ = tensor_gauss(3, m, last_fastest)
            vr, jr = o.evaluate(c, pts, jacobians=True)
            # k_eval_pair (both elements of a pair at once; 8 / 4 warps per block, block tiles), then k_eval_smem with its TMA modes
            # grids of 2 blocks: every warp walks through 8-15 items (buffer reuse, mbarrier phases, store waits)
            for pair, mode, cap in ((1, 1, 2), (2, 1, 2), (2, 1, 1), (4, 1, 2), (4, 1, 1), (-1, 1, -1), (0, 2, 2), (0, 1, 2), (0, 0, 2), (0, 1, -1)):
                L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_PAIR, pair)
                L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_TMA, mode)
                L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_BLOCKS, cap)
                v, j = f.evaluate(pts, jacobians=True)
                assert rel(v.cpu().numpy(), vr) <= TOL and rel(j.cpu().numpy(), jr) <= TOL, (m, pair, mode)
                assert rel(f.evaluate(pts).cpu().numpy(), vr) <= TOL, (m, pair, mode)
                j2 = f.evaluate(pts, values=False, jacobians=True)
                assert rel(j2.cpu().numpy(), jr) <= TOL, (m, pair, mode)
                for e0, e1 in ((2, ne), (1, ne - 1), (4, 6), (3, 4)):
                    assert rel(f.evaluate(pts, elem_begin=e0, elem_end=e1).cpu().numpy(), vr[e0:e1]) <= TOL, (m, pair, mode, e0)
                    js = f.evaluate(pts, values=False, jacobians=True, elem_begin=e0, elem_end=e1)
                    assert rel(js.cpu().numpy(), jr[e0:e1]) <= TOL, (m, pair, mode, e0)
    finally:
        L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_TMA, -1)
        L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_PAIR, -1)
        L.dfx_set_kernel_path(capi.OP_TUNE_SMEM_BLOCKS, -1)


@pytest.mark.parametrize("k", [2, 3, 4, 5])
@pytest.mark.parametrize("dm", [-1, 0, 1])
@pytest.mark.parametrize("last_fastest", [True, False])
def test_evaluate_shared_memory_path(kernel_path, k, dm, last_fastest):
    """path 2 (one element per warp, sum factorisation through shared memory): every
    instantiated (order, points-per-direction) pair, odd element counts, multi-leaf trees"""
    m = k + 1 + dm
    pts, _ = tensor_gauss(3, m, last_fastest)
    for grid in (YaspGrid([3, 3, 5], upper=[1.0, 2.0, 0.5]), GRIDS["3d_2x1x3"]):
        for tree in (lagrange(k), lagrangeDG(k), power(lagrange(k), 3, BI()),
                     composite(power(lagrangeDG(k), 2, FL()), lagrange(k), BL())):
            b = dfx.makeBasis(grid, tree)
            import ctypes
            kernel_path(capi.OP_EVALUATE, -1)
            path = capi.lib().dfx_evaluate_path(b._h, 0, pts.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(pts), 1)
            assert path == (1 if (k == 2 and m in (2, 3)) else 2)
            kernel_path(capi.OP_EVALUATE, 2)
            # gradients by collocation of the point values (automatic when m >= k + 1) and by
            # differentiated shape tables in every stage
            for grad_mode in (-1, 1):
                kernel_path(capi.OP_SMEM_GRADIENT, grad_mode)
                v, j, vr, jr = eval_both(grid, tree, pts, path=2)
                assert rel(v, vr) <= TOL and rel(j, jr) <= TOL, grad_mode
            kernel_path(capi.OP_SMEM_GRADIENT, -1)
            # values only + element sub-range with an odd start
            import torch
            o = orc.OracleBasis(grid, tree)
            c = random_coeff(b.dimension())
            kernel_path(capi.OP_EVALUATE, 2)
            f = dfx.makeDiscreteGlobalBasisFunction(b, torch.from_numpy(c).cuda())
            ne = b.numElements()
            got = f.evaluate(pts, elem_begin=1, elem_end=ne).cpu().numpy()
            assert rel(got, o.evaluate(c, pts, e0=1, e1=ne)) <= TOL


@pytest.mark.parametrize("gname", ["2d_10x10", "3d_3x4x5", "3d_2x1x3"])
def test_evaluate_dense_tensor_core_path(kernel_path, gname):
    """path 3 (FP64 mma.sync dense contraction): arbitrary points pick it automatically for
    elements with >= 16 local DOFs; forced on tensor points it must agree as well"""
    import ctypes
    grid = GRIDS[gname]
    rng = np.random.default_rng(11)
    pts = rng.uniform(0, 1, (7, grid.dim))
    tp, _ = tensor_gauss(grid.dim, 3)
    dim = grid.dim
    big = [lagrange(3), lagrangeDG(3), lagrange(4), power(lagrange(3), 2, BI()),
           composite(power(lagrangeDG(3), 2, FL()), lagrange(3), BL())]
    if dim == 3:
        big += [lagrange(2), lagrangeDG(4), taylorHood()]
    for tree in big:
        b = dfx.makeBasis(grid, tree)
        kernel_path(capi.OP_EVALUATE, -1)
        nodes = [0] if tree.kind != capi.NODE_TAYLOR_HOOD else [b.child(0, 0)]
        for node in nodes:
            expect = 3
            got = capi.lib().dfx_evaluate_path(b._h, node, pts.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(pts), 1)
            assert got == expect, (gname, got)
            o = orc.OracleBasis(grid, tree)
            c = random_coeff(b.dimension())
            import torch
            f = dfx.makeDiscreteGlobalBasisFunction(_Sub(b, node), torch.from_numpy(c).cuda())
            for p, path in ((pts, -1), (tp, 3)):
                kernel_path(capi.OP_EVALUATE, path)
                v, j = f.evaluate(p, jacobians=True)
                vr, jr = o.evaluate(c, p, node=node, jacobians=True)
                assert rel(v.cpu().numpy(), vr) <= TOL and rel(j.cpu().numpy(), jr) <= TOL
                v2 = f.evaluate(p, elem_begin=1, elem_end=b.numElements() - 1)
                assert rel(v2.cpu().numpy(), o.evaluate(c, p, node=node, e0=1, e1=b.numElements() - 1)) <= TOL
                j2 = f.evaluate(p, values=False, jacobians=True)
                assert rel(j2.cpu().numpy(), jr) <= TOL
