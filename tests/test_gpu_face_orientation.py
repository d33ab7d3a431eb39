"""GPU parity for SURVEY.md section 8(f) row 3: continuous Lagrange bases of order >= 3 on a grid whose
vertex ids are not monotone (dfx_basis_set_vertex_ids; Experimental::FaceOrientations and
LagrangeFaceDOFPermutation, lagrangebasis.hh:207-697).  The CUDA path finds a facet's orientation from
the vertex with the smallest id directly, the oracle through the reference's edge-bit tables; both
are compared bit for bit on indices and <= 1e-12 on coefficients / evaluations."""
import ctypes as C

import numpy as np
import pytest

import orc
import dune_functions_b200 as dfx
from dune_functions_b200 import _capi as capi
from dune_functions_b200 import (YaspGrid, lagrange, lagrangeDG, power, composite, tensor_gauss,
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


def ids_for(n, seed):
    nv = int(np.prod([m + 1 for m in n]))
    return (np.random.default_rng(seed).permutation(nv).astype(np.uint64) * 5 + 11)


def grid_with_ids(n, seed, **kw):
    return YaspGrid(list(n), vertex_ids=ids_for(n, seed), **kw)


BOXES = {"2d_3x2": (3, 2), "2d_1x1": (1, 1), "2d_5x4": (5, 4), "3d_2x2x2": (2, 2, 2), "3d_3x2x4": (3, 2, 4),
         "3d_1x1x1": (1, 1, 1)}


def trees(dim):
    return {
        "Q3": lagrange(3), "Q4": lagrange(4), "Q5": lagrange(5), "Q2": lagrange(2),
        "powFI_Q3": power(lagrange(3), dim, FI()), "powBL_Q4": power(lagrange(4), 2, BL()),
        "comp_Q4_Q2": composite(lagrange(4), lagrange(2), BL()),
        "comp_Q3_DG2_FL": composite(lagrange(3), lagrangeDG(2), FL()),
    }


def cases():
    for gname, n in BOXES.items():
        for tname, tree in trees(len(n)).items():
            yield pytest.param(n, tree, id=f"{gname}-{tname}")


@pytest.mark.parametrize("n,tree", list(cases()))
def test_indices_bit_exact_with_permuted_face_dofs(n, tree):
    import torch
    for seed in (0, 1):
        grid = grid_with_ids(n, seed)
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        dig, lens = o.indices()
        assert (b.indices().cpu().numpy().astype(np.uint64) == dig).all()
        assert (b.indices(dtype=torch.int64).cpu().numpy().astype(np.uint64) == dig).all()
        flat = o.flatIndices()
        assert (b.flatIndices().cpu().numpy().astype(np.uint64) == flat).all()
        assert (b.flatIndices(dtype=torch.int64).cpu().numpy().astype(np.uint64) == flat).all()
        out = np.zeros((b.numElements(), b.maxNodeSize()), dtype=np.uint32)
        capi.check(capi.lib().dfx_indices_flat_host(b._h, 0, b.numElements(), out.ctypes.data))
        assert (out.astype(np.uint64) == flat).all()


def test_ids_change_the_numbering_only_from_order_three_on():
    for n in ((3, 2), (2, 2, 2)):
        for k in (1, 2, 3, 4):
            plain = dfx.makeBasis(YaspGrid(list(n)), lagrange(k))
            base = plain.flatIndices().cpu().numpy()
            b = dfx.makeBasis(grid_with_ids(n, 3), lagrange(k))
            assert b.hasPermutedFaceDOFs() == (k >= 3)
            cur = b.flatIndices().cpu().numpy()
            assert (cur != base).any() == (k >= 3)
            # ids removed again / monotone ids given explicitly: the default numbering
            b.setVertexIds(None)
            assert not b.hasPermutedFaceDOFs()
            assert (b.flatIndices().cpu().numpy() == base).all()
            nv = int(np.prod([m + 1 for m in n]))
            b.setVertexIds(np.arange(nv, dtype=np.uint64) * 3)
            assert (b.flatIndices().cpu().numpy() == base).all()
    # 1-d: nothing to permute
    b = dfx.makeBasis(YaspGrid([5], vertex_ids=[5, 3, 9, 1, 0, 2]), lagrange(4))
    assert not b.hasPermutedFaceDOFs()
    assert (b.flatIndices().cpu().numpy() == dfx.makeBasis(YaspGrid([5]), lagrange(4)).flatIndices().cpu().numpy()).all()


def test_vertex_ids_from_device_memory_and_wrong_count():
    import torch
    n = (3, 2, 2)
    ids = ids_for(n, 5)
    b = dfx.makeBasis(YaspGrid(list(n)), lagrange(3))
    b.setVertexIds(torch.from_numpy(ids.astype(np.int64)).cuda())
    torch.cuda.synchronize()
    o = orc.OracleBasis(YaspGrid(list(n), vertex_ids=ids), lagrange(3))
    assert (b.flatIndices().cpu().numpy().astype(np.uint64) == o.flatIndices()).all()
    with pytest.raises(capi.RangeError):
        b.setVertexIds(ids[:-1])


FUNCS = {"gaussian": dfx.gaussian(1.0), "quadratic": dfx.quadratic(0.5, [1, -2, 3], [42, 7, 1, 13, -3, 5]),
         "coord1": dfx.coord(1)}


@pytest.mark.parametrize("n", [(3, 2), (5, 4), (2, 2, 2), (3, 2, 4)], ids=str)
@pytest.mark.parametrize("k", [3, 4, 5])
def test_interpolate_and_evaluate_with_permuted_face_dofs(n, k):
    import torch
    dim = len(n)
    grid = grid_with_ids(n, 7, upper=[1.0, 0.7, 1.3][:dim])
    b = dfx.makeBasis(grid, lagrange(k))
    o = orc.OracleBasis(grid, lagrange(k))
    rng = np.random.default_rng(2)
    for fname, f in FUNCS.items():
        x = dfx.interpolate(b, b.newVector(-7.0), f)
        ref = o.interpolate(f)
        if fname == "gaussian":
            assert rel(x.cpu().numpy(), ref) <= TOL
        else:
            assert (x.cpu().numpy() == ref).all(), fname            # polynomial f: bit-exact
        # masked
        mask = rng.random(b.dimension()) < 0.5
        xm = dfx.interpolate(b, b.newVector(-7.0), f, mask=torch.from_numpy(mask).cuda())
        refm = o.interpolate(f, coeff=np.full(b.dimension(), -7.0), mask=mask)
        assert rel(xm.cpu().numpy(), refm) <= TOL
    # form (2): node positions from the library, values from the caller
    xyz, leaf = dfx.interpolationPoints(b)
    vals = torch.exp(-(xyz ** 2).sum(dim=1))
    x2 = dfx.interpolateFromValues(b, b.newVector(0.0), vals)
    assert rel(x2.cpu().numpy(), o.interpolate(dfx.gaussian(1.0))) <= TOL
    # evaluation of a random vector: values and Jacobians at random and at tensor Gauss points
    c = rng.uniform(-1, 1, b.dimension())
    ct = torch.from_numpy(c).cuda()
    for pts in (rng.random((5, dim)), tensor_gauss(dim, 3)[0]):
        v, j = dfx.makeDiscreteGlobalBasisFunction(b, ct).evaluate(pts, jacobians=True)
        vr, jr = o.evaluate(c, pts, jacobians=True)
        assert rel(v.cpu().numpy(), vr) <= TOL
        assert rel(j.cpu().numpy(), jr) <= TOL
    # the permuted basis is conforming: a polynomial of degree <= k is reproduced at random points, and its
    # coefficients are those of the default numbering in another order
    up = [1.0, 0.7, 1.3][:dim]
    plain = dfx.makeBasis(YaspGrid(list(n), upper=up), lagrange(k))
    xq = dfx.interpolate(b, b.newVector(0.0), FUNCS["quadratic"])
    xp = dfx.interpolate(plain, plain.newVector(0.0), FUNCS["quadratic"])
    pts = rng.random((4, dim))
    vq = dfx.makeDiscreteGlobalBasisFunction(b, xq).evaluate(pts).cpu().numpy()
    vp = dfx.makeDiscreteGlobalBasisFunction(plain, xp).evaluate(pts).cpu().numpy()
    assert rel(vq, vp) <= TOL
    assert sorted(xq.cpu().numpy().tolist()) == sorted(xp.cpu().numpy().tolist())


def test_vector_valued_and_mixed_trees():
    import torch
    n = (3, 2, 2)
    grid = grid_with_ids(n, 9)
    rng = np.random.default_rng(4)
    for tree in (power(lagrange(3), 3, FI()), power(lagrange(4), 3, BL()), composite(lagrange(4), lagrange(2), BL()),
                 composite(power(lagrange(3), 3, BI()), lagrange(2), BL())):
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        assert b.hasPermutedFaceDOFs()
        nl = b.numLeaves()
        f = dfx.gaussian_vec(nl, 0.5) if nl > 1 else dfx.gaussian(0.5)
        x = dfx.interpolate(b, b.newVector(0.0), f)
        assert rel(x.cpu().numpy(), o.interpolate(f)) <= TOL
        c = rng.uniform(-1, 1, b.dimension())
        pts = tensor_gauss(3, 2)[0]
        v, j = dfx.makeDiscreteGlobalBasisFunction(b, torch.from_numpy(c).cuda()).evaluate(pts, jacobians=True)
        vr, jr = o.evaluate(c, pts, jacobians=True)
        assert rel(v.cpu().numpy(), vr) <= TOL and rel(j.cpu().numpy(), jr) <= TOL


def test_world_coordinates_grid_functions_boundary_and_pattern():
    import torch
    from dune_functions_b200 import operators as ops
    n = (3, 2, 3)
    grid = grid_with_ids(n, 13, upper=[1.0, 0.7, 1.3])
    rng = np.random.default_rng(6)
    b = dfx.makeBasis(grid, lagrange(4))
    o = orc.OracleBasis(grid, lagrange(4))
    c = rng.uniform(-1, 1, b.dimension())
    ct = torch.from_numpy(c).cuda()
    # DiscreteGlobalBasisFunction::operator()(x)
    xs = rng.random((40, 3)) * np.array([1.0, 0.7, 1.3])
    fn = dfx.makeDiscreteGlobalBasisFunction(b, ct)
    v, j = fn.evaluateGlobal(xs, jacobians=True)
    vr, jr, _ = o.evaluateGlobal(c, xs, jacobians=True)
    assert rel(v.cpu().numpy(), vr) <= TOL and rel(j.cpu().numpy(), jr) <= 1e-11
    # interpolate(target, coeff, gridFunction): permuted source -> DG target, plain source -> permuted target
    tdg, odg = dfx.makeBasis(grid, lagrangeDG(3)), orc.OracleBasis(grid, lagrangeDG(3))
    y = dfx.interpolate(tdg, tdg.newVector(0.0), fn)
    assert rel(y.cpu().numpy(), odg.interpolateGridFunction(o, c)) <= TOL
    s2, os2 = dfx.makeBasis(grid, lagrange(2)), orc.OracleBasis(grid, lagrange(2))
    c2 = rng.uniform(-1, 1, s2.dimension())
    z = dfx.interpolate(b, b.newVector(0.0), dfx.makeDiscreteGlobalBasisFunction(s2, torch.from_numpy(c2).cuda()))
    assert rel(z.cpu().numpy(), o.interpolateGridFunction(os2, c2)) <= TOL
    # boundary DOFs and the occupation pattern are sets of whole entities: compare with the oracle anyway
    m = dfx.boundaryDOFs(b).cpu().numpy()
    assert (m.astype(bool) == o.boundaryDOFs()[0].astype(bool)).all()
    rp, col = ops.getOccupationPattern(b)
    rpo, colo = o.occupationPattern()
    assert (rp.cpu().numpy().astype(np.uint64) == rpo).all() and (col.cpu().numpy().astype(np.uint32) == colo).all()
    # assembled / matrix-free Laplace are not available on a permuted basis: loud, not wrong
    with pytest.raises(capi.NotImplementedDfx):
        ops.assembleLaplaceMatrix(b, None, (rp, col))
    with pytest.raises(capi.NotImplementedDfx):
        ops.LaplaceOperator(b).mv(ct)


def test_slabs_with_their_share_of_the_ids_reproduce_the_whole_grid():
    whole = grid_with_ids((3, 2, 8), 17, upper=[1.0, 0.7, 1.3])
    tree = lagrange(4)
    og = orc.OracleBasis(whole, tree)
    ref = og.interpolate(FUNCS["quadratic"])
    gflat = og.flatIndices().reshape(8, 6, -1).astype(np.int64)
    for r in range(4):
        b = dfx.makeBasis(whole.slab(r, 4), tree)
        x = dfx.interpolate(b, b.newVector(0.0), FUNCS["quadratic"]).cpu().numpy()
        lflat = b.flatIndices().cpu().numpy().reshape(2, 6, -1)
        assert (x[lflat] == ref[gflat[2 * r:2 * r + 2]]).all()


def test_update_to_another_grid_matches_a_fresh_basis():
    """basis.update(gridView), defaultglobalbasis.hh:137-141 (lagrangebasistest.cc:163-175 updates its bases
    after adapting the grid and checks them again): same results as a basis made on the new grid; device
    tables of the old grid are dropped, vertex ids are taken from the new grid"""
    import torch
    pts = tensor_gauss(3, 3)[0]
    for tree in (lagrange(2), lagrange(4), composite(power(lagrange(2), 3, BI()), lagrange(1), BL()), lagrangeDG(3)):
        b = dfx.makeBasis(YaspGrid([5, 3, 2]), tree)
        x = dfx.interpolate(b, b.newVector(0.0), dfx.gaussian(1.0))              # fills the handle's device tables
        dfx.makeDiscreteGlobalBasisFunction(b, x).evaluate(pts, jacobians=True)
        for new in (YaspGrid([3, 4, 6], upper=[1.0, 0.7, 1.3]), grid_with_ids((2, 3, 3), 21)):
            b.update(new)
            f = dfx.makeBasis(new, tree)
            assert b.dimension() == f.dimension()
            assert (b.flatIndices().cpu().numpy() == f.flatIndices().cpu().numpy()).all()
            xb = dfx.interpolate(b, b.newVector(0.0), dfx.gaussian(1.0))
            xf = dfx.interpolate(f, f.newVector(0.0), dfx.gaussian(1.0))
            assert (xb.cpu().numpy() == xf.cpu().numpy()).all()
            vb, jb = dfx.makeDiscreteGlobalBasisFunction(b, xb).evaluate(pts, jacobians=True)
            vf, jf = dfx.makeDiscreteGlobalBasisFunction(f, xf).evaluate(pts, jacobians=True)
            assert (vb.cpu().numpy() == vf.cpu().numpy()).all() and (jb.cpu().numpy() == jf.cpu().numpy()).all()
            o = orc.OracleBasis(new, tree)
            assert (b.flatIndices().cpu().numpy().astype(np.uint64) == o.flatIndices()).all()
        torch.cuda.synchronize()


@pytest.mark.parametrize("n", [(3, 2), (3, 2, 4)], ids=str)
def test_evaluation_through_the_companion_basis_equals_the_general_kernel(n):
    """with permuted face DOFs dfx_evaluate gathers the element-local coefficients (k_gather_local) and
    evaluates them on the companion basis' fast kernels; kernel path 0 keeps the general kernel.  Both
    against the oracle, for sub-ranges of elements, subtrees and Jacobians."""
    import torch
    dim = len(n)
    grid = grid_with_ids(n, 23, upper=[1.0, 0.7, 1.3][:dim])
    rng = np.random.default_rng(8)
    trees_ = [lagrange(3), lagrange(5), power(lagrange(4), dim, FI()), composite(power(lagrange(3), dim, BI()), lagrange(2), BL()),
              composite(lagrange(4), lagrangeDG(2), FL())]
    try:
        for tree in trees_:
            b = dfx.makeBasis(grid, tree)
            o = orc.OracleBasis(grid, tree)
            c = rng.uniform(-1, 1, b.dimension())
            ct = torch.from_numpy(c).cuda()
            ne = b.numElements()
            for pts in (tensor_gauss(dim, 4)[0], tensor_gauss(dim, 3, last_fastest=False)[0], rng.random((6, dim))):
                vr, jr = o.evaluate(c, pts, jacobians=True)
                # (evaluate path, index path): companion basis with the fast / the general gather walk, general kernel
                for path, ipath in ((-1, -1), (-1, 0), (0, -1)):
                    capi.lib().dfx_set_kernel_path(capi.OP_EVALUATE, path)
                    capi.lib().dfx_set_kernel_path(capi.OP_INDICES, ipath)
                    f = dfx.makeDiscreteGlobalBasisFunction(b, ct)
                    v, j = f.evaluate(pts, jacobians=True)
                    assert rel(v.cpu().numpy(), vr) <= TOL and rel(j.cpu().numpy(), jr) <= TOL
                    v2 = f.evaluate(pts, elem_begin=1, elem_end=ne - 1)
                    assert rel(v2.cpu().numpy(), vr[1:ne - 1]) <= TOL
            if b.numLeaves() > 1:
                sub = dfx.subspaceBasis(b, 0)
                pts = tensor_gauss(dim, 3)[0]
                node = b.child(0, 0)
                vr = o.evaluate(c, pts, node=node)
                for path, ipath in ((-1, -1), (-1, 0), (0, -1)):
                    capi.lib().dfx_set_kernel_path(capi.OP_EVALUATE, path)
                    capi.lib().dfx_set_kernel_path(capi.OP_INDICES, ipath)
                    v = dfx.makeDiscreteGlobalBasisFunction(sub, ct).evaluate(pts)
                    assert rel(v.cpu().numpy(), vr) <= TOL
            # the host entry point (bands of z-layers) goes the same way
            capi.lib().dfx_set_kernel_path(capi.OP_EVALUATE, -1)
            capi.lib().dfx_set_kernel_path(capi.OP_INDICES, -1)
            pts = np.ascontiguousarray(tensor_gauss(dim, 3)[0])
            nc = b.numLeaves()
            out = np.zeros((ne, len(pts), nc))
            capi.check(capi.lib().dfx_evaluate_host(b._h, 0, c.ctypes.data, pts.ctypes.data_as(C.POINTER(C.c_double)), len(pts), 0, ne,
                                                    out.ctypes.data, None))
            assert rel(out, o.evaluate(c, pts)) <= TOL
    finally:
        capi.lib().dfx_set_kernel_path(capi.OP_EVALUATE, -1)
        capi.lib().dfx_set_kernel_path(capi.OP_INDICES, -1)


@pytest.mark.parametrize("n", [(5, 4), (3, 2, 4), (7, 3, 2)], ids=str)
def test_index_kernel_families_with_permuted_face_dofs(n):
    """the walking index kernel (one local index per thread, orientation per element) and the general one
    (one thread per entry) give the oracle's table, also for element sub-ranges and both index widths"""
    import torch
    dim = len(n)
    grid = grid_with_ids(n, 29)
    try:
        for tree in (lagrange(3), lagrange(6), power(lagrange(4), dim, BL()), composite(power(lagrange(3), dim, FI()), lagrange(1), BL())):
            b = dfx.makeBasis(grid, tree)
            o = orc.OracleBasis(grid, tree)
            dig, _ = o.indices()
            flat = o.flatIndices()
            ne = b.numElements()
            for path in (-1, 0):
                capi.lib().dfx_set_kernel_path(capi.OP_INDICES, path)
                for dt in (None, torch.int64):
                    assert (b.indices(dtype=dt).cpu().numpy().astype(np.uint64) == dig).all()
                    assert (b.flatIndices(dtype=dt).cpu().numpy().astype(np.uint64) == flat).all()
                    assert (b.flatIndices(2, ne - 1, dtype=dt).cpu().numpy().astype(np.uint64) == flat[2:ne - 1]).all()
    finally:
        capi.lib().dfx_set_kernel_path(capi.OP_INDICES, -1)


@pytest.mark.parametrize("n", [(70, 3, 2), (9, 5, 4), (1, 1, 1)], ids=str)
@pytest.mark.parametrize("last_fastest", [False, True])
def test_q3_register_kernel_with_permuted_face_dofs(n, last_fastest):
    """order 3 at 4^3 points, values only, with vertex ids: k_eval_q3<.., ORIENT> reads one orientation byte per
    edge / facet (the code table k_orient_codes builds once per set of ids) and permutes the in-entity index of
    its gathers.  Against the oracle — whole grid, sub-ranges, a component of an interleaved power tree — against
    the gather + companion route, and again after the ids have been replaced (the table is rebuilt) and removed."""
    import torch
    L = capi.lib()
    rng = np.random.default_rng(31)
    pts, _ = tensor_gauss(3, 4, last_fastest)
    for tree, sub in ((lagrange(3), None), (power(lagrange(3), 2, FI()), 1)):
        grid = grid_with_ids(n, 41, upper=[1.0, 0.7, 1.3])
        b = dfx.makeBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        c = rng.uniform(-1, 1, b.dimension())
        ct = torch.from_numpy(c).cuda()
        ne = b.numElements()
        space = b if sub is None else dfx.subspaceBasis(b, sub)
        node = 0 if sub is None else b.child(0, sub)
        f = dfx.makeDiscreteGlobalBasisFunction(space, ct)
        ref = o.evaluate(c, pts, node=node)
        try:
            for q3 in (1, 0):
                L.dfx_set_kernel_path(capi.OP_TUNE_EVAL_Q3, q3)
                assert rel(f.evaluate(pts).cpu().numpy(), ref) <= TOL, (n, q3)
                if ne > 2:
                    assert rel(f.evaluate(pts, elem_begin=1, elem_end=ne - 1).cpu().numpy(), ref[1:ne - 1]) <= TOL, (n, q3)
        finally:
            L.dfx_set_kernel_path(capi.OP_TUNE_EVAL_Q3, -1)
        # other ids: the numbering changes, the code table with it
        ids2 = ids_for(n, 77)
        b.setVertexIds(ids2)
        o2 = orc.OracleBasis(YaspGrid(list(n), vertex_ids=ids2, upper=[1.0, 0.7, 1.3]), tree)
        assert rel(f.evaluate(pts).cpu().numpy(), o2.evaluate(c, pts, node=node)) <= TOL
        # no ids: the plain kernel
        b.setVertexIds(None)
        o3 = orc.OracleBasis(YaspGrid(list(n), upper=[1.0, 0.7, 1.3]), tree)
        assert rel(f.evaluate(pts).cpu().numpy(), o3.evaluate(c, pts, node=node)) <= TOL
