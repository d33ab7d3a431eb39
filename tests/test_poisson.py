"""The element-loop callers of examples/poisson-pq2.cc: occupation
pattern, Laplace stiffness matrix, load vector, symmetric Dirichlet treatment, matrix-vector
product.  CPU part: the oracle restatement against known answers.  GPU part: the device path
(dfx_matrix_pattern / dfx_assemble_laplace / dfx_csr_mv / dfx_apply_dirichlet) against the oracle
— pattern bit-exact, values within 1e-12 — and a Poisson problem solved end to end."""
import numpy as np
import pytest
import scipy.sparse as sp

import orc
from cases import GRIDS, trees
import dune_functions_b200 as dfx
from dune_functions_b200 import YaspGrid, lagrange, lagrangeDG, constant, coord, gaussian, sinprod, quadratic

TOL = 1e-12
SCALAR = ["lagrange1", "lagrange2", "lagrange3", "lagrangeDG1", "lagrangeDG2"]


def rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def scipy_of(csr):
    row_ptr, col, val = csr
    n = len(row_ptr) - 1
    return sp.csr_matrix((val, col, row_ptr.astype(np.int64)), shape=(n, n))


# ---- oracle pinned by known answers -----------------------------------------------------------
def test_oracle_q1_1d_is_the_second_difference_matrix():
    n = 5
    ob = orc.OracleBasis(YaspGrid([n]), lagrange(1))
    csr, rhs = ob.assembleLaplace(constant(1.0))
    A = scipy_of(csr).toarray()
    h = 1.0 / n
    exp = (np.diag([1.0] + [2.0] * (n - 1) + [1.0]) - np.diag([1.0] * n, 1) - np.diag([1.0] * n, -1)) / h
    assert np.abs(A - exp).max() < 1e-13
    assert np.allclose(rhs, [h / 2] + [h] * (n - 1) + [h / 2], atol=1e-15)


@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_2x1x3"])
@pytest.mark.parametrize("tname", ["lagrange1", "lagrange2", "lagrange3"])
def test_oracle_stiffness_matrix_properties(gname, tname):
    grid = GRIDS[gname]
    ob = orc.OracleBasis(grid, trees(grid.dim)[tname])
    csr, rhs = ob.assembleLaplace(constant(1.0))
    A = scipy_of(csr)
    assert abs(A - A.T).max() < 1e-12                     # symmetric
    assert np.abs(A @ np.ones(A.shape[0])).max() < 1e-11    # constants are in the kernel
    lo = grid.lower if grid.lower is not None else [0.0] * grid.dim
    up = grid.upper if grid.upper is not None else [1.0] * grid.dim
    vol = np.prod([u - l for u, l in zip(up, lo)])
    assert abs(rhs.sum() - vol) < 1e-13                     # sum_i int phi_i = |Omega|
    # u = x_0 is reproduced: energy = |Omega|, and (A u)_i = 0 away from the boundary
    u = ob.interpolate(coord(0))
    assert abs(u @ (A @ u) - vol) < 1e-11
    bnd, _ = ob.boundaryDOFs()
    inner = (A @ u)[~bnd.astype(bool)]
    assert inner.size == 0 or np.abs(inner).max() < 1e-11


def test_oracle_pattern_row_lengths():
    ob = orc.OracleBasis(YaspGrid([4, 4]), lagrange(1))
    row_ptr, col = ob.occupationPattern()
    lens = np.diff(row_ptr.astype(np.int64)).reshape(5, 5)
    assert lens[2, 2] == 9 and lens[0, 0] == 4 and lens[0, 2] == 6
    ob = orc.OracleBasis(YaspGrid([3, 3, 3]), lagrange(2))
    row_ptr, col = ob.occupationPattern()
    lens = np.diff(row_ptr.astype(np.int64))
    assert sorted(set(lens.tolist())) == [27, 45, 75, 125]  # element / face / edge / vertex DOFs (boundary ones coincide)
    for r in range(len(lens)):
        seg = col[row_ptr[r]:row_ptr[r + 1]]
        assert (np.diff(seg.astype(np.int64)) > 0).all() and r in seg


def test_oracle_dirichlet_treatment():
    grid = YaspGrid([4, 3])
    ob = orc.OracleBasis(grid, lagrange(2))
    csr, rhs = ob.assembleLaplace(constant(10.0))
    bnd, _ = ob.boundaryDOFs()
    x = ob.interpolate(sinprod(2.0), coeff=np.zeros(ob.dimension()), mask=bnd)
    A0 = scipy_of(csr).toarray()
    rhs0 = rhs.copy()
    orc.oracle_dirichlet(csr, bnd, x, rhs)
    A = scipy_of(csr).toarray()
    b = bnd.astype(bool)
    assert np.array_equal(A[b][:, b], np.eye(b.sum())) and not A[b][:, ~b].any() and not A[~b][:, b].any()
    assert np.array_equal(A[~b][:, ~b], A0[~b][:, ~b])
    assert np.array_equal(rhs[b], x[b])
    assert np.allclose(rhs[~b], (rhs0 - A0 @ x)[~b], atol=1e-13)


def test_oracle_driven_cavity_of_the_stokes_example():
    """examples/stokes-taylorhood.cc main() (:299-480) step by step on the oracle — BL(BI) Taylor-Hood basis on
    YaspGrid 4x4, velocity boundary DOFs, Dirichlet data (0, [x0 < 1e-8]), identity rows, GMRES(500) to 1e-10
    from x0 = rhs.  These are the numbers tests/cpp/stokes_taylorhood.cc (the same main() on the façade,
    GPU suite) checks: 64 boundary DOFs, u(0.5, 0.5) = (0, -0.1567), discrete divergence ~ 0."""
    import scipy.sparse.linalg as spl
    from dune_functions_b200 import power, composite, blockedInterleaved as BI, blockedLexicographic as BL
    grid = YaspGrid([4, 4])
    ob = orc.OracleBasis(grid, composite(power(lagrange(2), 2, BI()), lagrange(1), BL()))
    n = ob.dimension()
    assert n == 2 * 81 + 25
    A = scipy_of(orc.oracle_stokes(ob)).toarray()
    assert np.abs(A - A.T).max() == 0.0 and not A[162:, 162:].any()
    bnd = ob.boundaryDOFs(node=1)[0].astype(bool)                 # forEachBoundaryDOF(subspaceBasis(basis, _0), ...)
    assert bnd.sum() == 64 and not bnd[162:].any()
    o2 = orc.OracleBasis(grid, lagrange(2))
    X, Y = o2.interpolate(coord(0)), o2.interpolate(coord(1))
    rhs = np.zeros(n)
    for i in range(81):                                           # (0, i, c) -> 2 i + c
        if bnd[2 * i + 1]:
            rhs[2 * i + 1] = float(X[i] < 1e-8)
    M = A.copy()
    M[bnd, :] = 0.0
    M[bnd, bnd] = 1.0
    its = [0]
    x, info = spl.gmres(sp.csr_matrix(M), rhs, x0=rhs.copy(), rtol=1e-10, restart=500, maxiter=2,
                        callback=lambda r: its.__setitem__(0, its[0] + 1), callback_type="pr_norm")
    assert info == 0 and its[0] < 187
    assert np.abs((A @ x)[162:]).max() < 1e-9                     # B u = 0
    mid = int(np.argmin((X - 0.5) ** 2 + (Y - 0.5) ** 2))
    left = int(np.argmin(X ** 2 + (Y - 0.5) ** 2))
    assert abs(x[2 * mid]) < 1e-9 and abs(x[2 * mid + 1] + 0.156722) < 1e-5
    assert x[2 * left] == 0.0 and x[2 * left + 1] == 1.0


@pytest.mark.parametrize("n", [[3, 2], [2, 2, 2]])
def test_oracle_stokes_matrix_properties(n):
    """assembleStokesMatrix restated (examples/stokes-taylorhood.cc): symmetric saddle-point matrix,
    velocity blocks = the Q2 Laplace matrix per component, no coupling between different velocity
    components or inside the pressure block, and B^T applied to a constant pressure vanishes on
    velocity rows away from the boundary (int div(phi) = 0 for interior shape functions)."""
    grid = YaspGrid(n)
    d = len(n)
    ob = orc.OracleBasis(grid, dfx.taylorHood())
    A = scipy_of(orc.oracle_stokes(ob)).toarray()
    assert np.abs(A - A.T).max() == 0.0
    o2 = orc.OracleBasis(grid, lagrange(2))
    csr2, _ = o2.assembleLaplace()
    K = scipy_of(csr2).toarray()
    n2 = o2.dimension()
    for c in range(d):                                        # TaylorHoodBasis = BL(FI): (0, i*d + c)
        assert np.abs(A[c:d * n2:d, c:d * n2:d] - K).max() <= 1e-13
        for c2 in range(d):
            if c2 != c:
                assert not A[c:d * n2:d, c2:d * n2:d].any()
    assert not A[d * n2:, d * n2:].any()
    bnd, _ = o2.boundaryDOFs()
    ones = np.ones(A.shape[0] - d * n2)
    bt1 = A[:d * n2, d * n2:] @ ones
    for c in range(d):
        assert np.abs(bt1[c::d][~bnd.astype(bool)]).max() <= 1e-13


# ---- GPU parity ----------------------------------------------------------------------------------
def all_cases():
    for gname, grid in GRIDS.items():
        for tname in trees(grid.dim):
            yield pytest.param(gname, tname, id=f"{gname}-{tname}")


@pytest.mark.gpu
@pytest.mark.parametrize("gname,tname", list(all_cases()))
def test_occupation_pattern_bit_exact(gname, tname):
    grid = GRIDS[gname]
    tree = trees(grid.dim)[tname]
    b = dfx.makeBasis(grid, tree)
    ob = orc.OracleBasis(grid, tree)
    row_ptr, col = dfx.getOccupationPattern(b)
    erp, ecol = ob.occupationPattern()
    assert np.array_equal(row_ptr.cpu().numpy().astype(np.uint64), erp)
    assert np.array_equal(col.cpu().numpy().view(np.uint32), ecol)


@pytest.mark.gpu
@pytest.mark.parametrize("gname", list(GRIDS))
@pytest.mark.parametrize("tname", SCALAR + ["lagrange4", "lagrange0"])
def test_laplace_matrix_and_load_vector(gname, tname):
    grid = GRIDS[gname]
    if tname == "lagrange4" and grid.dim == 3:
        tname = "lagrangeDG4"                             # (2k+1)^3 = 729 fits, keep the test small instead
    tree = trees(grid.dim)[tname]
    b = dfx.makeBasis(grid, tree)
    ob = orc.OracleBasis(grid, tree)
    for f in (gaussian(1.0), quadratic(1.0, [2.0, -1.0, 0.5], [1.0, 0.5, 0.25, -1.0, 2.0, 3.0])):
        A, rhs = dfx.assembleLaplaceMatrix(b, f)
        (erp, ecol, eval_), erhs = ob.assembleLaplace(f)
        assert np.array_equal(A.col_idx.cpu().numpy().view(np.uint32), ecol)
        assert rel(A.values.cpu().numpy(), eval_) <= TOL
        assert rel(rhs.cpu().numpy(), erhs) <= TOL


@pytest.mark.gpu
def test_laplace_of_a_slab_uses_the_global_geometry():
    whole = YaspGrid([3, 2, 6], upper=[1.0, 0.5, 2.0])
    for rank in range(3):
        mine = whole.slab(rank, 3)
        b = dfx.makeBasis(mine, lagrange(2))
        ob = orc.OracleBasis(mine, lagrange(2))
        A, rhs = dfx.assembleLaplaceMatrix(b, gaussian(1.0))
        (_, _, ev), er = ob.assembleLaplace(gaussian(1.0))
        assert rel(A.values.cpu().numpy(), ev) <= TOL and rel(rhs.cpu().numpy(), er) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("tname", ["lagrange1", "lagrange2", "lagrange3"])
def test_dirichlet_and_matrix_vector_product(tname):
    import torch
    grid = GRIDS["3d_3x4x5"]
    tree = trees(3)[tname]
    b = dfx.makeBasis(grid, tree)
    ob = orc.OracleBasis(grid, tree)
    f = constant(10.0)
    A, rhs = dfx.assembleLaplaceMatrix(b, f)
    csr, erhs = ob.assembleLaplace(f)
    xr = np.random.default_rng(5).uniform(-1, 1, b.dimension())
    y = A.mv(torch.from_numpy(xr).cuda()).cpu().numpy()
    assert rel(y, orc.oracle_mv(csr, xr)) <= TOL
    # Dirichlet values sin(2 x_0) sin(2 x_1) sin(2 x_2) on the boundary, poisson-pq2.cc:340-386
    dn = dfx.boundaryDOFs(b)
    x = b.newVector(0.0)
    dfx.interpolate(b, x, sinprod(2.0), mask=dn)
    dfx.applyDirichlet(A, dn, x, rhs)
    ebnd, _ = ob.boundaryDOFs()
    ex = ob.interpolate(sinprod(2.0), coeff=np.zeros(ob.dimension()), mask=ebnd)
    orc.oracle_dirichlet(csr, ebnd, ex, erhs)
    assert rel(A.values.cpu().numpy(), csr[2]) <= TOL
    assert rel(rhs.cpu().numpy(), erhs) <= TOL


@pytest.mark.gpu
def test_unsupported_operator_inputs_raise():
    from dune_functions_b200 import _capi as capi
    b = dfx.makeBasis(YaspGrid([2, 2]), dfx.taylorHood())
    with pytest.raises(capi.NotImplementedDfx):
        dfx.assembleLaplaceMatrix(b, constant(1.0))
    b = dfx.makeBasis(YaspGrid([2, 2, 2]), lagrange(6))
    with pytest.raises(capi.NotImplementedDfx):
        dfx.getOccupationPattern(b)                      # (2*6+1)^3 = 2197 columns per vertex row


@pytest.mark.gpu
@pytest.mark.parametrize("gname", ["2d_3x2", "2d_10x10", "3d_3x4x5", "3d_2x1x3", "3d_1x1x1"])
@pytest.mark.parametrize("k", [1, 2])
def test_matrix_free_laplace_equals_the_assembled_matrix(gname, k):
    """dfx_apply_laplace (never assembled, 3^d Gauss points) against the CSR product with the matrix
    of dfx_assemble_laplace and against the oracle's matrix, with and without Dirichlet rows"""
    import torch
    grid = GRIDS[gname]
    b = dfx.makeBasis(grid, lagrange(k))
    ob = orc.OracleBasis(grid, lagrange(k))
    A, _ = dfx.assembleLaplaceMatrix(b)
    csr, _ = ob.assembleLaplace()
    xr = np.random.default_rng(11).uniform(-1, 1, b.dimension())
    x = torch.from_numpy(xr).cuda()
    y = dfx.LaplaceOperator(b).mv(x).cpu().numpy()
    assert rel(y, orc.oracle_mv(csr, xr)) <= TOL
    assert rel(y, A.mv(x).cpu().numpy()) <= TOL
    # Dirichlet rows/columns: same as the product with the modified matrix
    dn = dfx.boundaryDOFs(b)
    rhs = torch.zeros_like(x)
    dfx.applyDirichlet(A, dn, torch.zeros_like(x), rhs)
    ym = dfx.LaplaceOperator(b, dn).mv(x).cpu().numpy()
    assert rel(ym, A.mv(x).cpu().numpy()) <= TOL


@pytest.mark.gpu
def test_matrix_free_laplace_slab_and_unsupported():
    import torch
    from dune_functions_b200 import _capi as capi
    whole = YaspGrid([3, 2, 6], upper=[1.0, 0.5, 2.0])
    mine = whole.slab(1, 3)
    b = dfx.makeBasis(mine, lagrange(2))
    ob = orc.OracleBasis(mine, lagrange(2))
    csr, _ = ob.assembleLaplace()
    xr = np.random.default_rng(3).uniform(-1, 1, b.dimension())
    y = dfx.LaplaceOperator(b).mv(torch.from_numpy(xr).cuda()).cpu().numpy()
    assert rel(y, orc.oracle_mv(csr, xr)) <= TOL
    for tree in (lagrange(3), lagrangeDG(1), dfx.taylorHood()):
        bb = dfx.makeBasis(YaspGrid([2, 2, 2]), tree)
        with pytest.raises(capi.NotImplementedDfx):
            dfx.LaplaceOperator(bb).mv(bb.newVector(1.0))


@pytest.mark.gpu
@pytest.mark.parametrize("gname", ["2d_3x2", "2d_10x10", "3d_3x4x5", "3d_2x1x3"])
@pytest.mark.parametrize("tname", ["taylorHood", "TH_FL(FL)", "TH_BL(BI)", "TH_FL(FI)", "TH_BL(BL)"])
def test_stokes_matrix(gname, tname):
    from dune_functions_b200 import _capi as capi
    grid = GRIDS[gname]
    tree = trees(grid.dim)[tname]
    b = dfx.makeBasis(grid, tree)
    ob = orc.OracleBasis(grid, tree)
    A = dfx.assembleStokesMatrix(b)
    erp, ecol, eval_ = orc.oracle_stokes(ob)
    assert np.array_equal(A.row_ptr.cpu().numpy().astype(np.uint64), erp)
    assert np.array_equal(A.col_idx.cpu().numpy().view(np.uint32), ecol)
    assert rel(A.values.cpu().numpy(), eval_) <= TOL
    if gname == "2d_3x2" and tname == "taylorHood":
        for bad in (lagrange(2), trees(2)["power2_q2_BL"], trees(2)["nested_FL_FI_BL"]):
            with pytest.raises(capi.NotImplementedDfx):
                dfx.assembleStokesMatrix(dfx.makeBasis(grid, bad))


def conjugate_gradient(A, rhs, x, tol=1e-12, maxit=2000):
    """plain CG on the device: A.mv is the library's SpMV, the vector updates are torch"""
    r = rhs - A.mv(x)
    p = r.clone()
    rr = torch_dot(r, r)
    r0 = rr
    for it in range(maxit):
        Ap = A.mv(p)
        alpha = rr / torch_dot(p, Ap)
        x += alpha * p
        r -= alpha * Ap
        rn = torch_dot(r, r)
        if rn <= tol * tol * r0:
            return it + 1
        p = r + (rn / rr) * p
        rr = rn
    return maxit


def torch_dot(a, b):
    return float((a * b).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("k,n,tol", [(1, 16, 1e-2), (2, 8, 2e-4), (3, 6, 2e-4)])
def test_poisson_problem_end_to_end(k, n, tol):
    """-Laplace u = 3 pi^2 sin(pi x) sin(pi y) sin(pi z), u = 0 on the boundary: the work flow of
    examples/poisson-pq2.cc (assemble, mark boundary DOFs, Dirichlet treatment, CG, evaluate)"""
    grid = YaspGrid([n, n, n])
    b = dfx.makeBasis(grid, lagrange(k))
    A, rhs = dfx.assembleLaplaceMatrix(b, sinprod(np.pi))
    rhs *= 3 * np.pi ** 2
    dn = dfx.boundaryDOFs(b)
    x = b.newVector(0.0)
    dfx.interpolate(b, x, constant(0.0), mask=dn)
    dfx.applyDirichlet(A, dn, x, rhs)
    x_mf = x.clone()
    its = conjugate_gradient(A, rhs, x)
    assert its < 2000
    if k <= 2:
        # the same solve without a matrix: CG on the matrix-free operator
        its_mf = conjugate_gradient(dfx.LaplaceOperator(b, dn), rhs, x_mf)
        assert its_mf < 2000 and float((x_mf - x).abs().max()) < 1e-9
    exact = b.newVector(0.0)
    dfx.interpolate(b, exact, sinprod(np.pi))
    assert float((x - exact).abs().max()) < tol
    # and through DiscreteGlobalBasisFunction at the element centres
    v = dfx.makeDiscreteGlobalBasisFunction(b, x).evaluate(np.array([[0.5, 0.5, 0.5]])).cpu().numpy().reshape(n, n, n)
    c = (np.arange(n) + 0.5) / n
    ex = np.sin(np.pi * c)[:, None, None] * np.sin(np.pi * c)[None, :, None] * np.sin(np.pi * c)[None, None, :]
    assert np.abs(v - ex).max() < tol
