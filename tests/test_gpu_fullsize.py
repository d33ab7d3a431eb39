"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the
oracle would need minutes to hours at these sizes): index-table invariants of basistest.hh,
the reference's known answers (integral of interpolated x_0 = 0.5, partition of unity,
constant interpolation), interpolate-evaluate round trips, exact gradients of polynomials."""
import numpy as np
import pytest

import dune_functions_b200 as dfx
from dune_functions_b200 import tensor_gauss

pytestmark = pytest.mark.gpu


def _free():
    import torch
    torch.cuda.empty_cache()


def check_index_table(basis, chunk=1 << 20):
    """every local index unique per element, indices cover 0..dimension-1 exactly (checkBasisIndices)"""
    import torch
    n = basis.dimension()
    hit = torch.zeros(n, dtype=torch.bool, device="cuda")
    ne = basis.numElements()
    for e0 in range(0, ne, chunk):
        idx = basis.flatIndices(e0, min(ne, e0 + chunk)).long()
        assert int(idx.min()) >= 0 and int(idx.max()) < n
        srt, _ = idx.sort(dim=1)
        assert bool((srt[:, 1:] != srt[:, :-1]).all()), "a local index appears twice in an element"
        hit[idx.reshape(-1)] = True
    assert bool(hit.all()), "some global index is never produced"


def test_c2_q2_256_interpolate_evaluate():
    """configs[1]: Q2 on 256^3, interpolate + evaluation at 3^3 Gauss points"""
    import torch
    basis = dfx.makeBasis(dfx.YaspGrid([256, 256, 256]), dfx.lagrange(2))
    assert basis.dimension() == 513 ** 3
    pts, w = tensor_gauss(3, 3)
    wt = torch.from_numpy(w).cuda()
    x = basis.newVector(0.0)
    # partition of unity / constant interpolation (makebasistest.cc:105-119)
    dfx.interpolate(basis, x, dfx.constant(1.0))
    assert bool((x == 1.0).all())
    v = dfx.makeDiscreteGlobalBasisFunction(basis, x).evaluate(pts)
    assert float((v - 1).abs().max()) < 1e-13
    # integral of the interpolant of x_0 (gridviewfunctionspacebasistest.cc:115-183)
    dfx.interpolate(basis, x, dfx.coord(0))
    dfx.makeDiscreteGlobalBasisFunction(basis, x).evaluate(pts, out_values=v)
    integral = float((v[:, :, 0] * wt).sum() / basis.numElements())
    assert abs(integral - 0.5) < 1e-10
    # the Gaussian: interpolant is within the Q2 interpolation error of f at the Gauss points
    dfx.interpolate(basis, x, dfx.gaussian(1.0))
    dfx.makeDiscreteGlobalBasisFunction(basis, x).evaluate(pts, out_values=v)
    e = torch.arange(0, basis.numElements(), 4099, device="cuda")
    ec = torch.stack([e % 256, (e // 256) % 256, e // 65536], 1).double()
    P = (ec[:, None, :] + torch.from_numpy(pts).cuda()[None]) / 256
    assert float((v[e, :, 0] - torch.exp(-(P ** 2).sum(-1))).abs().max()) < 1e-7
    del v
    _free()
    # round trip (discreteglobalbasisfunctiontest.cc:49-77): values at the Lagrange nodes are the coefficients
    nodes = np.array([[(i % 3) / 2, ((i // 3) % 3) / 2, (i // 9) / 2] for i in range(27)])
    ne = basis.numElements()
    for e0 in (0, ne // 2, ne - (1 << 20)):
        vn = dfx.makeDiscreteGlobalBasisFunction(basis, x).evaluate(nodes, elem_begin=e0, elem_end=e0 + (1 << 20))
        flat = basis.flatIndices(e0, e0 + (1 << 20)).long()
        assert float((vn[:, :, 0] - x[flat]).abs().max()) < 1e-13
    del vn, flat
    _free()
    check_index_table(basis)


def test_c3_taylor_hood_128_index_map_and_interpolation():
    """configs[2]: Taylor-Hood on 128^3: global multi-index map + interpolation"""
    import torch
    basis = dfx.makeBasis(dfx.YaspGrid([128, 128, 128]), dfx.taylorHood())
    n2, n1 = 257 ** 3, 129 ** 3
    assert basis.dimension() == 3 * n2 + n1 and basis.maxNodeSize() == 89
    assert basis.size(()) == 2 and basis.size((0,)) == 3 * n2 and basis.size((1,)) == n1 and basis.size((0, 5)) == 0
    check_index_table(basis, chunk=1 << 18)
    mi = basis.indices(0, 1 << 18)                       # [ne, 89, 2]
    assert bool((mi[:, :81, 0] == 0).all()) and bool((mi[:, 81:, 0] == 1).all())
    # velocity component c of Q2 DOF i -> (0, 3 i + c): the three components interleave
    assert bool((mi[:, 27:54, 1] == mi[:, :27, 1] + 1).all()) and bool((mi[:, 54:81, 1] == mi[:, :27, 1] + 2).all())
    del mi
    x = basis.newVector(float("nan"))
    dfx.interpolate(dfx.subspaceBasis(basis, 0), x, dfx.identity(3))
    dfx.interpolate(dfx.subspaceBasis(basis, 1), x, dfx.gaussian(1.0))
    assert not bool(torch.isnan(x).any())
    centre = np.array([[0.5, 0.5, 0.5]])
    v = dfx.makeDiscreteGlobalBasisFunction(dfx.subspaceBasis(basis, 0), x).evaluate(centre)   # [ne, 1, 3]
    e = torch.arange(basis.numElements(), device="cuda")
    exact = torch.stack([(e % 128 + 0.5) / 128, ((e // 128) % 128 + 0.5) / 128, (e // 16384 + 0.5) / 128], 1)
    assert float((v[:, 0, :] - exact).abs().max()) < 1e-13
    p = dfx.makeDiscreteGlobalBasisFunction(dfx.subspaceBasis(basis, 1), x).evaluate(centre)
    # Q1 interpolation error at the cell centre: <= 3 h^2/8 max|f''| = 3/(8*128^2)*2 = 4.6e-5
    assert float((p[:, 0, 0] - torch.exp(-(exact ** 2).sum(1))).abs().max()) < 6e-5


def test_c4_lagrange_dg4_192_values_and_gradients():
    """configs[3]: LagrangeDG order 4 on 192^3, values + gradients at 5^3 Gauss points.  The
    interpolant of a quadratic is exact, so values and gradients are known in closed form."""
    import torch
    basis = dfx.makeBasis(dfx.YaspGrid([192, 192, 192]), dfx.lagrangeDG(4))
    assert basis.dimension() == 192 ** 3 * 125
    x = basis.newVector(0.0)
    f = dfx.quadratic(0.0, [0, 0, 0], [42, 7, 0, 13, 0, 5])   # 42 x0^2 + 13 x1^2 + 7 x0 x1 + 5 x2^2
    dfx.interpolate(basis, x, f)
    pts, w = tensor_gauss(3, 5)
    ne = basis.numElements()
    tp = torch.from_numpy(pts).cuda()
    chunk = 1 << 19
    u = dfx.makeDiscreteGlobalBasisFunction(basis, x)
    for e0 in (0, ne // 3, ne - chunk):
        v, j = u.evaluate(pts, jacobians=True, elem_begin=e0, elem_end=e0 + chunk)
        e = torch.arange(e0, e0 + chunk, device="cuda")
        ec = torch.stack([e % 192, (e // 192) % 192, e // (192 * 192)], 1).double()
        P = (ec[:, None, :] + tp[None]) / 192
        X, Y, Z = P[..., 0], P[..., 1], P[..., 2]
        assert float((v[:, :, 0] - (42 * X * X + 13 * Y * Y + 7 * X * Y + 5 * Z * Z)).abs().max()) < 1e-11
        assert float((j[:, :, 0, 0] - (84 * X + 7 * Y)).abs().max()) < 1e-9
        assert float((j[:, :, 0, 1] - (26 * Y + 7 * X)).abs().max()) < 1e-9
        assert float((j[:, :, 0, 2] - (10 * Z)).abs().max()) < 1e-9
        del v, j
    idx = basis.flatIndices(ne - 1000, ne).long()
    assert bool((idx == (torch.arange(ne - 1000, ne, device="cuda")[:, None] * 125 + torch.arange(125, device="cuda")[None])).all())


def test_c5_q2_512_single_gpu():
    """configs[4] on one GPU: Q2 on 512^3 (1.08e9 DOFs, 3.6e9 point values)"""
    import torch
    basis = dfx.makeBasis(dfx.YaspGrid([512, 512, 512]), dfx.lagrange(2))
    assert basis.dimension() == 1025 ** 3
    x = basis.newVector(0.0)
    dfx.interpolate(basis, x, dfx.coord(2))
    pts, w = tensor_gauss(3, 3)
    v = dfx.makeDiscreteGlobalBasisFunction(basis, x).evaluate(pts)
    wt = torch.from_numpy(w).cuda()
    integral = float((v[:, :, 0] * wt).sum() / basis.numElements())
    assert abs(integral - 0.5) < 1e-10
    # x_2 is reproduced exactly up to rounding at every Gauss point of the last element layer
    last = v[-512 * 512:, :, 0]
    z = (511 + torch.from_numpy(pts[:, 2]).cuda()) / 512
    assert float((last - z[None, :]).abs().max()) < 1e-14
    del v
    _free()


def test_c2_q2_256_boundary_grid_function_and_matrix_free_operator():
    """the rows next to the path at the size of configs[1]: boundary-DOF
    count = lattice surface; interpolating the discrete function back returns its coefficients;
    the matrix-free Laplace operator annihilates constants, gives zero at interior nodes for a
    linear function and the energy |Omega| = 1 for u = x_0."""
    import torch
    basis = dfx.makeBasis(dfx.YaspGrid([256, 256, 256]), dfx.lagrange(2))
    n = basis.dimension()
    mask = dfx.boundaryDOFs(basis)
    assert int(mask.sum()) == 513 ** 3 - 511 ** 3
    x = basis.newVector(0.0)
    dfx.interpolate(basis, x, dfx.gaussian(1.0))
    y = basis.newVector(float("nan"))
    dfx.interpolate(basis, y, dfx.makeDiscreteGlobalBasisFunction(basis, x))
    assert float((y - x).abs().max()) < 1e-13
    del y
    _free()
    op = dfx.LaplaceOperator(basis)
    ones = basis.newVector(1.0)
    assert float(op.mv(ones).abs().max()) < 1e-12
    dfx.interpolate(basis, x, dfx.coord(0))
    ax = op.mv(x)
    assert abs(float((ax * x).sum()) - 1.0) < 1e-9
    assert float(ax[~mask.bool()].abs().max()) < 1e-12
    # with the Dirichlet rows: identity on the boundary, untouched interior rows
    opd = dfx.LaplaceOperator(basis, mask)
    z = torch.rand(n, dtype=torch.float64, device="cuda")
    az = opd.mv(z)
    assert bool((az[mask.bool()] == z[mask.bool()]).all())
    del op, opd, ax, az, z, ones
    _free()


def test_c2_world_coordinate_evaluation_of_many_points():
    """10^7 random world points on the Q2 256^3 function: the interpolant of a quadratic is exact"""
    import torch
    basis = dfx.makeBasis(dfx.YaspGrid([256, 256, 256]), dfx.lagrange(2))
    x = basis.newVector(0.0)
    dfx.interpolate(basis, x, dfx.quadratic(0.5, [1.0, -2.0, 3.0], [1.0, 0.5, -1.0, 2.0, 0.25, -0.5]))
    g = torch.Generator(device="cuda").manual_seed(5)
    p = torch.rand((10_000_000, 3), dtype=torch.float64, device="cuda", generator=g)
    f = dfx.makeDiscreteGlobalBasisFunction(basis, x)
    v, el = f.evaluateGlobal(p, return_elements=True)
    assert int(el.min()) >= 0 and int(el.max()) < basis.numElements()
    exact = (0.5 + p[:, 0] - 2 * p[:, 1] + 3 * p[:, 2] + p[:, 0] ** 2 + 0.5 * p[:, 0] * p[:, 1] - p[:, 0] * p[:, 2]
             + 2 * p[:, 1] ** 2 + 0.25 * p[:, 1] * p[:, 2] - 0.5 * p[:, 2] ** 2)
    assert float((v[:, 0] - exact).abs().max()) < 1e-12
    del v, el, p
    _free()
