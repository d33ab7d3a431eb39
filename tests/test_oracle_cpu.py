"""CPU suite, part 1: pin the oracle (oracle/dfx_oracle.cpp) against every golden vector
and known-answer test the reference holds for the path."""
import itertools

import numpy as np
import pytest

import orc
from cases import GRIDS, trees, th_variant
from dune_functions_b200 import (YaspGrid, lagrange, lagrangeDG, power, composite, taylorHood, tensor_gauss,
                                 flatLexicographic as FL, flatInterleaved as FI,
                                 blockedLexicographic as BL, blockedInterleaved as BI,
                                 gaussian, identity, constant, coord, quadratic)


def mi_list(basis):
    """all (element, local) multi-indices as tuples"""
    dig, lens = basis.indices()
    out = []
    for e in range(dig.shape[0]):
        out.append([tuple(int(x) for x in dig[e, i, :lens[e, i]]) for i in range(dig.shape[1])])
    return out


# ---- Appendix A of SURVEY.md: hand-derived numbering examples ------------------
def test_A1_q2_2d_2x1():
    b = orc.OracleBasis(YaspGrid([2, 1]), lagrange(2))
    assert b.dimension() == 15
    d, _ = b.indices()
    assert d[0, :, 0].tolist() == [9, 2, 10, 6, 0, 7, 12, 4, 13]
    assert d[1, :, 0].tolist() == [10, 3, 11, 7, 1, 8, 13, 5, 14]


def test_A2_q2_and_taylorhood_flfl_1x1():
    """what examples/interpolation.cc:42-79 builds"""
    b = orc.OracleBasis(YaspGrid([1, 1]), lagrange(2))
    assert b.dimension() == 9
    assert b.indices()[0][0, :, 0].tolist() == [5, 1, 6, 3, 0, 4, 7, 2, 8]
    th = orc.OracleBasis(YaspGrid([1, 1]), th_variant(FL, FL, 2))
    assert th.dimension() == 22
    assert th.indices()[0][0, :, 0].tolist() == [5, 1, 6, 3, 0, 4, 7, 2, 8, 14, 10, 15, 12, 9, 13, 16, 11, 17,
                                                  18, 19, 20, 21]


def test_A3_q2_3d_1x1x1():
    b = orc.OracleBasis(YaspGrid([1, 1, 1]), lagrange(2))
    expect = {0: 19, 1: 7, 2: 20, 3: 11, 4: 1, 5: 12, 6: 21, 7: 8, 8: 22, 9: 15, 10: 3, 11: 16, 12: 5, 13: 0, 14: 6,
              15: 17, 16: 4, 17: 18, 18: 23, 19: 9, 20: 24, 21: 13, 22: 2, 23: 14, 24: 25, 25: 10, 26: 26}
    got = b.indices()[0][0, :, 0].tolist()
    assert got == [expect[i] for i in range(27)]


@pytest.mark.parametrize("n", [[4], [3, 2], [3, 4, 5]])
def test_A4_q1_is_vertex_index(n):
    """P1 global index == indexSet.index(vertex): taylorhoodbasistest.cc:63-64, 121-122"""
    b = orc.OracleBasis(YaspGrid(n), lagrange(1))
    d, _ = b.indices()
    dim = len(n)
    ne = int(np.prod(n))
    for e in range(ne):
        ec = np.unravel_index(e, n[::-1])[::-1]
        for loc in range(2 ** dim):
            v, stride = 0, 1
            for j in range(dim):
                v += (ec[j] + ((loc >> j) & 1)) * stride
                stride *= n[j] + 1
            assert d[e, loc, 0] == v


@pytest.mark.parametrize("k", [0, 1, 2, 4])
def test_A5_lagrange_dg(k):
    """lagrangedgbasis.hh:120-130: elementIndex*(k+1)^d + i"""
    n = [3, 2, 2]
    b = orc.OracleBasis(YaspGrid(n), lagrangeDG(k))
    d, _ = b.indices()
    nl = (k + 1) ** 3
    assert b.dimension() == 12 * nl
    assert (d[:, :, 0] == (np.arange(12)[:, None] * nl + np.arange(nl)[None, :])).all()


# ---- the manual's golden table of the eight Taylor-Hood variants ------------------
TABLE = {
    # (outer, inner): (velocity (c, i) -> multi-index, pressure j -> multi-index) ; n2 = dim(Q2)
    ("BL", "BL"): (lambda c, i, n2: (0, c, i), lambda j, n2: (1, j)),
    ("BL", "BI"): (lambda c, i, n2: (0, i, c), lambda j, n2: (1, j)),
    ("BL", "FL"): (lambda c, i, n2: (0, c * n2 + i), lambda j, n2: (1, j)),
    ("BL", "FI"): (lambda c, i, n2: (0, 3 * i + c), lambda j, n2: (1, j)),
    ("FL", "BL"): (lambda c, i, n2: (c, i), lambda j, n2: (3 + j,)),
    ("FL", "BI"): (lambda c, i, n2: (i, c), lambda j, n2: (n2 + j,)),
    ("FL", "FL"): (lambda c, i, n2: (c * n2 + i,), lambda j, n2: (3 * n2 + j,)),
    ("FL", "FI"): (lambda c, i, n2: (3 * i + c,), lambda j, n2: (3 * n2 + j,)),
}
TAGS = {"BL": BL, "BI": BI, "FL": FL, "FI": FI}


@pytest.mark.parametrize("outer,inner", list(TABLE))
def test_manual_table_taylor_hood_variants(outer, inner):
    """doc/manual/dune-functions-bases.tex:1101-1292 (Table tab:th_indexing_variants)"""
    grid = YaspGrid([2, 3, 2])
    q2 = orc.OracleBasis(grid, lagrange(2))
    q1 = orc.OracleBasis(grid, lagrange(1))
    th = orc.OracleBasis(grid, th_variant(TAGS[outer], TAGS[inner], 3))
    n2 = q2.dimension()
    i2 = q2.indices()[0][:, :, 0]
    i1 = q1.indices()[0][:, :, 0]
    got = mi_list(th)
    vel, pres = TABLE[(outer, inner)]
    for e in range(th.numElements()):
        # local ordering: [v0 (27) | v1 (27) | v2 (27) | p (8)], nodes.hh:236-245
        for c in range(3):
            for i in range(27):
                assert got[e][c * 27 + i] == vel(c, int(i2[e, i]), n2)
        for j in range(8):
            assert got[e][81 + j] == pres(int(i1[e, j]), n2)
    assert th.dimension() == 3 * n2 + q1.dimension()


def test_taylorhood_prebasis_is_BL_FI():
    """TaylorHoodPreBasis::indicesImp<false> (taylorhoodbasis.hh:229-251) == column BL(FI)"""
    for n in ([2, 3], [2, 2, 3]):
        grid = YaspGrid(n)
        a = mi_list(orc.OracleBasis(grid, taylorHood()))
        b = mi_list(orc.OracleBasis(grid, th_variant(BL, FI, len(n))))
        assert a == b


# ---- basistest.hh invariants -------------------------------------------------------
def check_index_tree(basis):
    """checkBasisIndices (basistest.hh:190-227): multi-indices unique per element, they form a
    consecutive index tree starting at (0,...,0) (:103-143), size(prefix) is consistent
    (:150-180) and dimension() == number of distinct indices (:223)."""
    mis = mi_list(basis)
    allmi = set()
    for el in mis:
        assert len(set(el)) == len(el)          # each local index exactly once (checkLocalView :275-383)
        assert len(el) <= basis.maxNodeSize()
        allmi.update(el)
    assert len(allmi) == basis.dimension()
    lens = {len(m) for m in allmi}
    assert min(lens) == basis.minMultiIndexSize() and max(lens) == basis.maxMultiIndexSize()
    children = {}
    for m in allmi:
        for p in range(len(m)):
            children.setdefault(m[:p], set()).add(m[p])
    for prefix, ch in children.items():
        assert ch == set(range(len(ch))), f"non-consecutive digits below {prefix}"
        assert basis.size(prefix) == len(ch)
    for m in allmi:
        assert basis.size(m) == 0
        assert m not in children                 # no index is a strict prefix of another


@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_3x4x5"])
def test_index_tree_invariants(gname):
    grid = GRIDS[gname]
    for name, tree in trees(grid.dim).items():
        if "TH_" in name and grid.dim != 3:
            continue
        check_index_tree(orc.OracleBasis(grid, tree))


# ---- known answers: interpolation and evaluation -----------------------------------
def integrate(basis, coeff, order_pts=3, node=0):
    pts, w = tensor_gauss(basis.grid.dim, order_pts)
    v = basis.evaluate(coeff, pts, node=node)
    vol = 1.0 / basis.numElements()
    return np.einsum("eqc,q->c", v, w) * vol


@pytest.mark.parametrize("tree", [lagrange(0), lagrange(1), lagrange(2), lagrange(3), lagrange(4),
                                  lagrangeDG(1), lagrangeDG(2), lagrangeDG(3)])
@pytest.mark.parametrize("n", [[7], [4, 3], [3, 2, 2]])
def test_integral_of_interpolated_x0_is_half(tree, n):
    """gridviewfunctionspacebasistest.cc:115-183: int interpolate(x_0) = 0.5 (1e-10).
    (k = 0 interpolates the cell-centre value, exact for a linear function as well)"""
    b = orc.OracleBasis(YaspGrid(n), tree)
    c = b.interpolate(coord(0))
    assert abs(integrate(b, c)[0] - 0.5) < 1e-10


def test_taylorhood_pressure_integral_is_half():
    """taylorhoodbasistest.cc:58-64: pressure DOFs set to x_0 of the vertex with index
    gridView.indexSet().index(vertex) integrate to 0.5"""
    n = [10, 10]
    grid = YaspGrid(n)
    b = orc.OracleBasis(grid, taylorHood())
    flat = b.flatIndices()
    x = np.zeros(b.dimension())
    pressure_offset = int(flat[:, 18:].min())
    for vy in range(n[1] + 1):
        for vx in range(n[0] + 1):
            x[pressure_offset + vx + (n[0] + 1) * vy] = vx / n[0]
    p_node = 5   # composite(0) -> power(1) -> q2 (2,3) -> q1 (4)
    assert b.nodeInfo(4) == dict(local_offset=18, size=4, first_leaf=2, n_leaves=1)
    assert abs(integrate(b, x, node=4)[0] - 0.5) < 1e-10
    del p_node


@pytest.mark.parametrize("name", ["lagrange2", "power3_q2_BI", "taylorHood", "nested_BL_BI_FL", "nested_FL_FI_BL",
                                  "nested_pow_of_comp_BL"])
def test_constant_interpolation_gives_ones(name):
    """makebasistest.cc:105-119: interpolating the constant 1 makes every coefficient 1"""
    grid = GRIDS["2d_3x2"]
    b = orc.OracleBasis(grid, trees(2)[name])
    c = b.interpolate(constant(1.0), coeff=np.full(b.dimension(), -7.0))
    assert np.abs(c - 1).max() < 1e-10


@pytest.mark.parametrize("name", ["lagrange1", "lagrange2", "lagrange3", "lagrangeDG2", "power3_q2_BI", "taylorHood"])
def test_round_trip_interpolate_evaluate(name):
    """discreteglobalbasisfunctiontest.cc:49-77 / interpolatetest.hh:45-67: evaluating the
    discrete function at the basis' own interpolation nodes returns the coefficients"""
    grid = GRIDS["3d_2x1x3"]
    b = orc.OracleBasis(grid, trees(3)[name])
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, b.dimension())
    flat = b.flatIndices()
    nodes = b._nodes = None
    for leaf_node in range(b.numTreeNodes()):
        info = b.nodeInfo(leaf_node)
        if info["n_leaves"] != 1 or info["size"] == 0:
            continue
        # leaf order k from its size
        k = round(info["size"] ** (1 / 3)) - 1
        if k == 0:
            pts = np.full((1, 3), 0.5)
        else:
            pts = np.array([[(i % (k + 1)) / k, ((i // (k + 1)) % (k + 1)) / k, (i // (k + 1) ** 2) / k]
                            for i in range((k + 1) ** 3)])
        v = b.evaluate(x, pts, node=leaf_node)[:, :, 0]
        lo = info["local_offset"]
        expect = x[flat[:, lo:lo + info["size"]].astype(np.int64)]
        assert np.abs(v - expect).max() < 1e-10
    del nodes


def test_derivative_matches_analytic_gradient():
    """discreteglobalbasisfunctionderivativetest.cc:80-158: Q3 interpolant of a quadratic has the
    exact gradient (10x42 YaspGrid there; 5x7 here keeps the CPU suite short)"""
    grid = YaspGrid([5, 7])
    b = orc.OracleBasis(grid, lagrange(3))
    f = quadratic(0.0, [0, 0, 0], [42, 7, 0, 13, 0, 0])   # 42 x0^2 + 13 x1^2 + 7 x0 x1
    c = b.interpolate(f)
    pts, _ = tensor_gauss(2, 3)
    v, j = b.evaluate(c, pts, jacobians=True)
    for e in range(b.numElements()):
        ex, ey = e % 5, e // 5
        X = (ex + pts[:, 0]) / 5
        Y = (ey + pts[:, 1]) / 7
        assert np.abs(v[e, :, 0] - (42 * X * X + 13 * Y * Y + 7 * X * Y)).max() < 1e-10
        assert np.abs(j[e, :, 0, 0] - (84 * X + 7 * Y)).max() < 1e-9
        assert np.abs(j[e, :, 0, 1] - (26 * Y + 7 * X)).max() < 1e-9


def test_partition_of_unity_and_continuity():
    """gridviewfunctionspacebasistest.cc:36-110 (shape functions sum to 1) and the
    continuity check of basistest.hh:591-649 restated for values: a continuous Lagrange
    function evaluated on the common face from both sides agrees"""
    grid = YaspGrid([3, 2, 2])
    for k in (1, 2, 3):
        b = orc.OracleBasis(grid, lagrange(k))
        ones = np.ones(b.dimension())
        pts, _ = tensor_gauss(3, 3)
        assert np.abs(b.evaluate(ones, pts) - 1).max() < 1e-12
        x = np.random.default_rng(k).uniform(-1, 1, b.dimension())
        t = np.linspace(0.1, 0.9, 4)
        right = np.array([[1.0, a, c] for a in t for c in t])
        left = np.array([[0.0, a, c] for a in t for c in t])
        vr = b.evaluate(x, right)[:, :, 0].reshape(2, 2, 3, -1)
        vl = b.evaluate(x, left)[:, :, 0].reshape(2, 2, 3, -1)
        assert np.abs(vr[:, :, :-1] - vl[:, :, 1:]).max() < 1e-10


def test_masked_interpolation_and_subspace():
    """interpolate(basis, x, f, bitVector) (interpolate.hh:278-282, :169) on a subspace basis
    (examples/interpolation.cc:89-96): only marked entries of the selected subtree change"""
    grid = YaspGrid([3, 3])
    tree = th_variant(FL, FL, 2)
    b = orc.OracleBasis(grid, tree)
    n = b.dimension()
    x = np.full(n, 9.0)
    # velocity subtree = node 1 (power), vector-valued f(x) = x
    b.interpolate(identity(2), coeff=x, node=1)
    flat = b.flatIndices()
    vel = np.unique(flat[:, :18])
    pres = np.unique(flat[:, 18:])
    assert (x[pres] == 9.0).all() and (x[vel] != 9.0).any()
    y = np.full(n, 9.0)
    mask = np.zeros(n, dtype=np.uint8)
    mask[vel[::2]] = 1
    b.interpolate(identity(2), coeff=y, mask=mask, node=1)
    assert (y[mask == 0] == 9.0).all()
    assert (y[mask == 1] == x[mask == 1]).all()


def test_last_writer_wins_matches_serial_order():
    """interpolate.hh:254-259: elements are visited in grid-view order and every element
    writes all its local DOFs, so the value of a shared DOF is the one computed by the
    last element containing it; the coordinate it sees is lower + xhat*(upper-lower) of
    THAT element."""
    grid = YaspGrid([3], upper=[0.3])
    b = orc.OracleBasis(grid, lagrange(1))
    c = b.interpolate(coord(0))
    h = 0.3 / 3
    expect = [0.0 + 0.0 * (h - 0.0), h + 0.0 * (2 * h - h), 2 * h + 0.0 * (3 * h - 2 * h), 2 * h + 1.0 * (3 * h - 2 * h)]
    assert c.tolist() == expect


# ---- grid-function interpolation and world-coordinate evaluation of the oracle ---------------------
@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_2x1x3"])
def test_oracle_interpolation_consistency_of_discrete_functions(gname):
    """checkInterpolationConsistency (gridfunctions/test/discreteglobalbasisfunctiontest.cc:49-77):
    interpolate(basis, y, makeDiscreteGlobalBasisFunction(basis, x)) == x within 1e-10, for every tree"""
    grid = GRIDS[gname]
    rng = np.random.default_rng(2)
    for name, tree in trees(grid.dim).items():
        o = orc.OracleBasis(grid, tree)
        x = rng.uniform(-1, 1, o.dimension())
        y = o.interpolateGridFunction(o, x, coeff=np.full(o.dimension(), np.nan))
        assert np.abs(y - x).max() <= 1e-10, name


def test_oracle_grid_function_interpolation_between_bases():
    """a Q2 interpolant of a quadratic interpolated into Q1 / Q3 / DG2 bases equals the direct
    interpolation of the quadratic (all of them reproduce it at their nodes)"""
    grid = GRIDS["3d_2x1x3"]
    f = quadratic(0.5, [1.0, -2.0, 3.0], [1.0, 0.5, -1.0, 2.0, 0.25, -0.5])
    q2 = orc.OracleBasis(grid, lagrange(2))
    c2 = q2.interpolate(f)
    for tree in (lagrange(1), lagrange(3), lagrangeDG(2), lagrange(0)):
        o = orc.OracleBasis(grid, tree)
        assert np.abs(o.interpolateGridFunction(q2, c2) - o.interpolate(f)).max() <= 1e-12


@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_2x1x3"])
def test_oracle_world_coordinate_evaluation_agrees_with_local_evaluation(gname):
    """DiscreteGlobalBasisFunction::operator()(x) (discreteglobalbasisfunction.hh:402-410) = local
    evaluation in the element that contains x at geometry.local(x): take interior local points,
    map them to world coordinates and compare"""
    grid = GRIDS[gname]
    d = grid.dim
    lo = np.array(grid.lower if grid.lower is not None else [0.0] * d)
    up = np.array(grid.upper if grid.upper is not None else [1.0] * d)
    h = (up - lo) / np.array(grid.n)
    rng = np.random.default_rng(4)
    xhat = rng.uniform(0.1, 0.9, (3, d))
    for name in ("lagrange2", "lagrangeDG1", "taylorHood", "nested_FL_FI_BL"):
        o = orc.OracleBasis(grid, trees(d)[name])
        c = rng.uniform(-1, 1, o.dimension())
        v, j = o.evaluate(c, xhat, jacobians=True)
        ne = o.numElements()
        ec = np.stack(np.unravel_index(np.arange(ne), grid.n[::-1])[::-1], -1)          # x fastest
        world = (lo + (ec[:, None, :] + xhat[None]) * h).reshape(-1, d)
        vg, jg, el = o.evaluateGlobal(c, world, jacobians=True)
        assert np.array_equal(el, np.repeat(np.arange(ne), 3))
        assert np.abs(vg.reshape(v.shape) - v).max() <= 1e-12
        assert np.abs(jg.reshape(j.shape) - j).max() <= 1e-10
    # outside the grid: no element
    _, _, el = o.evaluateGlobal(c, np.array([up + 1.0]))
    assert el[0] == -1


# ---- the oracle's own bookkeeping: closed-form container position vs brute-force enumeration ----------
@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "2d_10x10", "3d_3x4x5", "3d_2x1x3"])
def test_flat_position_by_tree_rank_equals_enumeration(gname):
    """Basis::flat ranks a multi-index through size(prefix) (the way a nested container resolves
    vector[multiIndex], istlvectorbackend.hh:267-277); the brute-force route collects and sorts every
    multi-index of the grid.  Both must agree for every tree, and the positions must be a bijection
    onto 0..dimension()-1."""
    grid = GRIDS[gname]
    for name, tree in trees(grid.dim).items():
        if "TH_" in name and grid.dim != 3:
            continue
        b = orc.OracleBasis(grid, tree)
        fast, slow = b.flatIndices(), b.flatIndicesByEnumeration()
        assert np.array_equal(fast, slow), name
        assert set(fast.reshape(-1).tolist()) == set(range(b.dimension())), name


def test_interpolate_on_an_element_range_matches_the_full_loop():
    """orc_interpolate_range writes the DOFs of the elements it visits with the full loop's values
    (up to the rounding of the node coordinate on DOFs shared with elements outside the range)"""
    grid = YaspGrid([5, 4, 3])
    for tree in (lagrange(2), taylorHood(), lagrangeDG(2)):
        b = orc.OracleBasis(grid, tree)
        f = gaussian(1.0) if b.numLeaves() == 1 else None
        if f is None:
            full = np.zeros(b.dimension())
            b.interpolate(identity(3), coeff=full, node=1)
            b.interpolate(gaussian(1.0), coeff=full, node=5)
        else:
            full = b.interpolate(f)
        part = np.full(b.dimension(), np.nan)
        e0, e1 = 17, 41
        if f is None:
            b.interpolate(identity(3), coeff=part, node=1, e0=e0, e1=e1)
            b.interpolate(gaussian(1.0), coeff=part, node=5, e0=e0, e1=e1)
        else:
            b.interpolate(f, coeff=part, e0=e0, e1=e1)
        touched = np.unique(b.flatIndices(e0, e1).reshape(-1)).astype(np.int64)
        assert not np.isnan(part[touched]).any()
        assert np.abs(part[touched] - full[touched]).max() <= 1e-15
        untouched = np.setdiff1d(np.arange(b.dimension()), touched)
        assert np.isnan(part[untouched]).all()
