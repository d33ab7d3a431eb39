"""CPU suite for SURVEY.md section 8(f) row 3: continuous Lagrange bases of order >= 3 on a grid whose
vertex ids are NOT monotone in lexicographic position, i.e. Experimental::FaceOrientations and
LagrangeFaceDOFPermutation (lagrangebasis.hh:207-697).

Pins of the oracle's restatement:
  * the truth tables in the doc comment of FaceOrientations (:132-174), committed literally as
    tests/golden/face_orientations.json;
  * what lagrangebasistest.cc:134-310 checks with checkBasis(basis, EnableContinuityCheck()): the
    basis is conforming — here in the equivalent nodal form "every global index belongs to exactly one
    lattice node, whichever element looks at it";
  * interpolation of a polynomial of degree <= k is reproduced exactly (interpolatetest.hh:45-67)."""
import itertools
import json
import os

import numpy as np
import pytest

import orc
from dune_functions_b200 import YaspGrid, lagrange, power, quadratic, coord

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TABLE = json.load(open(os.path.join(GOLDEN, "face_orientations.json")))


@pytest.mark.parametrize("row", TABLE["quadrilateral"], ids=lambda r: "".join(map(str, r["vertex_ids"])))
def test_quadrilateral_truth_table(row):
    ids = row["vertex_ids"]
    # the quadrilateral as the facet z = 0 (number 4, vertices 0..3) of a cube; its edges in the facet's
    # own numbering x'=0, x'=1, y'=0, y'=1 are the cube's edges 4, 5, 6, 7  [EXT reference cube]
    edges, facets = orc.faceOrientations(3, ids + [4, 5, 6, 7])
    assert [edges[e] for e in (4, 5, 6, 7)] == row["edge_orientations"]
    assert facets[4] == row["index"]
    assert format(facets[4], "04b") == row["bits"]
    assert (facets[4] & 3) == row["rotations"] and bool(facets[4] & 4) == row["reflection"]
    # the same quadrilateral as a 2-d element: only edge bits exist
    edges2, _ = orc.faceOrientations(2, ids)
    assert edges2 == row["edge_orientations"]
    # any order-preserving relabelling of the ids gives the same orientation ("any selection of pairwise
    # different less-than comparable ids can be used")
    big = [1000 + 37 * i for i in ids]
    assert orc.faceOrientations(3, big + [7, 5, 3, 1])[1][4] == row["index"]


def test_all_24_orderings_are_covered_and_opposite_facets_agree():
    seen = {tuple(r["vertex_ids"]) for r in TABLE["quadrilateral"]}
    assert seen == set(itertools.permutations(range(4)))
    # facets 4 (z=0) and 5 (z=1) with the same id pattern have the same orientation; likewise x and y
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = rng.permutation(4)
        for lo, hi, verts in ((4, 5, ((0, 1, 2, 3), (4, 5, 6, 7))), (2, 3, ((0, 1, 4, 5), (2, 3, 6, 7))),
                              (0, 1, ((0, 2, 4, 6), (1, 3, 5, 7)))):
            ids = np.zeros(8, dtype=np.uint64)
            for i in range(4):
                ids[verts[0][i]] = q[i]
                ids[verts[1][i]] = 10 + q[i]
            f = orc.faceOrientations(3, ids)[1]
            assert f[lo] == f[hi]


@pytest.mark.parametrize("order", [2, 3, 4, 5, 8])
def test_quad_dof_tables_are_the_dihedral_group(order):
    m = order - 1
    tabs = [orc.quadDofTable(order, j) for j in range(8)]
    assert (tabs[0] == np.arange(m * m)).all()
    for t in tabs:
        assert sorted(t.tolist()) == list(range(m * m))          # a permutation
    if m > 1:
        assert len({tuple(t.tolist()) for t in tabs}) == 8       # all eight distinct
    r1, r3 = tabs[1], tabs[3]
    assert (r3[r1] == np.arange(m * m)).all()                     # one rotation and three rotations are inverse
    assert (tabs[2][tabs[2]] == np.arange(m * m)).all()
    assert (tabs[4][tabs[4]] == np.arange(m * m)).all()           # the reflection is an involution


def lattice_nodes(grid, k):
    """global lattice coordinates [ne, nloc, dim] of the local nodes (LagrangeCube: i = sum alpha_j (k+1)^j)"""
    N = grid.n
    dim = len(N)
    ne = int(np.prod(N))
    nloc = (k + 1) ** dim
    e = np.arange(ne)
    ec = np.zeros((ne, dim), dtype=np.int64)
    for j in range(dim):
        ec[:, j] = e % N[j]
        e = e // N[j]
    i = np.arange(nloc)
    al = np.zeros((nloc, dim), dtype=np.int64)
    for j in range(dim):
        al[:, j] = i % (k + 1)
        i = i // (k + 1)
    return k * ec[:, None, :] + al[None, :, :]


def random_ids(grid, seed):
    nv = int(np.prod([n + 1 for n in grid.n]))
    return np.random.default_rng(seed).permutation(nv).astype(np.uint64) * 3 + 5


def check_conforming(ob, grid, k):
    idx = ob.flatIndices().astype(np.int64)
    nodes = lattice_nodes(grid, k)
    dim = nodes.shape[-1]
    L = [k * n + 1 for n in grid.n]
    key = np.zeros(idx.shape, dtype=np.int64)
    stride = 1
    for j in range(dim):
        key += nodes[..., j] * stride
        stride *= L[j]
    n = ob.dimension()
    assert n == int(np.prod(L))
    owner = np.full(n, -1, dtype=np.int64)
    owner[idx.ravel()] = key.ravel()
    assert (owner >= 0).all()                                     # consecutive: every index is used
    assert (owner[idx] == key).all()                              # every element sees the same node behind an index
    assert len(np.unique(owner)) == n                             # and no two indices share a node
    for row in idx:
        assert len(set(row.tolist())) == len(row)                 # local indices unique (basistest.hh)


GRIDS = [(3, 2), (1, 1), (4, 3), (2, 2, 2), (3, 2, 2), (1, 1, 1), (2, 1, 3)]


@pytest.mark.parametrize("N", GRIDS, ids=str)
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_basis_is_conforming_for_any_vertex_numbering(N, k):
    grid = YaspGrid(list(N))
    ob = orc.OracleBasis(grid, lagrange(k))
    base = ob.flatIndices().copy()
    check_conforming(ob, grid, k)
    changed = False
    for seed in range(4):
        ob.setVertexIds(random_ids(grid, seed))
        check_conforming(ob, grid, k)
        cur = ob.flatIndices()
        changed = changed or (cur != base).any()
        if k <= 2:
            assert (cur == base).all()                            # nothing to permute below order 3
    if k >= 3 and len(N) > 1:
        assert changed                                            # the permutation is visible
    # monotone ids, given explicitly, are the default numbering
    nv = int(np.prod([n + 1 for n in N]))
    ob.setVertexIds(np.arange(nv, dtype=np.uint64) * 7 + 1)
    assert (ob.flatIndices() == base).all()
    ob.setVertexIds(None)
    assert (ob.flatIndices() == base).all()


def test_without_the_permutation_the_basis_would_not_conform():
    """the check above has teeth: swapping two ids changes the index table exactly on the entities that
    contain both vertices, and using the unpermuted table with the permuted one mixes nodes"""
    grid = YaspGrid([2, 2])
    ob = orc.OracleBasis(grid, lagrange(4))
    base = ob.flatIndices().copy()
    ids = np.arange(9, dtype=np.uint64)
    ids[[0, 1]] = ids[[1, 0]]                                     # flips the bottom edge of element 0
    ob.setVertexIds(ids)
    cur = ob.flatIndices()
    diff = np.argwhere(cur != base)
    assert set(diff[:, 0].tolist()) == {0}                        # only element 0 contains that edge
    assert sorted(diff[:, 1].tolist()) == [1, 3]                  # alpha = (1,0) and (3,0): the edge's outer DOFs swap
    assert cur[0, 1] == base[0, 3] and cur[0, 3] == base[0, 1] and cur[0, 2] == base[0, 2]


@pytest.mark.parametrize("N,k", [((3, 2), 3), ((3, 2), 5), ((2, 2, 2), 3), ((2, 1, 2), 4)], ids=str)
def test_polynomials_are_reproduced_with_permuted_numbering(N, k):
    grid = YaspGrid(list(N))
    dim = len(N)
    ob = orc.OracleBasis(grid, lagrange(k))
    ob.setVertexIds(random_ids(grid, 11))
    f = quadratic(0.5, [1.0, -2.0, 3.0], [42.0, 7.0, 1.0, 13.0, -3.0, 5.0])
    x = ob.interpolate(f)
    rng = np.random.default_rng(3)
    pts = rng.random((7, dim))
    v, j = ob.evaluate(x, pts, jacobians=True)
    # the same function through the default numbering: same point values, different coefficient order
    ob0 = orc.OracleBasis(grid, lagrange(k))
    x0 = ob0.interpolate(f)
    v0, j0 = ob0.evaluate(x0, pts, jacobians=True)
    assert np.abs(v - v0).max() <= 1e-12 * max(1.0, np.abs(v0).max())
    assert np.abs(j - j0).max() <= 1e-11 * max(1.0, np.abs(j0).max())
    assert sorted(x.tolist()) == sorted(x0.tolist())
    assert (x != x0).any()
