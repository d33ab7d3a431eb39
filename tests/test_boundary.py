"""Boundary DOFs — SURVEY.md section 8(f) row 1: forEachBoundaryDOF (boundarydofs.hh:39-127) and
SubEntityDOFs (subentitydofs.hh:45-236).

CPU part: the oracle restatement against known answers (a Lagrange DOF is visited iff its node lies
on the boundary of the domain; counts of nodes on the surface of the lattice) and the host-side
dfx_subentity_dofs against the oracle.  GPU part: dfx_boundary_dofs bit-exact against the oracle
for every grid x tree, sub-space bases, slab partitions, and the Dirichlet flow of
examples/interpolation.cc:99-122 (mark velocity boundary DOFs, masked interpolation)."""
import ctypes as C

import numpy as np
import pytest

import orc
from cases import GRIDS, trees, th_variant
import dune_functions_b200 as dfx
from dune_functions_b200 import (YaspGrid, lagrange, lagrangeDG, power, composite, taylorHood,
                                 flatLexicographic as FL, blockedLexicographic as BL, blockedInterleaved as BI,
                                 coord, identity)
from dune_functions_b200 import _capi as capi


def node_on_boundary(ob, grid):
    """geometric criterion from the interpolated coordinate functions (scalar bases)"""
    d = grid.dim
    lo = grid.lower if grid.lower is not None else [0.0] * d
    up = grid.upper if grid.upper is not None else [1.0] * d
    on = np.zeros(ob.dimension(), dtype=bool)
    for j in range(d):
        xj = ob.interpolate(coord(j))
        on |= np.isclose(xj, lo[j], atol=1e-13) | np.isclose(xj, up[j], atol=1e-13)
    return on


# ---- oracle pinned by known answers ---------------------------------------------------------
@pytest.mark.parametrize("gname", list(GRIDS))
@pytest.mark.parametrize("tname", ["lagrange1", "lagrange2", "lagrange3", "lagrange4", "lagrangeDG1", "lagrangeDG2"])
def test_oracle_boundary_dofs_are_the_nodes_on_the_domain_boundary(gname, tname):
    grid = GRIDS[gname]
    ob = orc.OracleBasis(grid, trees(grid.dim)[tname])
    mask, visits = ob.boundaryDOFs()
    assert (mask.astype(bool) == node_on_boundary(ob, grid)).all()
    assert visits >= mask.sum()


@pytest.mark.parametrize("n,k", [([4], 1), ([4], 3), ([3, 2], 1), ([3, 2], 2), ([3, 4, 5], 1), ([3, 4, 5], 2), ([2, 2, 2], 3)])
def test_oracle_boundary_dof_count_is_lattice_surface(n, k):
    ob = orc.OracleBasis(YaspGrid(n), lagrange(k))
    mask, _ = ob.boundaryDOFs()
    total = np.prod([k * x + 1 for x in n])
    inner = np.prod([k * x - 1 for x in n])
    assert mask.sum() == total - inner


@pytest.mark.parametrize("tname", ["lagrange0", "lagrangeDG0"])
def test_oracle_order_zero_has_no_boundary_dofs(tname):
    ob = orc.OracleBasis(YaspGrid([3, 2]), trees(2)[tname])
    mask, visits = ob.boundaryDOFs()
    assert mask.sum() == 0 and visits == 0


def test_oracle_visit_count_single_element():
    # 1x1 grid, Q1: 4 boundary intersections x 2 vertices each; Q2: 4 x 3
    assert orc.OracleBasis(YaspGrid([1, 1]), lagrange(1)).boundaryDOFs()[1] == 8
    assert orc.OracleBasis(YaspGrid([1, 1]), lagrange(2)).boundaryDOFs()[1] == 12
    # 1x1x1 grid, Q2: 6 faces x 9
    assert orc.OracleBasis(YaspGrid([1, 1, 1]), lagrange(2)).boundaryDOFs()[1] == 54


def test_oracle_subentity_dofs_q2_cube():
    ob = orc.OracleBasis(YaspGrid([1, 1, 1]), lagrange(2))
    # face 0 (x = 0): lattice points with alpha_0 = 0, local index a + 3b + 9c
    lst, flags = ob.subEntityDOFs(0, 1)
    assert lst == [i for i in range(27) if i % 3 == 0]
    assert flags.sum() == 9
    # face 5 (z = 1)
    assert ob.subEntityDOFs(5, 1)[0] == list(range(18, 27))
    # the element itself contains everything, a vertex only itself
    assert ob.subEntityDOFs(0, 0)[0] == list(range(27))
    assert ob.subEntityDOFs(7, 3)[0] == [26]
    # edge 0: z-directed at (x, y) = (0, 0)
    assert ob.subEntityDOFs(0, 2)[0] == [0, 9, 18]
    # edge 6: x-directed at y = 0, z = 0
    assert ob.subEntityDOFs(6, 2)[0] == [0, 1, 2]


def test_oracle_taylor_hood_velocity_subspace():
    """examples/interpolation.cc:99-108: only velocity DOFs are marked"""
    grid = YaspGrid([2, 2])
    tree = th_variant(FL, FL, 2)
    ob = orc.OracleBasis(grid, tree)
    n2 = 25
    mask, _ = ob.boundaryDOFs(node=1)         # expanded tree: 0 composite, 1 power, 2,3 Q2 leaves, 4 Q1
    assert mask[:2 * n2].sum() == 2 * 16 and mask[2 * n2:].sum() == 0
    pmask, _ = ob.boundaryDOFs(node=4)
    assert pmask[:2 * n2].sum() == 0 and pmask[2 * n2:].sum() == 8


# ---- host bookkeeping of the library (no GPU needed) ------------------------------------------
@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_2x1x3"])
def test_subentity_dofs_host_matches_oracle(gname):
    grid = GRIDS[gname]
    d = grid.dim
    nsub = {1: {0: 1, 1: 2}, 2: {0: 1, 1: 4, 2: 4}, 3: {0: 1, 1: 6, 2: 12, 3: 8}}[d]
    for tname, tree in trees(d).items():
        b = dfx.makeBasis(grid, tree)
        ob = orc.OracleBasis(grid, tree)
        nodes = range(b.numTreeNodes())
        for codim, cnt in nsub.items():
            for i in range(cnt):
                for node in nodes:
                    root = b if node == 0 else _Sub(b, node)
                    got = dfx.subEntityDOFs(root, i, codim)
                    exp = ob.subEntityDOFs(i, codim, node=node)
                    assert got[0] == exp[0], (tname, codim, i, node)
                    assert (got[1] == exp[1]).all()


class _Sub(dfx.SubspaceBasis):
    def __init__(self, root, node):
        self.root, self.node, self.path = root, node, ()


def test_subentity_dofs_bad_subentity_raises():
    b = dfx.makeBasis(YaspGrid([2, 2]), lagrange(2))
    with pytest.raises(capi.RangeError):
        dfx.subEntityDOFs(b, 4, 1)
    with pytest.raises(capi.RangeError):
        dfx.subEntityDOFs(b, 0, 3)


# ---- GPU parity ---------------------------------------------------------------------------------
def all_cases():
    for gname, grid in GRIDS.items():
        for tname in trees(grid.dim):
            yield pytest.param(gname, tname, id=f"{gname}-{tname}")


@pytest.mark.gpu
@pytest.mark.parametrize("gname,tname", list(all_cases()))
def test_boundary_dofs_bit_exact(gname, tname):
    grid = GRIDS[gname]
    tree = trees(grid.dim)[tname]
    b = dfx.makeBasis(grid, tree)
    ob = orc.OracleBasis(grid, tree)
    got = dfx.boundaryDOFs(b).cpu().numpy()
    exp, _ = ob.boundaryDOFs()
    assert (got == exp).all()
    # every sub-space basis (subspaceBasis(basis, path)) marks only its own leaves
    for node in range(1, b.numTreeNodes()):
        got = dfx.boundaryDOFs(_Sub(b, node)).cpu().numpy()
        exp, _ = ob.boundaryDOFs(node=node)
        assert (got == exp).all(), node


@pytest.mark.gpu
def test_boundary_dofs_keeps_other_mask_entries_and_host_entry_point():
    import torch
    grid = GRIDS["3d_3x4x5"]
    tree = taylorHood()
    b = dfx.makeBasis(grid, tree)
    ob = orc.OracleBasis(grid, tree)
    pre = (np.arange(b.dimension()) % 3 == 0).astype(np.uint8) * 7
    exp = pre.copy()
    bm, _ = ob.boundaryDOFs(node=1)
    exp[bm.astype(bool)] = 1
    m = torch.from_numpy(pre.copy()).cuda()
    dfx.boundaryDOFs(dfx.subspaceBasis(b, 0), m)
    assert (m.cpu().numpy() == exp).all()
    hm = pre.copy()
    capi.check(b._lib.dfx_boundary_dofs_host(b._h, b.child(0, 0), C.c_void_p(hm.ctypes.data)))
    assert (hm == exp).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tname", ["lagrange2", "lagrangeDG2", "TH_BL(BI)"])
def test_boundary_dofs_of_slabs_use_the_global_boundary(tname):
    whole = YaspGrid([3, 2, 6], upper=[1.0, 0.5, 2.0])
    tree = trees(3)[tname]
    og = orc.OracleBasis(whole, tree)
    gmask, _ = og.boundaryDOFs()
    gflat = og.flatIndices().reshape(6, 6, -1).astype(np.int64)
    for world in (2, 3):
        nz = 6 // world
        for rank in range(world):
            mine = whole.slab(rank, world)
            b = dfx.makeBasis(mine, tree)
            got = dfx.boundaryDOFs(b).cpu().numpy()
            ol = orc.OracleBasis(mine, tree)
            exp, _ = ol.boundaryDOFs()
            assert (got == exp).all()
            # and it is the restriction of the serial mask to the slab
            lflat = b.flatIndices().cpu().numpy().reshape(nz, 6, -1).astype(np.int64)
            assert (got[lflat] == gmask[gflat[nz * rank:nz * (rank + 1)]]).all()


@pytest.mark.gpu
def test_dirichlet_flow_of_the_interpolation_example():
    """examples/interpolation.cc:99-122: mark the velocity boundary DOFs of a Taylor-Hood basis,
    then interpolate x -> x into the marked entries only."""
    grid = YaspGrid([4, 3])
    tree = th_variant(FL, FL, 2)
    b = dfx.makeBasis(grid, tree)
    ob = orc.OracleBasis(grid, tree)
    vel = dfx.subspaceBasis(b, 0)
    is_boundary = dfx.boundaryDOFs(vel)
    x = b.newVector(-3.0)
    dfx.interpolate(vel, x, identity(2), mask=is_boundary)
    em, _ = ob.boundaryDOFs(node=1)
    ex = ob.interpolate(identity(2), coeff=np.full(ob.dimension(), -3.0), mask=em, node=1)
    assert np.array_equal(x.cpu().numpy(), ex)
    assert (ex == -3.0).sum() == ob.dimension() - em.sum()
