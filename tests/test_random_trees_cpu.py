"""CPU suite: randomly generated basis trees (seeded) — the library's host-side flattening of the tree into
per-leaf affine index maps (csrc/dfx_basis.cu, walked by tests/harness) and its containerDescriptor() against
the oracle's recursive PreBasis::indices / size(prefix) (dynamicpowerbasis.hh:138-371, compositebasis.hh:184-338).
The fixed tree lists of the other tests cover the named cases; this one covers nestings nobody wrote down."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc
from dune_functions_b200 import _capi as capi
from dune_functions_b200 import (YaspGrid, GlobalBasis, lagrange, lagrangeDG, power, composite, tree_array,
                                 flatLexicographic as FL, flatInterleaved as FI,
                                 blockedLexicographic as BL, blockedInterleaved as BI)
from test_host_cpu import harness, lex_rank  # noqa: F401  (fixture + helper)

LEAVES = [lagrange(1), lagrange(2), lagrangeDG(1), lagrangeDG(0)]      # <= 4 distinct (kind, order) layouts


def random_tree(rng, depth, budget):
    """(factory, number of leaves, multi-index length); budget = leaves still allowed"""
    if depth == 0 or budget <= 1 or rng.random() < 0.25:
        return LEAVES[rng.integers(len(LEAVES))], 1, 1
    if rng.random() < 0.5:
        c = int(rng.integers(1, 4))
        sub, n, d = random_tree(rng, depth - 1, max(1, budget // c))
        ims = [FL(), FI(), BL(), BI()][rng.integers(4)]
        return power(sub, c, ims), n * c, d + (ims in (BL(), BI()))
    k = int(rng.integers(1, 4))
    subs, n, d = [], 0, 0
    for _ in range(k):
        s, ns, ds = random_tree(rng, depth - 1, max(1, budget // k))
        subs.append(s); n += ns; d = max(d, ds)
    ims = [FL(), BL()][rng.integers(2)]
    return composite(*subs, ims), n, d + (ims == BL())


def cases():
    rng = np.random.default_rng(20261020)
    out = []
    while len(out) < 120:
        tree, nleaves, digits = random_tree(rng, 3, 16)
        if nleaves <= 16 and digits <= 6:
            out.append(tree)
    return out


TREES = cases()


def describe(f):
    names = {capi.NODE_LAGRANGE: "Q", capi.NODE_LAGRANGE_DG: "DG", capi.NODE_POWER: "pow", capi.NODE_COMPOSITE: "comp"}
    return "/".join(f"{names[k]}{o if k in (capi.NODE_LAGRANGE, capi.NODE_LAGRANGE_DG) else c}.{m}" for k, o, c, m in f.nodes())


@pytest.mark.parametrize("chunk", range(6))
def test_random_trees_flattening_and_descriptor(harness, chunk):
    harness.harness_indices.argtypes = [C.POINTER(capi.GridDesc), C.POINTER(capi.TreeNode), C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.POINTER(C.c_uint64)]
    for t, tree in enumerate(TREES[chunk * 20:(chunk + 1) * 20]):
        grid = YaspGrid([2, 2]) if t % 2 == 0 else YaspGrid([2, 1, 2])
        what = describe(tree)
        ob = orc.OracleBasis(grid, tree)
        dig, lens = ob.indices()
        st = ob.maxMultiIndexSize()
        ne, nl = dig.shape[:2]
        # the product's flattening, walked on the CPU
        hd = np.zeros((ne, nl, st), dtype=np.uint64)
        hf = np.zeros((ne, nl), dtype=np.uint64)
        dim = C.c_uint64()
        desc = grid.desc()
        arr, n = tree_array(tree)
        assert harness.harness_indices(C.byref(desc), arr, n, hd.ctypes.data, st, hf.ctypes.data, C.byref(dim)) == 0, what
        assert dim.value == ob.dimension(), what
        assert (hd == dig).all(), what
        assert (hf.astype(np.int64) == lex_rank(dig, lens)).all(), what
        # host queries of the handle
        b = GlobalBasis(grid, tree)
        assert b.dimension() == ob.dimension() and b.maxNodeSize() == nl, what
        assert b.maxMultiIndexSize() == st and b.minMultiIndexSize() == ob.minMultiIndexSize(), what
        prefixes = {()}
        for row, ln in zip(dig.reshape(-1, st), lens.reshape(-1)):
            for j in range(ln + 1):
                prefixes.add(tuple(int(v) for v in row[:j]))
        for p in prefixes:
            assert b.size(p) == ob.size(p), (what, p)
        # containerDescriptor: known unless a flat composite sits on blocked children (containerdescriptors.hh:291-298)
        cd = b.containerDescriptor()
        if cd.kind == capi.CD_UNKNOWN:
            assert "comp" in what
            continue
        for p in prefixes:
            d = cd
            for digit in p:
                if d.kind == capi.CD_UNKNOWN:              # an Unknown sub-tree (below a blocked node) says nothing
                    break
                assert digit < d.size(), (what, p)
                d = d[digit]
            if d.kind != capi.CD_UNKNOWN:
                assert d.size() == ob.size(p), (what, p, cd.type())


def test_random_trees_cover_the_interesting_shapes():
    kinds = [describe(t) for t in TREES]
    assert sum("pow" in k and "comp" in k for k in kinds) >= 30            # mixed nestings
    assert any(GlobalBasis(YaspGrid([2, 2]), t).containerDescriptor().kind == capi.CD_UNKNOWN for t in TREES)
    assert any(GlobalBasis(YaspGrid([2, 2]), t).containerDescriptor().kind == capi.CD_TUPLE for t in TREES)
    assert max(GlobalBasis(YaspGrid([2, 2]), t).maxMultiIndexSize() for t in TREES) >= 4


def test_random_trees_with_permuted_face_dofs(harness):
    """the same for trees with continuous leaves of order 3 and 4 on grids with random vertex ids: the kernels'
    layoutIndexOriented / dofOwnerNode (walked on the CPU) against the oracle's FaceOrientations restatement"""
    global LEAVES
    harness.harness_oriented.restype = C.c_longlong
    harness.harness_oriented.argtypes = [C.POINTER(capi.GridDesc), C.POINTER(capi.TreeNode), C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(7)
    saved = LEAVES
    LEAVES = [lagrange(3), lagrange(4), lagrange(2), lagrangeDG(1)]
    try:
        done = 0
        while done < 40:
            tree, nleaves, digits = random_tree(rng, 2, 8)
            if nleaves > 8 or digits > 6:
                continue
            n = (2, 2) if done % 2 == 0 else (2, 1, 2)
            nv = int(np.prod([m + 1 for m in n]))
            ids = rng.permutation(nv).astype(np.uint64) + 100
            grid = YaspGrid(list(n), vertex_ids=ids)
            ob = orc.OracleBasis(grid, tree)
            ref = ob.flatIndices()
            out = np.zeros(ref.shape, dtype=np.uint64)
            desc = grid.desc()
            arr, cnt = tree_array(tree)
            bad = harness.harness_oriented(C.byref(desc), arr, cnt, ids.ctypes.data, out.ctypes.data)
            assert bad == 0 and (out == ref).all(), describe(tree)
            done += 1
    finally:
        LEAVES = saved
