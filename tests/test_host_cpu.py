"""CPU suite, part 2: the product's host logic (tree flattening into per-leaf affine index
maps, size(prefix), node tree, error codes) against the oracle, and the C-ABI surface.
No compute call needs a GPU here; on a machine without one the compute entry points must
fail loudly with DFX_ERR_CUDA instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import orc
from cases import GRIDS, trees
from dune_functions_b200 import _capi as capi
from dune_functions_b200 import (YaspGrid, GlobalBasis, lagrange, lagrangeDG, power, composite, tree_array,
                                 blockedInterleaved as BI, flatInterleaved as FI)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS_SRC = os.path.join(ROOT, "tests", "harness", "host_index_harness.cpp")
HARNESS_SO = os.path.join(ROOT, "tests", "harness", "_build_harness.so")
BASIS_SRC = os.path.join(ROOT, "dune_functions_b200", "csrc", "dfx_basis.cu")


@pytest.fixture(scope="module")
def harness():
    srcs = [HARNESS_SRC, BASIS_SRC, os.path.join(ROOT, "dune_functions_b200", "csrc", "dfx_internal.h"),
            os.path.join(ROOT, "dune_functions_b200", "csrc", "dfx_hostbands.h")]
    if not os.path.exists(HARNESS_SO) or any(os.path.getmtime(s) > os.path.getmtime(HARNESS_SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", HARNESS_SO,
                               HARNESS_SRC, BASIS_SRC])
    L = C.CDLL(HARNESS_SO)
    L.harness_indices.argtypes = [C.POINTER(capi.GridDesc), C.POINTER(capi.TreeNode), C.c_int, C.c_void_p, C.c_int,
                                  C.c_void_p, C.POINTER(C.c_uint64)]
    return L


def lex_rank(dig, lens):
    """flat container position = lexicographic rank of the multi-index among all indices"""
    ne, nl, st = dig.shape
    keys = np.full((ne * nl, st), -1, dtype=np.int64)
    d = dig.reshape(-1, st).astype(np.int64)
    ln = lens.reshape(-1)
    for p in range(st):
        keys[:, p] = np.where(p < ln, d[:, p], -1)
    uniq, inv = np.unique(keys, axis=0, return_inverse=True)
    return inv.reshape(ne, nl)


@pytest.mark.parametrize("gname", list(GRIDS))
def test_flattened_index_maps_match_oracle(harness, gname):
    grid = GRIDS[gname]
    for name, tree in trees(grid.dim).items():
        if "TH_" in name and grid.dim != 3:
            continue
        ob = orc.OracleBasis(grid, tree)
        dig, lens = ob.indices()
        st = ob.maxMultiIndexSize()
        ne, nl = dig.shape[:2]
        hd = np.zeros((ne, nl, st), dtype=np.uint64)
        hf = np.zeros((ne, nl), dtype=np.uint64)
        dim = C.c_uint64()
        desc = grid.desc()
        arr, n = tree_array(tree)
        rc = harness.harness_indices(C.byref(desc), arr, n, hd.ctypes.data, st, hf.ctypes.data, C.byref(dim))
        assert rc == 0, name
        assert dim.value == ob.dimension(), name
        assert (hd == dig).all(), f"{gname}/{name}: multi-index digits differ"
        assert (hf.astype(np.int64) == lex_rank(dig, lens)).all(), f"{gname}/{name}: flat positions differ"
        assert (hf == ob.flatIndices()).all(), name


def test_partitioned_box_numbering_equals_local_grid(harness):
    """a slab of a larger grid is numbered like a grid of the slab's size"""
    tree = composite(power(lagrange(2), 3, BI()), lagrange(1))
    whole = YaspGrid([3, 2, 8])
    slab = whole.slab(1, 4)
    assert list(slab.local_begin) == [0, 0, 2] and list(slab.local_n) == [3, 2, 2]
    ref = orc.OracleBasis(YaspGrid([3, 2, 2]), tree)
    dig, _ = ref.indices()
    st = ref.maxMultiIndexSize()
    hd = np.zeros_like(dig)
    hf = np.zeros(dig.shape[:2], dtype=np.uint64)
    dim = C.c_uint64()
    desc = slab.desc()
    arr, n = tree_array(tree)
    assert harness.harness_indices(C.byref(desc), arr, n, hd.ctypes.data, st, hf.ctypes.data, C.byref(dim)) == 0
    assert (hd == dig).all() and dim.value == ref.dimension()


# ---- the shared library itself --------------------------------------------------------
def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dfx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dfx_[a-z0-9_]+)\s*\(", hdr)))


@pytest.mark.parametrize("gname", ["1d_5", "2d_3x2", "3d_3x4x5", "3d_2x1x3"])
def test_pipeline_bands_upload_what_their_elements_gather(harness, gname):
    """dfx_evaluate_host cuts the element range into bands and uploads, per band, only the
    coefficient ranges not yet on the device (csrc/dfx_hostbands.h).  With the oracle's index table:
    every coefficient an element of band i gathers is uploaded by band i or an earlier one, nothing
    is uploaded twice, and everything that is gathered at all gets uploaded."""
    grid = GRIDS[gname]
    desc = grid.desc()
    rng = np.random.default_rng(3)
    harness.harness_band_uploads.argtypes = [C.POINTER(capi.GridDesc), C.POINTER(capi.TreeNode), C.c_int, C.c_void_p,
                                             C.c_int, C.c_void_p]
    for name, tree in trees(grid.dim).items():
        o = orc.OracleBasis(grid, tree)
        ne = o.numElements()
        flat = o.flatIndices().astype(np.int64)
        arr, n = tree_array(tree)
        for trial in range(3):
            # element-range boundaries: whole layers, ragged cuts inside layers, single elements
            cuts = sorted(set([0, ne] + rng.integers(0, ne + 1, size=rng.integers(0, 5)).tolist()))
            bounds = np.array(cuts, dtype=np.uint64)
            band = np.zeros(o.dimension(), dtype=np.int32)
            rc = harness.harness_band_uploads(C.byref(desc), arr, n, bounds.ctypes.data, len(cuts) - 1, band.ctypes.data)
            assert rc == 0, (name, rc)
            for i in range(len(cuts) - 1):
                used = np.unique(flat[cuts[i]:cuts[i + 1]])
                assert (band[used] >= 0).all() and (band[used] <= i).all(), (name, cuts, i)
            assert (band[np.unique(flat)] >= 0).all()


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/dfx.h but not exported"
        assert name in capi.SIGNATURES, f"{name} has no ctypes signature"
    assert L.dfx_abi_version() == 1


@pytest.mark.parametrize("gname", ["2d_3x2", "3d_3x4x5"])
def test_global_basis_queries_match_oracle(gname):
    grid = GRIDS[gname]
    for name, tree in trees(grid.dim).items():
        if "TH_" in name and grid.dim != 3:
            continue
        b = GlobalBasis(grid, tree)
        o = orc.OracleBasis(grid, tree)
        assert b.dimension() == o.dimension()
        assert b.maxNodeSize() == o.maxNodeSize()
        assert b.maxMultiIndexSize() == o.maxMultiIndexSize()
        assert b.minMultiIndexSize() == o.minMultiIndexSize()
        assert b.numElements() == o.numElements()
        assert b.numLeaves() == o.numLeaves()
        assert b.numTreeNodes() == o.numTreeNodes()
        for node in range(b.numTreeNodes()):
            assert b.nodeInfo(node) == o.nodeInfo(node)
        # size(prefix) for every prefix of a sample of multi-indices (basistest.hh:150-180)
        dig, lens = o.indices(0, min(3, o.numElements()))
        seen = set()
        for e in range(dig.shape[0]):
            for i in range(dig.shape[1]):
                m = tuple(int(x) for x in dig[e, i, :lens[e, i]])
                for p in range(len(m) + 1):
                    if m[:p] not in seen:
                        seen.add(m[:p])
                        assert b.size(m[:p]) == o.size(m[:p]), (name, m[:p])
        assert (b.localIndexSizes() == lens[0]).all()


def test_error_codes():
    g = YaspGrid([2, 2])
    with pytest.raises(capi.RangeError):
        GlobalBasis(g, lagrange(18))               # lagrangebasis.hh:795-796
    with pytest.raises(capi.RangeError):
        GlobalBasis(g, lagrange(-1))
    with pytest.raises(capi.DfxError):
        GlobalBasis(g, composite(lagrange(1), lagrange(2), FI()))   # composite has no interleaved strategy
    with pytest.raises(capi.DfxError):
        GlobalBasis(YaspGrid([2, 0]), lagrange(1))
    with pytest.raises(capi.DfxError):
        GlobalBasis(YaspGrid([2, 2], upper=[1.0, -1.0]), lagrange(1))
    b = GlobalBasis(g, power(lagrange(1), 2, BI()))
    with pytest.raises(capi.RangeError):
        b.child(0, 5)
    assert b.child(0, 1) == 2


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    b = GlobalBasis(YaspGrid([2, 2]), lagrange(2))
    L = capi.lib()
    buf = (C.c_double * b.dimension())()
    f = capi.Function()
    f.id, f.n_comp = capi.FN_GAUSSIAN, 1
    rc = L.dfx_interpolate_host(b._h, 0, C.byref(f), buf, None)
    assert rc == capi.DFX_ERR_CUDA
    assert b"no CPU fallback" in L.dfx_last_error() or b"CUDA" in L.dfx_last_error()
    xh = (C.c_double * 2)(0.5, 0.5)
    out = (C.c_double * 4)()
    assert L.dfx_evaluate_host(b._h, 0, buf, xh, 1, 0, 4, out, None) == capi.DFX_ERR_CUDA
    idx = (C.c_uint32 * 36)()
    assert L.dfx_indices_flat_host(b._h, 0, 4, idx) == capi.DFX_ERR_CUDA
    # the newer entry points: boundary mask, operator set-up, peer link
    mask = (C.c_uint8 * b.dimension())()
    assert L.dfx_boundary_dofs_host(b._h, 0, mask) == capi.DFX_ERR_CUDA
    row_ptr = (C.c_uint64 * (b.dimension() + 1))()
    nnz = C.c_uint64()
    assert L.dfx_matrix_pattern_host(b._h, row_ptr, None, C.byref(nnz)) == capi.DFX_ERR_CUDA
    col = (C.c_uint32 * 4)()
    assert L.dfx_assemble_laplace_host(b._h, row_ptr, col, buf, C.byref(f), buf) == capi.DFX_ERR_CUDA
    link = C.c_void_p()
    assert L.dfx_halo_link_create(b._h, C.byref(link)) == capi.DFX_ERR_CUDA
    # host bookkeeping still works without a device
    n = C.c_int32()
    lst = (C.c_int32 * b.maxNodeSize())()
    assert L.dfx_subentity_dofs(b._h, 0, 1, 1, None, lst, C.byref(n)) == capi.DFX_OK and n.value == 3


@pytest.mark.parametrize("n", [(3, 2), (1, 1), (5, 4), (2, 2, 2), (3, 2, 4), (1, 1, 1), (2, 1, 3)], ids=str)
def test_face_dof_permutation_of_the_kernels_matches_oracle(harness, n):
    """layoutIndexOriented / orientFaceDOF / dofOwnerNode (csrc/dfx_internal.h, what k_indices, k_interpolate,
    k_points, k_eval_generic run when vertex ids are set) walked on the CPU: indices equal the oracle's
    FaceOrientations / LagrangeFaceDOFPermutation restatement bit for bit, every DOF's owner node maps
    back to the DOF, and the inverse permutation inverts."""
    harness.harness_oriented.restype = C.c_longlong
    harness.harness_oriented.argtypes = [C.POINTER(capi.GridDesc), C.POINTER(capi.TreeNode), C.c_int, C.c_void_p, C.c_void_p]
    dim = len(n)
    nv = int(np.prod([m + 1 for m in n]))
    some_trees = {"Q1": lagrange(1), "Q2": lagrange(2), "Q3": lagrange(3), "Q4": lagrange(4), "Q5": lagrange(5), "Q7": lagrange(7),
                  "powFI_Q3": power(lagrange(3), dim, FI()), "comp_Q4_Q2_DG2": composite(lagrange(4), lagrange(2), lagrangeDG(2))}
    for seed in range(3):
        ids = np.random.default_rng(seed).permutation(nv).astype(np.uint64) * 3 + 2
        for slab in (None, (0, 2), (1, 2)):
            whole = YaspGrid(list(n), vertex_ids=ids)
            if slab is not None:
                if n[-1] < 2:
                    continue
                whole = whole.slab(*slab)
            for name, tree in some_trees.items():
                ob = orc.OracleBasis(whole, tree)
                ref = ob.flatIndices()
                out = np.zeros(ref.shape, dtype=np.uint64)
                desc = whole.desc()
                arr, cnt = tree_array(tree)
                vid = np.ascontiguousarray(whole.vertex_ids, dtype=np.uint64)
                bad = harness.harness_oriented(C.byref(desc), arr, cnt, vid.ctypes.data, out.ctypes.data)
                assert bad == 0, (name, seed, slab, bad)
                assert (out == ref).all(), (name, seed, slab)


def test_update_moves_the_handle_to_another_grid():
    """basis.update(gridView) (defaultglobalbasis.hh:137-141): same tree, new grid; host bookkeeping only here"""
    from dune_functions_b200 import taylorHood
    for tree in (lagrange(3), taylorHood(), composite(power(lagrange(2), 3, BI()), lagrangeDG(1))):
        a, bgrid = YaspGrid([2, 3, 2]), YaspGrid([4, 2, 5], upper=[1.0, 2.0, 0.5])
        b = GlobalBasis(a, tree)
        fresh = GlobalBasis(bgrid, tree)
        assert b.dimension() != fresh.dimension()
        b.update(bgrid)
        assert b.dimension() == fresh.dimension() and b.numElements() == fresh.numElements() == 40
        assert b.containerDescriptor().type() == fresh.containerDescriptor().type()
        ob = orc.OracleBasis(bgrid, tree)
        assert b.dimension() == ob.dimension()
        for prefix in ((), (0,), (1,), (0, 0)):
            assert b.size(prefix) == ob.size(prefix)
        # a descriptor the library rejects leaves the handle as it was
        bad = bgrid.desc()
        bad.n_global[1] = 0
        bad.local_n[1] = 0
        assert capi.lib().dfx_basis_update(b._h, C.byref(bad)) != 0
        assert b.dimension() == fresh.dimension()


def test_vertex_ids_are_host_bookkeeping_until_a_device_call():
    """dfx_basis_set_vertex_ids_host needs no device (slab set-up on a CPU-only rank, tests/test_dist_cpu.py);
    wrong counts are a RangeError either way"""
    grid = YaspGrid([3, 2, 2], vertex_ids=np.random.default_rng(0).permutation(36))
    b = GlobalBasis(grid, lagrange(4))
    assert b.hasPermutedFaceDOFs()
    assert not GlobalBasis(grid, lagrange(2)).hasPermutedFaceDOFs()
    assert not GlobalBasis(YaspGrid([5], vertex_ids=range(6)), lagrange(4)).hasPermutedFaceDOFs()
    with pytest.raises(capi.RangeError):
        b.setVertexIds(np.arange(35))
    b.setVertexIds(None)
    assert not b.hasPermutedFaceDOFs()
    # slabs take contiguous runs of the id table
    s1 = grid.slab(1, 2)
    assert list(s1.vertex_ids) == list(grid.vertex_ids[12:36]) and list(grid.slab(0, 2).vertex_ids) == list(grid.vertex_ids[:24])


def test_kernel_orientation_code_on_all_24_orderings_of_every_facet(harness):
    """the truth table's 24 orderings (tests/golden/face_orientations.json) put on each of the three facet
    directions of a single cube in turn, orders 3-6: the kernels' orientFaceDOF (direct search of the smallest
    id) gives the oracle's table-driven numbering for every one of them"""
    import itertools
    harness.harness_oriented.restype = C.c_longlong
    harness.harness_oriented.argtypes = [C.POINTER(capi.GridDesc), C.POINTER(capi.TreeNode), C.c_int, C.c_void_p, C.c_void_p]
    facets = {"z": ((0, 1, 2, 3), (4, 5, 6, 7)), "y": ((0, 1, 4, 5), (2, 3, 6, 7)), "x": ((0, 2, 4, 6), (1, 3, 5, 7))}
    seen = set()
    for name, (lo, hi) in facets.items():
        for perm in itertools.permutations(range(4)):
            ids = np.zeros(8, dtype=np.uint64)
            for i in range(4):
                ids[lo[i]] = 10 + perm[i]
                ids[hi[i]] = 3 - perm[i]          # the opposite facet runs through the mirrored orderings
            grid = YaspGrid([1, 1, 1], vertex_ids=ids)
            for k in (3, 4, 5, 6):
                tree = lagrange(k)
                ob = orc.OracleBasis(grid, tree)
                ref = ob.flatIndices()
                out = np.zeros(ref.shape, dtype=np.uint64)
                desc = grid.desc()
                arr, cnt = tree_array(tree)
                bad = harness.harness_oriented(C.byref(desc), arr, cnt, ids.ctypes.data, out.ctypes.data)
                assert bad == 0 and (out == ref).all(), (name, perm, k)
            seen.add((name, orc.faceOrientations(3, ids)[1][{"z": 4, "y": 2, "x": 0}[name]]))
    # all eight quadrilateral orientations occurred on every facet direction
    assert seen == {(n, o) for n in facets for o in range(8, 16)}


@pytest.mark.parametrize("gname", list(GRIDS))
def test_every_dof_owner_node_maps_back_to_the_dof(harness, gname):
    """dofOwnerNode (what k_interpolate / k_points use to find the node behind a DOF) and layoutIndex are
    inverse to each other on every grid / tree of the parity suite and on z-slabs of them — identity
    orientations here, the permuted case is test_face_dof_permutation_of_the_kernels_matches_oracle"""
    harness.harness_oriented.restype = C.c_longlong
    harness.harness_oriented.argtypes = [C.POINTER(capi.GridDesc), C.POINTER(capi.TreeNode), C.c_int, C.c_void_p, C.c_void_p]
    whole = GRIDS[gname]
    grids = [whole]
    if whole.n[-1] >= 2:
        grids += [whole.slab(0, 2), whole.slab(1, 2)]
    for grid in grids:
        for name, tree in trees(grid.dim).items():
            if "TH_" in name and grid.dim != 3:
                continue
            ob = orc.OracleBasis(YaspGrid(list(grid.local_n)) if grid.local_n is not None else grid, tree)
            ref = ob.flatIndices()
            out = np.zeros(ref.shape, dtype=np.uint64)
            desc = grid.desc()
            arr, cnt = tree_array(tree)
            bad = harness.harness_oriented(C.byref(desc), arr, cnt, None, out.ctypes.data)
            assert bad == 0, (gname, name, bad)
            assert (out == ref).all(), (gname, name)
