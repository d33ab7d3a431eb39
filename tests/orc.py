"""ctypes binding of the CPU oracle (oracle/dfx_oracle.cpp).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from dune_functions_b200 import _capi as capi
from dune_functions_b200.basis import tree_array

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_PATH = os.path.join(ROOT, "oracle", "_build", "liborc.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORC_PATH):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(ORC_PATH)
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int32
        L.orc_basis_create.restype = vp
        L.orc_basis_create.argtypes = [C.POINTER(capi.GridDesc), C.POINTER(capi.TreeNode), i32]
        L.orc_basis_destroy.argtypes = [vp]
        L.orc_dimension.restype = u64
        L.orc_dimension.argtypes = [vp]
        L.orc_size_prefix.restype = u64
        L.orc_size_prefix.argtypes = [vp, C.POINTER(u64), i32]
        for n in ("orc_max_node_size", "orc_max_digits", "orc_min_digits", "orc_num_leaves", "orc_num_nodes"):
            getattr(L, n).restype = i32
            getattr(L, n).argtypes = [vp]
        L.orc_num_elements.restype = u64
        L.orc_num_elements.argtypes = [vp]
        L.orc_node_info.argtypes = [vp, i32, C.POINTER(i32)]
        L.orc_indices.argtypes = [vp, u64, u64, vp, vp, i32]
        L.orc_indices_flat.argtypes = [vp, u64, u64, vp]
        L.orc_interpolate.argtypes = [vp, i32, C.POINTER(capi.Function), vp, vp, i32]
        L.orc_interpolate_range.argtypes = [vp, i32, C.POINTER(capi.Function), vp, vp, u64, u64, i32]
        L.orc_indices_flat_by_enumeration.argtypes = [vp, u64, u64, vp]
        L.orc_evaluate.argtypes = [vp, i32, vp, vp, i32, u64, u64, vp, vp, i32]
        L.orc_max_threads.restype = i32
        L.orc_subentity_dofs.restype = i32
        L.orc_subentity_dofs.argtypes = [vp, i32, i32, i32, vp, vp]
        L.orc_boundary_dofs.restype = u64
        L.orc_boundary_dofs.argtypes = [vp, i32, vp]
        L.orc_interpolate_dgbf.restype = i32
        L.orc_interpolate_dgbf.argtypes = [vp, i32, vp, i32, vp, vp, vp]
        L.orc_evaluate_global.argtypes = [vp, i32, vp, vp, u64, vp, vp, vp]
        L.orc_occupation_pattern.restype = u64
        L.orc_occupation_pattern.argtypes = [vp, vp, vp]
        L.orc_assemble_laplace.restype = i32
        L.orc_assemble_laplace.argtypes = [vp, vp, vp, vp, C.POINTER(capi.Function), vp]
        L.orc_assemble_stokes.restype = i32
        L.orc_assemble_stokes.argtypes = [vp, vp, vp, vp]
        L.orc_dirichlet.argtypes = [u64, vp, vp, vp, vp, vp, vp]
        L.orc_csr_mv.argtypes = [u64, vp, vp, vp, vp, vp]
        L.orc_set_vertex_ids.restype = i32
        L.orc_set_vertex_ids.argtypes = [vp, vp, u64]
        L.orc_face_orientations.argtypes = [i32, vp, vp, vp]
        L.orc_quad_dof_table.argtypes = [i32, i32, vp]
        _lib = L
    return _lib


class OracleBasis:
    def __init__(self, grid, factory):
        self.grid = grid
        self._desc = grid.desc()
        arr, n = tree_array(factory)
        self._h = lib().orc_basis_create(C.byref(self._desc), arr, n)
        if not self._h:
            raise ValueError("oracle: bad tree")
        if getattr(grid, "vertex_ids", None) is not None:
            self.setVertexIds(grid.vertex_ids)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_basis_destroy(self._h)
            self._h = None

    def dimension(self):
        return int(lib().orc_dimension(self._h))

    def setVertexIds(self, ids):
        """ids of the grid's globalIdSet for the vertices of the local box (lexicographic order);
        None restores the structured grid's own monotone ids"""
        if ids is None:
            rc = lib().orc_set_vertex_ids(self._h, None, 0)
        else:
            a = np.ascontiguousarray(ids, dtype=np.uint64)
            rc = lib().orc_set_vertex_ids(self._h, a.ctypes.data, a.size)
        if rc:
            raise ValueError("oracle: wrong number of vertex ids")

    def size(self, prefix=()):
        p = (C.c_uint64 * max(1, len(prefix)))(*prefix)
        return int(lib().orc_size_prefix(self._h, p, len(prefix)))

    def maxNodeSize(self):
        return int(lib().orc_max_node_size(self._h))

    def maxMultiIndexSize(self):
        return int(lib().orc_max_digits(self._h))

    def minMultiIndexSize(self):
        return int(lib().orc_min_digits(self._h))

    def numElements(self):
        return int(lib().orc_num_elements(self._h))

    def numLeaves(self):
        return int(lib().orc_num_leaves(self._h))

    def numTreeNodes(self):
        return int(lib().orc_num_nodes(self._h))

    def nodeInfo(self, node):
        out = (C.c_int32 * 4)()
        lib().orc_node_info(self._h, node, out)
        return dict(local_offset=out[0], size=out[1], first_leaf=out[2], n_leaves=out[3])

    def indices(self, e0=0, e1=None):
        """(digits [ne,nloc,maxlen] uint64, lens [ne,nloc] uint8)"""
        e1 = self.numElements() if e1 is None else e1
        nl, st = self.maxNodeSize(), self.maxMultiIndexSize()
        dig = np.zeros((e1 - e0, nl, st), dtype=np.uint64)
        lens = np.zeros((e1 - e0, nl), dtype=np.uint8)
        lib().orc_indices(self._h, e0, e1, dig.ctypes.data, lens.ctypes.data, st)
        return dig, lens

    def flatIndices(self, e0=0, e1=None):
        e1 = self.numElements() if e1 is None else e1
        out = np.zeros((e1 - e0, self.maxNodeSize()), dtype=np.uint64)
        lib().orc_indices_flat(self._h, e0, e1, out.ctypes.data)
        return out

    def flatIndicesByEnumeration(self, e0=0, e1=None):
        """flat positions by collecting and sorting every multi-index of the grid (small grids only)"""
        e1 = self.numElements() if e1 is None else e1
        out = np.zeros((e1 - e0, self.maxNodeSize()), dtype=np.uint64)
        lib().orc_indices_flat_by_enumeration(self._h, e0, e1, out.ctypes.data)
        return out

    def interpolate(self, f, coeff=None, mask=None, node=0, nthreads=1, e0=None, e1=None):
        """e0/e1: restrict the element loop to [e0, e1) (slices of grids too large to run whole)"""
        if coeff is None:
            coeff = np.zeros(self.dimension(), dtype=np.float64)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        if e0 is None and e1 is None:
            lib().orc_interpolate(self._h, node, C.byref(f), coeff.ctypes.data,
                                  None if m is None else m.ctypes.data, nthreads)
        else:
            e0 = 0 if e0 is None else e0
            e1 = self.numElements() if e1 is None else e1
            lib().orc_interpolate_range(self._h, node, C.byref(f), coeff.ctypes.data,
                                        None if m is None else m.ctypes.data, e0, e1, nthreads)
        return coeff

    def evaluate(self, coeff, xhat, node=0, values=True, jacobians=False, e0=0, e1=None, nthreads=1):
        e1 = self.numElements() if e1 is None else e1
        xh = np.ascontiguousarray(np.asarray(xhat, dtype=np.float64).reshape(-1, self.grid.dim))
        nq = xh.shape[0]
        nc = self.nodeInfo(node)["n_leaves"]
        coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        v = np.zeros((e1 - e0, nq, nc)) if values else None
        j = np.zeros((e1 - e0, nq, nc, self.grid.dim)) if jacobians else None
        lib().orc_evaluate(self._h, node, coeff.ctypes.data, xh.ctypes.data, nq, e0, e1,
                           None if v is None else v.ctypes.data, None if j is None else j.ctypes.data, nthreads)
        if values and jacobians:
            return v, j
        return v if values else j

    def subEntityDOFs(self, subEntityIndex, subEntityCodim, node=0):
        nl = self.maxNodeSize()
        flags = np.zeros(nl, dtype=np.uint8)
        lst = np.zeros(nl, dtype=np.int32)
        n = lib().orc_subentity_dofs(self._h, node, subEntityIndex, subEntityCodim, flags.ctypes.data, lst.ctypes.data)
        return lst[:n].tolist(), flags.astype(bool)

    def boundaryDOFs(self, node=0, mask=None):
        """(mask, number of callback invocations) of forEachBoundaryDOF"""
        if mask is None:
            mask = np.zeros(self.dimension(), dtype=np.uint8)
        visits = lib().orc_boundary_dofs(self._h, node, mask.ctypes.data)
        return mask, int(visits)

    def interpolateGridFunction(self, source, sourceCoeff, coeff=None, mask=None, node=0, sourceNode=0):
        """interpolate(self, coeff, makeDiscreteGlobalBasisFunction(source, sourceCoeff))"""
        if coeff is None:
            coeff = np.zeros(self.dimension())
        sc = np.ascontiguousarray(sourceCoeff, dtype=np.float64)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = lib().orc_interpolate_dgbf(self._h, node, source._h, sourceNode, sc.ctypes.data, coeff.ctypes.data,
                                        None if m is None else m.ctypes.data)
        if rc:
            raise ValueError(f"oracle: grid-function interpolation failed ({rc})")
        return coeff

    def evaluateGlobal(self, coeff, x, node=0, jacobians=False):
        """DiscreteGlobalBasisFunction::operator()(x) at world coordinates x [n, dim]:
        (values [n, ncomp], jacobians [n, ncomp, dim] or None, element [n])"""
        xs = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1, self.grid.dim))
        n = xs.shape[0]
        nc = self.nodeInfo(node)["n_leaves"]
        coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        v = np.zeros((n, nc))
        j = np.zeros((n, nc, self.grid.dim)) if jacobians else None
        el = np.zeros(n, dtype=np.int64)
        lib().orc_evaluate_global(self._h, node, coeff.ctypes.data, xs.ctypes.data, n, v.ctypes.data,
                                  None if j is None else j.ctypes.data, el.ctypes.data)
        return v, j, el

    # ---- examples/poisson-pq2.cc restated ----
    def occupationPattern(self):
        n = self.dimension()
        row_ptr = np.zeros(n + 1, dtype=np.uint64)
        nnz = lib().orc_occupation_pattern(self._h, row_ptr.ctypes.data, None)
        col = np.zeros(nnz, dtype=np.uint32)
        lib().orc_occupation_pattern(self._h, row_ptr.ctypes.data, col.ctypes.data)
        return row_ptr, col

    def assembleLaplace(self, f=None, pattern=None):
        row_ptr, col = pattern if pattern is not None else self.occupationPattern()
        values = np.zeros(len(col))
        rhs = np.zeros(self.dimension()) if f is not None else None
        rc = lib().orc_assemble_laplace(self._h, row_ptr.ctypes.data, col.ctypes.data, values.ctypes.data,
                                        C.byref(f) if f is not None else None, None if rhs is None else rhs.ctypes.data)
        if rc:
            raise ValueError("oracle: the Laplace assembler needs a single-leaf basis")
        return (row_ptr, col, values), rhs


def oracle_stokes(basis, pattern=None):
    """assembleStokesMatrix of examples/stokes-taylorhood.cc on flat positions: (row_ptr, col, values)"""
    row_ptr, col = pattern if pattern is not None else basis.occupationPattern()
    values = np.zeros(len(col))
    if lib().orc_assemble_stokes(basis._h, row_ptr.ctypes.data, col.ctypes.data, values.ctypes.data):
        raise ValueError("oracle: the Stokes assembler needs a Taylor-Hood tree")
    return row_ptr, col, values


def oracle_dirichlet(csr, dirichlet, x, rhs):
    row_ptr, col, values = csr
    d = np.ascontiguousarray(dirichlet, dtype=np.uint8)
    lib().orc_dirichlet(len(row_ptr) - 1, row_ptr.ctypes.data, col.ctypes.data, values.ctypes.data, d.ctypes.data,
                        x.ctypes.data, rhs.ctypes.data)


def oracle_mv(csr, x):
    row_ptr, col, values = csr
    y = np.zeros(len(row_ptr) - 1)
    lib().orc_csr_mv(len(row_ptr) - 1, row_ptr.ctypes.data, col.ctypes.data, values.ctypes.data, x.ctypes.data, y.ctypes.data)
    return y


def faceOrientations(dim, ids):
    """(edge bits, facet orientation indices) of one cube with the given vertex ids"""
    a = np.ascontiguousarray(ids, dtype=np.uint64)
    assert a.size == 1 << dim
    edges = np.zeros(12, dtype=np.int32)
    facets = np.zeros(6, dtype=np.int32)
    lib().orc_face_orientations(dim, a.ctypes.data, edges.ctypes.data, facets.ctypes.data)
    ne = {1: 0, 2: 4, 3: 12}[dim]
    return edges[:ne].tolist(), (facets.tolist() if dim == 3 else [])


def quadDofTable(order, j):
    out = np.zeros((order - 1) ** 2, dtype=np.uint32)
    lib().orc_quad_dof_table(order, j, out.ctypes.data)
    return out
