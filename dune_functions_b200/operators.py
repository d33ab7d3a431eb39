"""The element-loop callers next to the hot path: what
examples/poisson-pq2.cc builds on a global basis — occupation pattern, Laplace stiffness matrix
and load vector, symmetric Dirichlet treatment, matrix-vector products — on the device through
the C ABI (dfx_matrix_pattern, dfx_assemble_laplace, dfx_csr_mv, dfx_apply_dirichlet).

Rows and columns are flat container positions.  row_ptr is an int64 tensor (the ABI's uint64
bits), col_idx an int32 tensor (uint32 bits), values float64 — all CUDA tensors."""
import ctypes as C
from dataclasses import dataclass

from . import _capi as capi
from .basis import _ptr, _stream, _root_and_node


@dataclass
class CSRMatrix:
    """BCRSMatrix<double> stand-in: compressed rows, ascending column indices per row"""
    basis: object
    row_ptr: object
    col_idx: object
    values: object

    @property
    def nnz(self):
        return int(self.col_idx.numel())

    def mv(self, x, y=None):
        """y = A x   (BCRSMatrix::mv)"""
        import torch
        if y is None:
            y = torch.empty_like(x)
        b = self.basis
        capi.check(b._lib.dfx_csr_mv(b._h, _ptr(self.row_ptr), _ptr(self.col_idx), _ptr(self.values), _ptr(x), _ptr(y), 0,
                                     _stream()))
        return y

    def mmv(self, x, y):
        """y -= A x   (BCRSMatrix::mmv)"""
        b = self.basis
        capi.check(b._lib.dfx_csr_mv(b._h, _ptr(self.row_ptr), _ptr(self.col_idx), _ptr(self.values), _ptr(x), _ptr(y), 1,
                                     _stream()))
        return y

    def to_scipy(self):
        import scipy.sparse as sp
        n = self.row_ptr.numel() - 1
        return sp.csr_matrix((self.values.cpu().numpy(), self.col_idx.cpu().numpy().view("uint32"),
                              self.row_ptr.cpu().numpy()), shape=(n, n))


def getOccupationPattern(basis):
    """getOccupationPattern(feBasis, nb) + nb.exportIdx(matrix), poisson-pq2.cc:137-174, 199-203.
    Returns (row_ptr, col_idx)."""
    import torch
    root, node = _root_and_node(basis)
    if node != 0:
        raise capi.NotImplementedDfx("occupation pattern of a sub-space basis")
    dev = f"cuda:{root.device}"
    n = root.dimension()
    row_ptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    nnz = C.c_uint64()
    capi.check(root._lib.dfx_matrix_pattern(root._h, _ptr(row_ptr), None, C.byref(nnz), _stream()))
    col_idx = torch.empty(nnz.value, dtype=torch.int32, device=dev)
    capi.check(root._lib.dfx_matrix_pattern(root._h, _ptr(row_ptr), _ptr(col_idx), None, _stream()))
    return row_ptr, col_idx


def assembleLaplaceMatrix(basis, volumeTerm=None, pattern=None):
    """assembleLaplaceMatrix(feBasis, matrix, rhs, volumeTerm), poisson-pq2.cc:176-255.
    volumeTerm: a registry function (dfx_function) or None (matrix only).  Returns (CSRMatrix, rhs)."""
    import torch
    row_ptr, col_idx = pattern if pattern is not None else getOccupationPattern(basis)
    dev = row_ptr.device
    values = torch.empty(col_idx.numel(), dtype=torch.float64, device=dev)
    rhs = torch.empty(basis.dimension(), dtype=torch.float64, device=dev) if volumeTerm is not None else None
    capi.check(basis._lib.dfx_assemble_laplace(basis._h, _ptr(row_ptr), _ptr(col_idx), _ptr(values),
                                               C.byref(volumeTerm) if volumeTerm is not None else None, _ptr(rhs), _stream()))
    return CSRMatrix(basis, row_ptr, col_idx, values), rhs


def assembleStokesMatrix(basis, pattern=None):
    """assembleStokesMatrix(basis, matrix), examples/stokes-taylorhood.cc:252-297, for a Taylor-Hood
    basis; the nested matrix of the example flattened to container positions.  Returns a CSRMatrix."""
    import torch
    row_ptr, col_idx = pattern if pattern is not None else getOccupationPattern(basis)
    values = torch.empty(col_idx.numel(), dtype=torch.float64, device=row_ptr.device)
    capi.check(basis._lib.dfx_assemble_stokes(basis._h, _ptr(row_ptr), _ptr(col_idx), _ptr(values), _stream()))
    return CSRMatrix(basis, row_ptr, col_idx, values)


def applyDirichlet(matrix, dirichletNodes, x, rhs):
    """poisson-pq2.cc:356-386: rhs -= A x, Dirichlet rows/columns -> identity, rhs = x on them"""
    import torch
    b = matrix.basis
    m = dirichletNodes.view(torch.uint8) if dirichletNodes.dtype == torch.bool else dirichletNodes
    capi.check(b._lib.dfx_apply_dirichlet(b._h, _ptr(matrix.row_ptr), _ptr(matrix.col_idx), _ptr(matrix.values), _ptr(m),
                                          _ptr(x), _ptr(rhs), _stream()))
    return matrix, rhs


class LaplaceOperator:
    """The Laplace stiffness matrix of assembleLaplaceMatrix as a matrix-free operator
    (dfx_apply_laplace): mv(x) = A x without ever storing A.  With dirichletNodes the operator is the
    symmetrically modified matrix of poisson-pq2.cc:356-386 (identity on those rows and columns)."""

    def __init__(self, basis, dirichletNodes=None):
        import torch
        self.basis = basis
        self.mask = None
        if dirichletNodes is not None:
            self.mask = dirichletNodes.view(torch.uint8) if dirichletNodes.dtype == torch.bool else dirichletNodes
        self.scratch = torch.empty(basis.maxNodeSize() * basis.numElements() + basis.dimension(), dtype=torch.float64,
                                   device=f"cuda:{basis.device}")

    def mv(self, x, y=None):
        import torch
        if y is None:
            y = torch.empty_like(x)
        b = self.basis
        capi.check(b._lib.dfx_apply_laplace(b._h, _ptr(x), _ptr(y), _ptr(self.mask), _ptr(self.scratch), _stream()))
        return y
