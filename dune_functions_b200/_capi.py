"""ctypes binding of the C ABI declared in include/dfx.h.

The shared library is the product; there is no Python or CPU fallback.  If
`lib/libdfx.so` is missing the import of this module raises, and every compute
entry point of the library itself fails with DFX_ERR_CUDA when no sm_100 device
is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DFX_LIB selects another build of the same library (tuning variants under lib/)
LIB_PATH = os.environ.get("DFX_LIB") or os.path.join(HERE, "lib", "libdfx.so")

DFX_MAX_DIM = 3
DFX_MAX_DIGITS = 6
DFX_MAX_LEAVES = 16
DFX_MAX_FN_PARAMS = 16

# dfx_status
DFX_OK, DFX_ERR_INVALID, DFX_ERR_RANGE, DFX_ERR_NOT_IMPLEMENTED, DFX_ERR_CUDA, DFX_ERR_NCCL = range(6)
# dfx_node_kind
NODE_LAGRANGE, NODE_LAGRANGE_DG, NODE_POWER, NODE_COMPOSITE, NODE_TAYLOR_HOOD = range(5)
# dfx_ims
IMS_NONE, IMS_FLAT_LEXICOGRAPHIC, IMS_FLAT_INTERLEAVED, IMS_BLOCKED_LEXICOGRAPHIC, IMS_BLOCKED_INTERLEAVED = range(5)
# dfx_cd_kind
CD_UNKNOWN, CD_VALUE, CD_TUPLE, CD_ARRAY, CD_VECTOR, CD_UNIFORM_ARRAY, CD_UNIFORM_VECTOR = range(7)
# dfx_fn_id
FN_CONSTANT, FN_GAUSSIAN, FN_IDENTITY, FN_COORD, FN_QUADRATIC, FN_SINPROD, FN_GAUSSIAN_VEC = range(7)


class GridDesc(C.Structure):
    _fields_ = [("dim", C.c_int32),
                ("n_global", C.c_int32 * DFX_MAX_DIM),
                ("local_begin", C.c_int32 * DFX_MAX_DIM),
                ("local_n", C.c_int32 * DFX_MAX_DIM),
                ("lower", C.c_double * DFX_MAX_DIM),
                ("upper", C.c_double * DFX_MAX_DIM)]


class TreeNode(C.Structure):
    _fields_ = [("kind", C.c_int32), ("order", C.c_int32),
                ("n_children", C.c_int32), ("ims", C.c_int32)]


class ContainerNode(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_stored", C.c_int32), ("size", C.c_uint64)]


class Function(C.Structure):
    _fields_ = [("id", C.c_int32), ("n_comp", C.c_int32),
                ("p", C.c_double * DFX_MAX_FN_PARAMS)]


# every symbol include/dfx.h declares: name -> (restype, argtypes)
_vp, _u64, _i32 = C.c_void_p, C.c_uint64, C.c_int32
_pu64, _pi32, _pu8 = C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
SIGNATURES = {
    "dfx_abi_version": (C.c_int, []),
    "dfx_last_error": (C.c_char_p, []),
    "dfx_kernel_launch_count": (_u64, []),
    "dfx_basis_create": (C.c_int, [C.POINTER(GridDesc), C.POINTER(TreeNode), _i32, _i32, C.POINTER(_vp)]),
    "dfx_basis_destroy": (None, [_vp]),
    "dfx_basis_update": (C.c_int, [_vp, C.POINTER(GridDesc)]),
    "dfx_basis_set_vertex_ids": (C.c_int, [_vp, _vp, _u64, _vp]),
    "dfx_basis_set_vertex_ids_host": (C.c_int, [_vp, _vp, _u64]),
    "dfx_basis_has_permuted_face_dofs": (_i32, [_vp]),
    "dfx_basis_dimension": (_u64, [_vp]),
    "dfx_basis_size": (_u64, [_vp, _pu64, _i32]),
    "dfx_basis_container_descriptor": (C.c_int, [_vp, C.POINTER(ContainerNode), _i32, _pi32, C.c_char_p, _i32]),
    "dfx_basis_max_node_size": (_i32, [_vp]),
    "dfx_basis_max_multi_index_size": (_i32, [_vp]),
    "dfx_basis_min_multi_index_size": (_i32, [_vp]),
    "dfx_basis_num_elements": (_u64, [_vp]),
    "dfx_basis_num_leaves": (_i32, [_vp]),
    "dfx_basis_num_tree_nodes": (_i32, [_vp]),
    "dfx_basis_node_info": (C.c_int, [_vp, _i32, _pi32, _pi32, _pi32, _pi32]),
    "dfx_basis_node_child": (_i32, [_vp, _i32, _i32]),
    "dfx_basis_local_index_sizes": (C.c_int, [_vp, _pu8]),
    "dfx_indices_multi": (C.c_int, [_vp, _u64, _u64, _vp, _i32, _vp]),
    "dfx_indices_multi_u64": (C.c_int, [_vp, _u64, _u64, _vp, _i32, _vp]),
    "dfx_indices_flat": (C.c_int, [_vp, _u64, _u64, _vp, _vp]),
    "dfx_indices_flat_u64": (C.c_int, [_vp, _u64, _u64, _vp, _vp]),
    "dfx_interpolate": (C.c_int, [_vp, _i32, C.POINTER(Function), _vp, _vp, _vp]),
    "dfx_interpolate_separable": (C.c_int, [_vp, _i32, C.POINTER(Function), _vp, _vp, _vp]),
    "dfx_interpolate_batch": (C.c_int, [_vp, _i32, _pi32, C.POINTER(Function), _vp, _vp, _vp]),
    "dfx_interpolation_points": (C.c_int, [_vp, _vp, _vp, _vp]),
    "dfx_interpolate_from_values": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "dfx_interpolate_dgbf": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _u64, _vp]),
    "dfx_interpolate_dgbf_host": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "dfx_evaluate": (C.c_int, [_vp, _i32, _vp, C.POINTER(C.c_double), _i32, _u64, _u64, _vp, _vp, _vp]),
    "dfx_evaluate_global": (C.c_int, [_vp, _i32, _vp, _vp, _u64, _vp, _vp, _vp, _vp]),
    "dfx_evaluate_global_host": (C.c_int, [_vp, _i32, _vp, _vp, _u64, _vp, _vp, _vp]),
    "dfx_evaluate_path": (C.c_int, [_vp, _i32, C.POINTER(C.c_double), _i32, _i32]),
    "dfx_set_kernel_path": (None, [_i32, _i32]),
    "dfx_interpolate_host": (C.c_int, [_vp, _i32, C.POINTER(Function), _vp, _vp]),
    "dfx_evaluate_host": (C.c_int, [_vp, _i32, _vp, C.POINTER(C.c_double), _i32, _u64, _u64, _vp, _vp]),
    "dfx_indices_flat_host": (C.c_int, [_vp, _u64, _u64, _vp]),
    "dfx_indices_multi_host": (C.c_int, [_vp, _u64, _u64, _vp, _i32]),
    "dfx_interpolation_points_host": (C.c_int, [_vp, _vp, _vp]),
    "dfx_subentity_dofs": (C.c_int, [_vp, _i32, _i32, _i32, _pu8, _pi32, _pi32]),
    "dfx_boundary_dofs": (C.c_int, [_vp, _i32, _vp, _vp]),
    "dfx_boundary_dofs_host": (C.c_int, [_vp, _i32, _vp]),
    "dfx_matrix_pattern": (C.c_int, [_vp, _vp, _vp, _pu64, _vp]),
    "dfx_assemble_laplace": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(Function), _vp, _vp]),
    "dfx_matrix_pattern_host": (C.c_int, [_vp, _vp, _vp, _pu64]),
    "dfx_assemble_laplace_host": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(Function), _vp]),
    "dfx_assemble_stokes": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "dfx_assemble_stokes_host": (C.c_int, [_vp, _vp, _vp, _vp]),
    "dfx_apply_laplace": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "dfx_csr_mv": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "dfx_apply_dirichlet": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dfx_halo_ranges": (C.c_int, [_vp, _i32, _pu64, _pu64, _i32, _pi32]),
    "dfx_halo_size": (_u64, [_vp]),
    "dfx_halo_pack": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "dfx_halo_unpack": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "dfx_halo_link_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "dfx_halo_link_destroy": (None, [_vp]),
    "dfx_halo_link_export": (C.c_int, [_vp, _vp]),
    "dfx_halo_link_block": (_vp, [_vp]),
    "dfx_halo_link_attach": (C.c_int, [_vp, _i32, _vp, _vp]),
    "dfx_halo_link_error": (_u64, [_vp]),
    "dfx_halo_link_reset_error": (C.c_int, [_vp]),
    "dfx_halo_link_set_timeout": (C.c_int, [_vp, C.c_double]),
    "dfx_halo_push": (C.c_int, [_vp, _vp, _vp, _u64, _vp]),
    "dfx_halo_pull": (C.c_int, [_vp, _vp, _vp, _u64, _vp]),
    "dfx_owned_ranges": (C.c_int, [_vp, _pu64, _pu64, _pu64, _i32, _pi32]),
}

OP_EVALUATE, OP_INTERPOLATE, OP_INDICES, OP_HOST_BAND_KIB, OP_SMEM_GRADIENT, OP_TUNE_ROWS, OP_TUNE_SWIZZLE, OP_TUNE_STREAM_HINT, OP_TUNE_SMEM_TMA, OP_TUNE_EVAL_BLOCK, OP_TUNE_SMEM_PAIR, OP_TUNE_SMEM_BLOCKS, OP_TUNE_EVAL_Q3, OP_TUNE_SPLIT_MAP, OP_TUNE_PAIR_PREFETCH = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14

_lib = None


def lib():
    """Load lib/libdfx.so (built by __graft_entry__.build()); loud failure if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class DfxError(RuntimeError):
    """Base of the errors raised for non-zero dfx_status codes."""


class RangeError(DfxError):
    """Dune::RangeError equivalent (bad order, size mismatch, ...)."""


class NotImplementedDfx(DfxError):
    """Dune::NotImplemented equivalent."""


class CudaError(DfxError):
    pass


def check(status):
    if status == DFX_OK:
        return
    msg = lib().dfx_last_error().decode()
    exc = {DFX_ERR_RANGE: RangeError, DFX_ERR_NOT_IMPLEMENTED: NotImplementedDfx,
           DFX_ERR_CUDA: CudaError}.get(status, DfxError)
    raise exc(f"dfx status {status}: {msg}")
