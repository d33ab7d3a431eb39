/* dfx.h — C ABI of the B200-native element-loop hot path of dune-functions.
 *
 * This is the drop-in boundary.  The reference (dune-functions, header-only C++)
 * has no FFI of its own; its "operator API" for this path is the set of C++
 * concepts and free functions cited next to each entry point below
 * (paths relative to the reference checkout).  A C++ façade that keeps the
 * reference's names on top of this ABI lives in include/dfx/dune_functions.hh;
 * INTEGRATION.md shows how a maintainer binds it.
 *
 * Conventions
 *  - every function returns a dfx_status (0 = ok); dfx_last_error() gives the
 *    text of the last failure on the calling thread.  No exception crosses
 *    the ABI.
 *  - pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *  - `stream` is a cudaStream_t passed as void*; all device work is
 *    stream-ordered, nothing synchronises unless documented.
 *  - buffers are caller-owned; the library never reallocates them.
 *  - a handle is thread-COMPATIBLE, not thread-safe: it owns device caches (shape tables, scratch,
 *    the uploaded descriptor) that calls rewrite stream-ordered on the stream they are given, so use
 *    one stream and one thread at a time per handle (LocalView / LocalFunction are per-thread
 *    objects in the reference too, defaultlocalview.hh:95-101).  Several handles — one per thread or
 *    per stream — on the same grid are independent.
 *  - there is NO CPU fallback: every compute entry point fails with
 *    DFX_ERR_CUDA when no sm_100 device is usable.
 *
 * Coefficient-vector layout ("flat container position"): the basis' index
 * tree is laid out contiguously in lexicographic order of the multi-indices,
 * i.e. exactly how nested ISTL containers resolve `vector[multiIndex]`
 * (reference: dune/functions/backends/istlvectorbackend.hh:267-277,
 * dune/functions/common/indexaccess.hh:376-403).  Flat bases: position ==
 * the single index digit.  Blocked bases: e.g. BL(BI) Taylor-Hood
 * (0,i,c) -> 3*i+c, (1,j) -> 3*n2+j.
 */
#ifndef DFX_H
#define DFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFX_ABI_VERSION 1
#define DFX_MAX_DIM 3
#define DFX_MAX_DIGITS 6      /* longest multi-index the ABI represents          */
#define DFX_MAX_LEAVES 16     /* leaf nodes of one basis tree                    */
#define DFX_MAX_FN_PARAMS 16

typedef enum {
  DFX_OK = 0,
  DFX_ERR_INVALID = 1,        /* malformed descriptor / argument                 */
  DFX_ERR_RANGE = 2,          /* Dune::RangeError equivalents (bad order, ...)    */
  DFX_ERR_NOT_IMPLEMENTED = 3,/* Dune::NotImplemented equivalents                 */
  DFX_ERR_CUDA = 4,           /* CUDA runtime failure / no device                 */
  DFX_ERR_NCCL = 5
} dfx_status;

/* ---- grid: an unrefined equidistant YaspGrid-style box --------------------
 * n_global/lower/upper describe the whole grid; [local_begin, local_begin +
 * local_n) is the box of elements this handle (one GPU rank) works on.  The
 * basis is numbered over the LOCAL box (like a basis built on that rank's grid
 * view) while node coordinates and the last-writer rule of interpolate() are
 * those of the GLOBAL grid, so a slab-partitioned run reproduces the serial
 * one bit for bit.  Unpartitioned: local_begin = 0, local_n = n_global.
 * Replaces: Dune::YaspGrid<dim>(L, N) + leafGridView() as consumed by
 * lagrangebasis.hh:783-797 (MCMGMapper over the grid view's index set).     */
typedef struct {
  int32_t dim;                      /* 1..3                                    */
  int32_t n_global[DFX_MAX_DIM];    /* elements per direction, whole grid      */
  int32_t local_begin[DFX_MAX_DIM];
  int32_t local_n[DFX_MAX_DIM];
  double  lower[DFX_MAX_DIM];       /* lower-left corner of the whole grid     */
  double  upper[DFX_MAX_DIM];       /* upper-right corner of the whole grid    */
} dfx_grid_desc;

/* ---- basis tree: preorder array of nodes ----------------------------------
 * Replaces the compile-time pre-basis types
 *   LagrangePreBasis<GV,k,R>      lagrangebasis.hh:740-742
 *   LagrangeDGPreBasis<GV,k,R>    lagrangedgbasis.hh:48-50
 *   PowerPreBasis<IMS,SPB,C>      powerbasis.hh:51-53  (one child subtree follows)
 *   CompositePreBasis<IMS,SPB...> compositebasis.hh:55-56 (n_children subtrees follow)
 *   TaylorHoodPreBasis<GV>        taylorhoodbasis.hh:59-281 (no children in the array)
 */
typedef enum {
  DFX_NODE_LAGRANGE = 0,
  DFX_NODE_LAGRANGE_DG = 1,
  DFX_NODE_POWER = 2,
  DFX_NODE_COMPOSITE = 3,
  DFX_NODE_TAYLOR_HOOD = 4
} dfx_node_kind;

/* index-merging strategies, functionspacebases/basistags.hh:52-184 */
typedef enum {
  DFX_IMS_NONE = 0,
  DFX_IMS_FLAT_LEXICOGRAPHIC = 1,
  DFX_IMS_FLAT_INTERLEAVED = 2,
  DFX_IMS_BLOCKED_LEXICOGRAPHIC = 3,
  DFX_IMS_BLOCKED_INTERLEAVED = 4
} dfx_ims;

typedef struct {
  int32_t kind;        /* dfx_node_kind                                         */
  int32_t order;       /* polynomial order k (leaves)                           */
  int32_t n_children;  /* power: exponent C; composite: number of children      */
  int32_t ims;         /* dfx_ims (inner nodes)                                 */
} dfx_tree_node;

/* ---- functions f for interpolate() ----------------------------------------
 * A host callable cannot run on the device, so f crosses the ABI either as
 * (1) an entry of the built-in registry of __device__ functors below, or
 * (2) nodal values produced by the caller (dfx_interpolation_points +
 *     dfx_interpolate_from_values), or
 * (3) for C++ callers compiling with nvcc: the header template
 *     dfx::interpolate(basis, coeff, DeviceFunctor) (include/dfx/device_functor.cuh).
 * Replaces the callable argument of interpolate(), interpolate.hh:204-211.   */
typedef enum {
  DFX_FN_CONSTANT = 0,     /* y_c = p[c]                      (n_comp values)        */
  DFX_FN_GAUSSIAN = 1,     /* y   = exp(-p[0] * |x|^2)         examples/interpolation.cc:49-52 */
  DFX_FN_IDENTITY = 2,     /* y_c = x_c                        examples/interpolation.cc:93-95 */
  DFX_FN_COORD = 3,        /* y   = x[(int)p[0]]               gridviewfunctionspacebasistest.cc */
  DFX_FN_QUADRATIC = 4,    /* y = p0 + sum p[1+j] x_j + sum_{i<=j} p[4+idx(i,j)] x_i x_j,
                              idx order (00,01,02,11,12,22)   cf. discreteglobalbasisfunctionderivativetest.cc:89-91 */
  DFX_FN_SINPROD = 5,      /* y = prod_j sin(p[0] * x_j)                              */
  DFX_FN_GAUSSIAN_VEC = 6  /* y_c = x_c * exp(-p[0]*|x|^2)     (vector valued, smooth) */
} dfx_fn_id;

typedef struct {
  int32_t id;              /* dfx_fn_id                                          */
  int32_t n_comp;          /* number of range entries; 1 = scalar (broadcast)    */
  double  p[DFX_MAX_FN_PARAMS];
} dfx_function;

typedef struct dfx_basis dfx_basis;   /* opaque, caller-owned, explicit destroy */

/* ---- life cycle ------------------------------------------------------------*/
int         dfx_abi_version(void);
const char* dfx_last_error(void);
/* number of CUDA kernels this library has launched in the calling process    */
uint64_t    dfx_kernel_launch_count(void);

/* makeBasis(gridView, preBasisFactory): defaultglobalbasis.hh:205-209, ctor :89-95.
 * `device` is the CUDA ordinal the handle is bound to.
 * Deliberate restrictions (DFX_ERR_RANGE): continuous Lagrange order <= 17 in EVERY dimension (the reference
 * restricts 3-d only, lagrangebasis.hh:795; the kernels' tables are sized for 17), LagrangeDG order <= 12,
 * at most DFX_MAX_LEAVES leaves and DFX_MAX_DIGITS multi-index digits.                                  */
int  dfx_basis_create(const dfx_grid_desc* grid_host, const dfx_tree_node* tree_host,
                      int32_t n_nodes, int32_t device, dfx_basis** out);
void dfx_basis_destroy(dfx_basis* b);
/* basis.update(gridView): defaultglobalbasis.hh:137-141 (preBasis_.update(gv); initializeIndices()).
 * The same tree on another grid box; the handle stays valid, everything that belonged to the old
 * grid (index tables, coefficient vectors, vertex ids set with dfx_basis_set_vertex_ids) does not.
 * On failure the handle keeps its old state.  Not stream-ordered: work in flight on the handle must
 * have finished (the reference has the same rule, :134-136).                                       */
int  dfx_basis_update(dfx_basis* b, const dfx_grid_desc* grid_host);

/* Vertex ids for the face-DOF permutation of continuous Lagrange bases of order >= 3.
 * Replaces: the `gridView.grid().globalIdSet()` argument of Experimental::LagrangeFaceDOFPermutation
 * (lagrangebasis.hh:787, :585-604) as read by computeFaceOrientations (:616-634).  `ids` holds one
 * pairwise different id per vertex of the LOCAL box in lexicographic order (x fastest),
 * n = prod (local_n[j] + 1); the library keeps its own copy.  Without ids (the default, or
 * ids = NULL) the grid's own ids are used: they are monotone in lexicographic position, every edge
 * and facet orientation is the identity and permuteFaceDOF (:652-675) is a no-op.  With ids — a grid
 * manager whose ids are not ordered like the lattice (UGGrid / ALUGrid cube meshes, refined grids;
 * lagrangebasistest.cc:34-70) — edge DOFs are flipped (:232-235, :660-668) and the DOFs of
 * quadrilateral facets run through the table of :518-559 selected by the facet's orientation index
 * (:279-331) in bind/index, interpolate and evaluate.  Orders <= 2 and 1-d grids are unaffected.
 * Set ids before computing anything that depends on the numbering; in a slab-partitioned run every
 * rank passes the ids of its own box (same id for the same vertex on both sides of an interface).
 * dfx_assemble_laplace and dfx_apply_laplace return DFX_ERR_NOT_IMPLEMENTED on a basis whose face
 * DOFs are permuted.  The _host form is host bookkeeping (it works without a device, like
 * dfx_basis_create; the table is uploaded by the next device call), the device form copies
 * stream-ordered.                                                                                  */
int     dfx_basis_set_vertex_ids(dfx_basis* b, const uint64_t* ids, uint64_t n, void* stream);
int     dfx_basis_set_vertex_ids_host(dfx_basis* b, const uint64_t* ids_host, uint64_t n);
/* 1 iff ids are set AND some continuous leaf has order >= 3 on a 2-d/3-d grid.  The orientation of every edge and
 * quadrilateral facet of the local box is then computed ONCE per set of ids into a one-byte-per-entity table on the
 * device (at the first index table / interpolate / evaluate call after the ids were set) and the kernels permute the
 * in-entity index through it: index table and evaluation gather on the walking kernel, a single continuous leaf of
 * order 3 - 5 interpolated by the run sweep, order 3 at 4^3 points evaluated in registers; everything else (masks,
 * several leaves per call, other point sets) takes the general kernels or the gather + companion-basis route. */
int32_t dfx_basis_has_permuted_face_dofs(const dfx_basis* b);

/* ---- GlobalBasis concept: functionspacebases/concepts.hh:253-273 -----------*/
uint64_t dfx_basis_dimension(const dfx_basis* b);                 /* defaultglobalbasis.hh:144-147 */
uint64_t dfx_basis_size(const dfx_basis* b, const uint64_t* prefix_host, int32_t len); /* :150-160 */
/* basis.containerDescriptor(): defaultglobalbasis.hh:180-186 -> PreBasis::containerDescriptor
 * (leafprebasismixin.hh:66-69, dynamicpowerbasis.hh:373-387, compositebasis.hh:274-289,
 * taylorhoodbasis.hh:204-216) with the descriptor types of containerdescriptors.hh:91-209 as a
 * run-time tree: preorder array of nodes; a node is followed by its n_stored children — all `size`
 * of them for TUPLE / ARRAY / VECTOR, the single prototype for UNIFORM_ARRAY / UNIFORM_VECTOR, none
 * for VALUE / UNKNOWN.  FlatVector = UNIFORM_VECTOR of VALUE, FlatArray<n> = UNIFORM_ARRAY of VALUE.
 * Host-only bookkeeping (no device needed).  *n_nodes receives the number of nodes the descriptor has;
 * nothing is written beyond `capacity`.  type_out (optional, type_capacity bytes) receives the
 * spelled-out reference type, e.g. "Tuple<UniformVector<UniformArray<Value,3>>,UniformVector<Value>>". */
typedef enum {
  DFX_CD_UNKNOWN = 0, DFX_CD_VALUE = 1, DFX_CD_TUPLE = 2, DFX_CD_ARRAY = 3, DFX_CD_VECTOR = 4,
  DFX_CD_UNIFORM_ARRAY = 5, DFX_CD_UNIFORM_VECTOR = 6
} dfx_cd_kind;
typedef struct {
  int32_t  kind;       /* dfx_cd_kind                                           */
  int32_t  n_stored;   /* children stored behind this node in the array         */
  uint64_t size;       /* number of children the container has at this level    */
} dfx_container_node;
int dfx_basis_container_descriptor(const dfx_basis* b, dfx_container_node* out_host, int32_t capacity,
                                   int32_t* n_nodes, char* type_out, int32_t type_capacity);
int32_t  dfx_basis_max_node_size(const dfx_basis* b);             /* LocalView::maxSize, defaultlocalview.hh:147-150 */
int32_t  dfx_basis_max_multi_index_size(const dfx_basis* b);      /* PreBasis::maxMultiIndexSize */
int32_t  dfx_basis_min_multi_index_size(const dfx_basis* b);
uint64_t dfx_basis_num_elements(const dfx_basis* b);              /* elements(gridView) of the local box */
int32_t  dfx_basis_num_leaves(const dfx_basis* b);
int32_t  dfx_basis_num_tree_nodes(const dfx_basis* b);            /* nodes of the expanded local ansatz tree */
/* node tree (nodes.hh:130-209): for expanded-tree node `node` (preorder
 * "treeIndex"), its first local index, its size and its leaf range.           */
int dfx_basis_node_info(const dfx_basis* b, int32_t node, int32_t* local_offset,
                        int32_t* size, int32_t* first_leaf, int32_t* n_leaves);
/* child `child` of expanded-tree node `node` -> node id (or -1)                */
int32_t dfx_basis_node_child(const dfx_basis* b, int32_t node, int32_t child);
/* multi-index length of local index i (same for every element)                */
int dfx_basis_local_index_sizes(const dfx_basis* b, uint8_t* sizes_host /*[max_node_size]*/);

/* ---- LocalView::bind + index, batched over elements ------------------------
 * Replaces DefaultLocalView::bind (defaultlocalview.hh:95-101) -> PreBasis::indices
 * (lagrangebasis.hh:843-873, lagrangedgbasis.hh:120-130, dynamicpowerbasis.hh:263-371,
 * compositebasis.hh:293-338, taylorhoodbasis.hh:229-251) for elements
 * [elem_begin, elem_end) of the local box in grid-view iteration order
 * (x fastest).  out_digits[(e-elem_begin)*n_loc*stride + i*stride + p] is digit
 * p of localView.index(i); digits past the multi-index length are 0.          */
int dfx_indices_multi(const dfx_basis* b, uint64_t elem_begin, uint64_t elem_end,
                      uint32_t* out_digits, int32_t stride, void* stream);
int dfx_indices_multi_u64(const dfx_basis* b, uint64_t elem_begin, uint64_t elem_end,
                          uint64_t* out_digits, int32_t stride, void* stream);
/* flat container position of every local index: out[(e-elem_begin)*n_loc + i]  */
int dfx_indices_flat(const dfx_basis* b, uint64_t elem_begin, uint64_t elem_end,
                     uint32_t* out, void* stream);
int dfx_indices_flat_u64(const dfx_basis* b, uint64_t elem_begin, uint64_t elem_end,
                         uint64_t* out, void* stream);

/* ---- interpolate(basis, coeff, f [, bitVector]) ----------------------------
 * Replaces interpolate.hh:204-302 (+ Imp::interpolateLocal :151-173,
 * AnalyticGridViewFunction analyticgridviewfunction.hh:77-109,
 * HierarchicNodeToRangeMap hierarchicnodetorangemap.hh:33-48).
 * `subtree_node` = expanded-tree node whose leaves are interpolated (0 = whole
 * basis; otherwise what subspaceBasis(basis, path) selects, subspacebasis.hh:147-157).
 * Leaf number l (within the subtree) takes range entry l of f, a scalar f is
 * broadcast.  coeff has dimension() doubles; entries of other leaves and
 * entries whose mask byte is 0 are left untouched (mask may be NULL = all).   */
int dfx_interpolate(const dfx_basis* b, int32_t subtree_node, const dfx_function* f_host,
                    double* coeff, const uint8_t* mask, void* stream);
/* The same call with an algorithmic shortcut the caller opts into PER CALL: registry functions whose
 * range entries are products of per-direction factors (exp(-a|x|^2) = prod_j exp(-a x_j^2), prod_j
 * sin(w x_j), x_j, constants, x -> x, x -> x exp(-a|x|^2)) are tabulated per direction — O(N)
 * transcendental evaluations instead of one per node — and every coefficient is a product of table
 * entries.  dfx_interpolate evaluates f at every node like Imp::interpolateLocal (interpolate.hh:151-173);
 * this variant is NOT bit-identical to it (<= 3 ulp for the Gaussians, identical for polynomials and
 * constants) and is not what an arbitrary callable can use, so it is never picked implicitly.  A
 * function without such a form is interpolated exactly like dfx_interpolate.                        */
int dfx_interpolate_separable(const dfx_basis* b, int32_t subtree_node, const dfx_function* f_host,
                              double* coeff, const uint8_t* mask, void* stream);
/* Several interpolate calls on the same vector as one call: part p interpolates fns_host[p] into the leaves
 * below subtree_nodes_host[p], in order, exactly like n_parts dfx_interpolate calls (the reference's
 * examples do this by hand: interpolate(subspaceBasis(basis, _0), x, f_velocity); interpolate(subspaceBasis(
 * basis, _1), x, f_pressure), examples/interpolation.cc:89-96).  Consecutive parts on disjoint subtrees whose
 * kernels would each be shorter than a launch gap share ONE launch (Taylor-Hood velocity + pressure).      */
int dfx_interpolate_batch(const dfx_basis* b, int32_t n_parts, const int32_t* subtree_nodes_host,
                          const dfx_function* fns_host, double* coeff, const uint8_t* mask, void* stream);
/* form (2): positions of the interpolation nodes and range entry of every
 * coefficient, in flat container order; then a masked copy of caller values. */
int dfx_interpolation_points(const dfx_basis* b, double* out_xyz /*[dimension][dim]*/,
                             int32_t* out_leaf /*[dimension]*/, void* stream);
int dfx_interpolate_from_values(const dfx_basis* b, int32_t subtree_node, const double* values,
                                double* coeff, const uint8_t* mask, void* stream);

/* interpolate(targetBasis, coeff, gridFunction) where the grid function is a
 * DiscreteGlobalBasisFunction over `source` (any basis on the SAME grid box): makeGridViewFunction
 * hands grid functions through (gridfunctions/gridviewfunction.hh:68-75), so interpolateLocal
 * (interpolate.hh:151-173) evaluates localFunction(f) at the target's local interpolation nodes,
 * element by element — the first variant of checkInterpolationConsistency
 * (gridfunctions/test/discreteglobalbasisfunctiontest.cc:49-77).  Leaf l below target_node takes
 * range entry l of f (= leaf l below source_node); a one-leaf source is broadcast.
 * `scratch` (device, may be NULL = allocated per call): scratch_doubles doubles; the element
 * loop runs in bands that fit it (at least one z-layer: layer elements * max target n_loc *
 * source leaves).                                                                       */
int dfx_interpolate_dgbf(const dfx_basis* target, int32_t target_node, const dfx_basis* source, int32_t source_node,
                         const double* source_coeff, double* target_coeff, const uint8_t* mask,
                         double* scratch, uint64_t scratch_doubles, void* stream);
int dfx_interpolate_dgbf_host(const dfx_basis* target, int32_t target_node, const dfx_basis* source, int32_t source_node,
                              const double* source_coeff_host, double* target_coeff_host, const uint8_t* mask_host);

/* ---- DiscreteGlobalBasisFunction local evaluation --------------------------
 * Replaces LocalFunctionBase::bind (discreteglobalbasisfunction.hh:124-148),
 * LocalFunction::operator() (:336-371) and
 * DiscreteGlobalBasisFunctionDerivative::LocalFunction::operator() (:583-627)
 * for all elements [elem_begin, elem_end) at the n_q reference-element points
 * xhat_host[q*dim + j].
 *   out_values[(e-elem_begin)*n_q*n_comp + q*n_comp + c]          (may be NULL)
 *   out_jacobians[((e-elem_begin)*n_q + q)*n_comp*dim + c*dim + j]  (may be NULL)
 * n_comp = number of leaves below subtree_node (range entry per leaf).        */
int dfx_evaluate(const dfx_basis* b, int32_t subtree_node, const double* coeff,
                 const double* xhat_host, int32_t n_q,
                 uint64_t elem_begin, uint64_t elem_end,
                 double* out_values, double* out_jacobians, void* stream);
/* DiscreteGlobalBasisFunction::operator()(x) and its derivative for points in WORLD coordinates
 * (discreteglobalbasisfunction.hh:402-410, :637-650).  The reference locates the element with a
 * HierarchicSearch ("very slow" by its own comment); on the structured grid the element is a
 * closed form: per direction the lowest element whose interval contains x up to 64 eps (the first
 * match of the reference's search), then the local function is evaluated at geometry.local(x).
 *   x[p*dim + j]                         n_points world coordinates (device)
 *   out_values[p*n_comp + c]             (may be NULL)
 *   out_jacobians[(p*n_comp + c)*dim+j]  (may be NULL)
 *   out_element[p]                       element index in the local box, -1 if no element of the
 *                                        local box contains the point (results NaN; the reference
 *                                        throws GridError) — may be NULL                          */
int dfx_evaluate_global(const dfx_basis* b, int32_t subtree_node, const double* coeff, const double* x,
                        uint64_t n_points, double* out_values, double* out_jacobians, int32_t* out_element, void* stream);
/* which kernel family dfx_evaluate picks for this configuration:
 * 0 generic, 1 sum-factorised register kernel, 2 sum-factorised shared-memory
 * (TMA bulk) kernel, 3 dense FP64 tensor-core (DMMA) kernel.                   */
int dfx_evaluate_path(const dfx_basis* b, int32_t subtree_node, const double* xhat_host,
                      int32_t n_q, int32_t want_jacobians);
/* force a kernel family (testing / profiling); path -1 = automatic.
 * These knobs are process-wide and meant for tests and profiling only (product code selects
 * behaviour per call, e.g. dfx_interpolate vs dfx_interpolate_separable).
 * op DFX_OP_EVALUATE: paths as above; DFX_OP_INTERPOLATE: 0 generic per-DOF
 * kernel, 1 structured kernel evaluating f at every node (the automatic choice), 2 structured
 * kernel with per-direction tables for separable f; DFX_OP_INDICES: 0 generic, 1 fast (shared-memory tile + bulk stores),
 * 2 fast with direct stores;
 * DFX_OP_HOST_BAND_KIB: `path` = KiB of results per pipeline band of dfx_evaluate_host
 * (-1 = default 192 MiB); DFX_OP_SMEM_GRADIENT: how the shared-memory evaluation kernel
 * forms gradients: 1 = differentiated shape tables in every stage, -1 = automatic
 * (collocation derivative of the point values when every direction has at least
 * order+1 distinct points).                                                        */
#define DFX_OP_EVALUATE 0
#define DFX_OP_INTERPOLATE 1
#define DFX_OP_INDICES 2
#define DFX_OP_HOST_BAND_KIB 3
#define DFX_OP_SMEM_GRADIENT 4
#define DFX_OP_TUNE_ROWS 5      /* lattice rows per block of the per-node interpolation kernel (-1 = default) */
#define DFX_OP_TUNE_EVAL_BLOCK 9  /* elements per block of the values-only register evaluation kernel: 32, 64 or 128 (-1 = default: 32) */
#define DFX_OP_TUNE_SMEM_TMA 8    /* shared-memory evaluation kernel: 0 no TMA pair loads of element-local coefficients, 2 also with gradients */
#define DFX_OP_TUNE_SMEM_PAIR 10   /* element-local DOFs, one leaf: 0 one element per warp at a time (k_eval_smem), 1 / 4 = k_eval_pair also for values-only calls with 8 / 4 warps per block, 2 = 8 warps with one bulk load per warp instead of one per block (-1 = default: k_eval_pair when gradients are asked for) */
#define DFX_OP_TUNE_SMEM_BLOCKS 11 /* > 0: cap on the grid of the persistent-warp evaluation kernels (testing: many items per warp); -1 = none */
#define DFX_OP_TUNE_EVAL_Q3 12     /* 0: continuous order 3 at 4^3 points stays on the shared-memory kernel (-1 = default: k_eval_q3) */
#define DFX_OP_TUNE_SPLIT_MAP 13   /* fused Taylor-Hood evaluation: 0 = thread per (element, leaf) with the leaf fastest in every warp (-1 / 1 = default: whole warps per leaf group — velocity warps, pressure warp) */
#define DFX_OP_TUNE_PAIR_PREFETCH 14 /* k_eval_pair: iterations ahead of the L2 prefetch of the coefficients (-1 / 0 = default: none) */
#define DFX_OP_TUNE_STREAM_HINT 7 /* 0: result stores without the L2 evict-first policy (-1 = default: with) */
#define DFX_OP_TUNE_SWIZZLE 6   /* evaluation block schedule: 0 grid-view order, 2 strips whenever the shape allows, -1 / 1 = automatic */
void dfx_set_kernel_path(int32_t op, int32_t path);

/* ---- host-buffer entry points (what a CPU caller of the reference uses) ----
 * Same semantics with HOST buffers (pinned buffers make the copies asynchronous);
 * the call returns when the result is in the host buffer.  dfx_evaluate_host is
 * pipelined over bands of z-layers: the coefficients a band needs travel to the
 * device on one stream while the previous band is evaluated and its results
 * travel back on another, so both PCIe directions are busy at once.           */
int dfx_interpolate_host(const dfx_basis* b, int32_t subtree_node, const dfx_function* f_host,
                         double* coeff_host, const uint8_t* mask_host);
int dfx_evaluate_host(const dfx_basis* b, int32_t subtree_node, const double* coeff_host,
                      const double* xhat_host, int32_t n_q,
                      uint64_t elem_begin, uint64_t elem_end,
                      double* out_values_host, double* out_jacobians_host);
int dfx_evaluate_global_host(const dfx_basis* b, int32_t subtree_node, const double* coeff_host, const double* x_host,
                             uint64_t n_points, double* out_values_host, double* out_jacobians_host,
                             int32_t* out_element_host);
int dfx_indices_flat_host(const dfx_basis* b, uint64_t elem_begin, uint64_t elem_end,
                          uint32_t* out_host);
int dfx_indices_multi_host(const dfx_basis* b, uint64_t elem_begin, uint64_t elem_end,
                           uint64_t* out_digits_host, int32_t stride);
int dfx_interpolation_points_host(const dfx_basis* b, double* out_xyz_host /*[dimension][dim]*/,
                                  int32_t* out_leaf_host /*[dimension]*/);

/* ---- boundary DOFs (Dirichlet set-up) ---------------------------------------
 * SubEntityDOFs::bind(localView, subEntityIndex, subEntityCodim), subentitydofs.hh:71-95:
 * local indices (below subtree_node) of the DOFs attached to sub-entity
 * (sub_entity_index, sub_entity_codim) of the reference cube or to one of its
 * sub-sub-entities.  On a cube grid the answer is the same for every element, so this
 * is host bookkeeping.  contained_host[max_node_size] receives dofIsContained_ (may be
 * NULL), out_local_host[max_node_size] the contained local indices in the reference's
 * order (may be NULL), *n their number.  An intersection is (indexInInside, 1).       */
int dfx_subentity_dofs(const dfx_basis* b, int32_t subtree_node, int32_t sub_entity_index,
                       int32_t sub_entity_codim, uint8_t* contained_host,
                       int32_t* out_local_host, int32_t* n);
/* forEachBoundaryDOF(basis, [&](auto&& index){ mask[index] = 1; }), boundarydofs.hh:110-128
 * (usage: examples/poisson-pq2.cc:272-276, stokes-taylorhood.cc:419-438 with a sub-space
 * basis = subtree_node): sets mask[flat position] = 1 for every DOF of the subtree that is
 * attached to a sub-entity of a boundary intersection of the GLOBAL grid; all other mask
 * bytes are left untouched.  mask has dimension() bytes.                               */
int dfx_boundary_dofs(const dfx_basis* b, int32_t subtree_node, uint8_t* mask, void* stream);
int dfx_boundary_dofs_host(const dfx_basis* b, int32_t subtree_node, uint8_t* mask_host);

/* ---- element-loop callers: sparse operator set-up ---------------------------
 * What examples/poisson-pq2.cc builds on top of a global basis, on the device.  Rows and
 * columns are flat container positions; all pointers are device pointers.
 *
 * dfx_matrix_pattern — getOccupationPattern (poisson-pq2.cc:137-174) + MatrixIndexSet::
 * exportIdx: CSR pattern of every pair of DOFs that share an element, ascending columns per
 * row.  Call once with col_idx == NULL to get row_ptr[dimension()+1] (and the number of
 * non-zeros in *nnz_host, may be NULL), allocate, call again to fill col_idx[nnz].      */
int dfx_matrix_pattern(const dfx_basis* b, uint64_t* row_ptr, uint32_t* col_idx, uint64_t* nnz_host, void* stream);
/* assembleLaplaceMatrix (poisson-pq2.cc:176-255, getLocalMatrix :35-87, getVolumeTerm :91-134)
 * for a single-leaf (scalar) basis: values[nnz] in the order of the pattern and rhs[dimension()]
 * for the volume term f (registry function); either output may be NULL.                     */
int dfx_assemble_laplace(const dfx_basis* b, const uint64_t* row_ptr, const uint32_t* col_idx, double* values,
                         const dfx_function* volume_term_host, double* rhs, void* stream);
/* BCRSMatrix::mv (subtract == 0: y = A x) / mmv (subtract != 0: y -= A x), rows = dimension() */
int dfx_csr_mv(const dfx_basis* b, const uint64_t* row_ptr, const uint32_t* col_idx, const double* values,
               const double* x, double* y, int32_t subtract, void* stream);
/* symmetric incorporation of Dirichlet values (poisson-pq2.cc:356-386): rhs -= A x (x holds the
 * Dirichlet values, zero elsewhere), then Dirichlet rows and columns become those of the identity
 * and rhs[i] = x[i] on them.  dirichlet: mask bytes as produced by dfx_boundary_dofs.          */
int dfx_apply_dirichlet(const dfx_basis* b, const uint64_t* row_ptr, const uint32_t* col_idx, double* values,
                        const uint8_t* dirichlet, const double* x, double* rhs, void* stream);

/* assembleStokesMatrix (examples/stokes-taylorhood.cc:252-297, getLocalMatrix :49-160) for a
 * Taylor-Hood tree composite(power<dim>(lagrange<2>), lagrange<1>) (or taylorHood()) with any index
 * merging strategies: values[nnz] in the order of the pattern of dfx_matrix_pattern — the example's
 * nested matrix flattened to container positions; velocity-velocity blocks of different
 * components and the pressure-pressure block are structural zeros.                           */
int dfx_assemble_stokes(const dfx_basis* b, const uint64_t* row_ptr, const uint32_t* col_idx, double* values, void* stream);
int dfx_assemble_stokes_host(const dfx_basis* b, const uint64_t* row_ptr_host, const uint32_t* col_idx_host,
                             double* values_host);
/* y = A x with the same Laplace stiffness matrix, never assembled (matrix-free): per element
 * gather, sum-factorised gradients at (order+1)^dim Gauss points, scale, transposed contraction;
 * every DOF then sums its element contributions in ascending element order (deterministic).
 * Continuous Lagrange order 1 or 2 in 2-d / 3-d.  dirichlet (may be NULL): mask bytes as produced by
 * dfx_boundary_dofs — those rows and columns behave like the identity, i.e. the result equals the
 * product with the matrix after dfx_apply_dirichlet.  `scratch` holds max_node_size() *
 * num_elements() + dimension() doubles.                                                      */
int dfx_apply_laplace(const dfx_basis* b, const double* x, double* y, const uint8_t* dirichlet, double* scratch,
                      void* stream);
/* host-buffer forms of the two set-up calls (what a CPU caller of the example uses): same
 * semantics with HOST arrays; col_idx_host == NULL returns row_ptr and *nnz only.          */
int dfx_matrix_pattern_host(const dfx_basis* b, uint64_t* row_ptr_host, uint32_t* col_idx_host, uint64_t* nnz_host);
int dfx_assemble_laplace_host(const dfx_basis* b, const uint64_t* row_ptr_host, const uint32_t* col_idx_host,
                              double* values_host, const dfx_function* volume_term_host, double* rhs_host);

/* ---- slab partition helpers (multi-GPU, one handle per rank) ---------------
 * When the local box spans the whole grid in every direction but the last,
 * the DOFs of one z-layer of every Yasp index-set component are contiguous.
 * dfx_halo_ranges lists, per leaf, the contiguous flat-position ranges of the
 * lowest (`side`=0) or highest (`side`=1) lattice plane of the local box:
 * the data one rank sends to / receives from its z-neighbour.
 * out_begin/out_count: [max_ranges]; returns the number of ranges in *n.     */
int dfx_halo_ranges(const dfx_basis* b, int32_t side, uint64_t* out_begin_host,
                    uint64_t* out_count_host, int32_t max_ranges, int32_t* n);
/* pack/unpack those ranges to/from one contiguous buffer (device)            */
uint64_t dfx_halo_size(const dfx_basis* b);
int dfx_halo_pack(const dfx_basis* b, int32_t side, const double* coeff, double* buf, void* stream);
int dfx_halo_unpack(const dfx_basis* b, int32_t side, const double* buf, double* coeff, void* stream);
/* owned range bookkeeping for gathering a distributed vector: per leaf and
 * index-set component, (local flat begin, count, global flat begin) of the
 * DOFs this rank owns (the top plane belongs to the upper neighbour unless
 * this is the last slab).                                                     */
int dfx_owned_ranges(const dfx_basis* b, uint64_t* local_begin_host, uint64_t* count_host,
                     uint64_t* global_begin_host, int32_t max_ranges, int32_t* n);

/* ---- peer-memory halo link (one process per GPU, NVLink) -------------------------
 * The sender packs its bottom lattice plane straight into the lower neighbour's receive buffer
 * (P2P stores from the pack kernel) and raises an arrival counter there; the receiver waits on its
 * own counter on the stream, unpacks into its top plane and acknowledges into the sender's block —
 * one kernel per direction (wait, move, fence, flag).
 * No rendezvous and no host synchronisation; two receive buffers let a sender run a step ahead.
 * Set-up: every rank creates a link (allocates its receive block), exports the 64-byte CUDA IPC
 * handle, and attaches the handles of its lower (`which` = 0) and upper (`which` = 1) neighbours;
 * inside one process a raw device pointer (dfx_halo_link_block) can be attached instead.
 * `step` counts exchanges from 1 and must be the same on both sides of a link.              */
typedef struct dfx_halo_link dfx_halo_link;
int   dfx_halo_link_create(const dfx_basis* b, dfx_halo_link** out);
void  dfx_halo_link_destroy(dfx_halo_link* l);
int   dfx_halo_link_export(dfx_halo_link* l, uint8_t ipc_handle_out[64]);
void* dfx_halo_link_block(dfx_halo_link* l);
int   dfx_halo_link_attach(dfx_halo_link* l, int32_t which, const uint8_t* ipc_handle, void* raw_block);
/* 0 if no wait of this link has timed out so far, else the step that did (synchronises with the
 * device).  The error is sticky: the exchange that timed out and every later dfx_halo_push / _pull on
 * the link move NOTHING (no stale plane is unpacked, no unconsumed buffer is overwritten) until
 * dfx_halo_link_reset_error — so check it before results that depend on the plane are used.  The wait
 * limit defaults to 4 s (dfx_halo_link_set_timeout; a slow neighbour, e.g. first-call set-up, needs more). */
uint64_t dfx_halo_link_error(dfx_halo_link* l);
int   dfx_halo_link_reset_error(dfx_halo_link* l);
int   dfx_halo_link_set_timeout(dfx_halo_link* l, double seconds);
int dfx_halo_push(const dfx_basis* b, dfx_halo_link* l, const double* coeff, uint64_t step, void* stream);
int dfx_halo_pull(const dfx_basis* b, dfx_halo_link* l, double* coeff, uint64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFX_H */
