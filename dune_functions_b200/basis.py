"""Host-side mirror of the reference's interface for the element-loop path.

Names follow dune-functions (BasisFactory: makeBasis / lagrange / lagrangeDG /
power / composite / taylorHood and the four index-merging tags; interpolate;
makeDiscreteGlobalBasisFunction; subspaceBasis), cf.
  dune/functions/functionspacebases/defaultglobalbasis.hh:205-209
  dune/functions/functionspacebases/basistags.hh:192-225
  dune/functions/functionspacebases/interpolate.hh:204-302
  dune/functions/gridfunctions/discreteglobalbasisfunction.hh:457-478
  dune/python/functions/globalbasis.hh:141-252 (the Python front door)
Everything computes through the C ABI (include/dfx.h) on device memory held in
torch tensors; torch is used for allocation, streams and torch.distributed only.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _capi as capi


# ---- index merging strategy tags (basistags.hh:192-225) ----------------------
def flatLexicographic():
    return capi.IMS_FLAT_LEXICOGRAPHIC


def flatInterleaved():
    return capi.IMS_FLAT_INTERLEAVED


def blockedLexicographic():
    return capi.IMS_BLOCKED_LEXICOGRAPHIC


def blockedInterleaved():
    return capi.IMS_BLOCKED_INTERLEAVED


# ---- pre-basis factories -----------------------------------------------------
@dataclass
class PreBasisFactory:
    kind: int
    order: int = 0
    ims: int = capi.IMS_NONE
    children: List["PreBasisFactory"] = field(default_factory=list)
    exponent: int = 0

    def nodes(self):
        """preorder dfx_tree_node list"""
        if self.kind == capi.NODE_POWER:
            return [(self.kind, 0, self.exponent, self.ims)] + self.children[0].nodes()
        if self.kind == capi.NODE_COMPOSITE:
            out = [(self.kind, 0, len(self.children), self.ims)]
            for c in self.children:
                out += c.nodes()
            return out
        return [(self.kind, self.order, 0, capi.IMS_NONE)]


def lagrange(k):
    """lagrange<k>() / lagrange(k): lagrangebasis.hh:1006-1027"""
    return PreBasisFactory(capi.NODE_LAGRANGE, order=int(k))


def lagrangeDG(k):
    """lagrangeDG<k>() / lagrangeDG(k): lagrangedgbasis.hh:150-170"""
    return PreBasisFactory(capi.NODE_LAGRANGE_DG, order=int(k))


def power(child, exponent, ims=None):
    """power<k>(child, ims): powerbasis.hh:140-171.  The reference deprecates the
    implicit default; it is BlockedInterleaved (powerbasis.hh:160-171)."""
    if ims is None:
        ims = capi.IMS_BLOCKED_INTERLEAVED
    return PreBasisFactory(capi.NODE_POWER, ims=ims, children=[child], exponent=int(exponent))


def composite(*args):
    """composite(children..., ims): compositebasis.hh:420-445 (default BlockedLexicographic)."""
    args = list(args)
    ims = capi.IMS_BLOCKED_LEXICOGRAPHIC
    if args and isinstance(args[-1], int):
        ims = args.pop()
    return PreBasisFactory(capi.NODE_COMPOSITE, ims=ims, children=args)


def taylorHood():
    """taylorHood(): taylorhoodbasis.hh:330-335"""
    return PreBasisFactory(capi.NODE_TAYLOR_HOOD)


# ---- grid --------------------------------------------------------------------
@dataclass
class YaspGrid:
    """Unrefined equidistant structured grid (Dune::YaspGrid<dim>(L, N) stand-in).

    `local_begin/local_n` select the box of elements owned by this rank; see
    dfx_grid_desc in include/dfx.h."""
    n: Sequence[int]
    upper: Optional[Sequence[float]] = None
    lower: Optional[Sequence[float]] = None
    local_begin: Optional[Sequence[int]] = None
    local_n: Optional[Sequence[int]] = None
    # ids of the grid's globalIdSet for the vertices of the local box (lexicographic, x fastest);
    # None = the grid's own monotone ids.  See dfx_basis_set_vertex_ids in include/dfx.h.
    vertex_ids: Optional[Sequence[int]] = None

    @property
    def dim(self):
        return len(self.n)

    def desc(self):
        d = self.dim
        g = capi.GridDesc()
        g.dim = d
        lo = list(self.lower) if self.lower is not None else [0.0] * d
        up = list(self.upper) if self.upper is not None else [1.0] * d
        lb = list(self.local_begin) if self.local_begin is not None else [0] * d
        ln = list(self.local_n) if self.local_n is not None else list(self.n)
        for j in range(3):
            g.n_global[j] = self.n[j] if j < d else 1
            g.local_begin[j] = lb[j] if j < d else 0
            g.local_n[j] = ln[j] if j < d else 1
            g.lower[j] = lo[j] if j < d else 0.0
            g.upper[j] = up[j] if j < d else 1.0
        return g

    def slab(self, rank, world):
        """z-slab (last direction) of this grid for `rank` of `world`."""
        d = self.dim
        nz = self.n[d - 1]
        b = (nz * rank) // world
        e = (nz * (rank + 1)) // world
        lb = [0] * d
        ln = list(self.n)
        lb[d - 1] = b
        ln[d - 1] = e - b
        ids = None
        if self.vertex_ids is not None:
            # z is the slowest direction of the vertex lattice: a slab's vertices are one contiguous run
            plane = 1
            for j in range(d - 1):
                plane *= self.n[j] + 1
            ids = self.vertex_ids[b * plane:(e + 1) * plane]
        return YaspGrid(self.n, self.upper, self.lower, lb, ln, ids)


def tree_array(factory):
    nodes = factory.nodes()
    arr = (capi.TreeNode * len(nodes))()
    for i, (k, o, c, m) in enumerate(nodes):
        arr[i].kind, arr[i].order, arr[i].n_children, arr[i].ims = k, o, c, m
    return arr, len(nodes)


# ---- functions ---------------------------------------------------------------
def make_function(fn_id, n_comp=1, params=()):
    f = capi.Function()
    f.id = fn_id
    f.n_comp = n_comp
    for i, p in enumerate(params):
        f.p[i] = float(p)
    return f


def gaussian(a=1.0):
    """x -> exp(-a |x|^2)   (examples/interpolation.cc:49-52)"""
    return make_function(capi.FN_GAUSSIAN, 1, [a])


def identity(n_comp):
    """x -> x   (examples/interpolation.cc:93-95)"""
    return make_function(capi.FN_IDENTITY, n_comp)


def constant(*values):
    return make_function(capi.FN_CONSTANT, len(values), values)


def coord(j):
    return make_function(capi.FN_COORD, 1, [j])


def quadratic(c0, lin, quad):
    """c0 + lin.x + sum_{i<=j} quad[(i,j)] x_i x_j, quad order (00,01,02,11,12,22)"""
    return make_function(capi.FN_QUADRATIC, 1, [c0] + list(lin) + list(quad))


def sinprod(w):
    return make_function(capi.FN_SINPROD, 1, [w])


def gaussian_vec(n_comp, a=1.0):
    return make_function(capi.FN_GAUSSIAN_VEC, n_comp, [a])


# ---- basis -------------------------------------------------------------------
def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _check_range(begin, end, n):
    """the element range check of the ABI (DFX_ERR_RANGE), done before any buffer is sized from it"""
    if not (0 <= begin <= end <= n):
        raise capi.RangeError(f"element range [{begin}, {end}) outside the local box of {n} elements")


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class GlobalBasis:
    """DefaultGlobalBasis (defaultglobalbasis.hh:51-191) over the C ABI."""

    def __init__(self, grid: YaspGrid, factory: PreBasisFactory, device: int = 0):
        self.grid = grid
        self.factory = factory
        self.device = device
        self._lib = capi.lib()
        self._desc = grid.desc()
        arr, n = tree_array(factory)
        h = C.c_void_p()
        capi.check(self._lib.dfx_basis_create(C.byref(self._desc), arr, n, device, C.byref(h)))
        self._h = h
        if getattr(grid, "vertex_ids", None) is not None:
            self.setVertexIds(grid.vertex_ids)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.dfx_basis_destroy(h)
            self._h = None

    def update(self, grid):
        """basis.update(gridView), defaultglobalbasis.hh:137-141: the same tree on another grid"""
        desc = grid.desc()
        capi.check(self._lib.dfx_basis_update(self._h, C.byref(desc)))
        self.grid, self._desc = grid, desc
        self._numbering = getattr(self, "_numbering", 0) + 1
        if getattr(grid, "vertex_ids", None) is not None:
            self.setVertexIds(grid.vertex_ids)

    def setVertexIds(self, ids):
        """ids of grid().globalIdSet() for the vertices of the local box, the input of
        LagrangeFaceDOFPermutation (lagrangebasis.hh:787); None restores the grid's own ids.
        Accepts a host sequence / numpy array or a CUDA uint64/int64 torch tensor."""
        self._numbering = getattr(self, "_numbering", 0) + 1           # local views refetch their index window
        if ids is None:
            capi.check(self._lib.dfx_basis_set_vertex_ids_host(self._h, None, 0))
            return
        try:
            import torch
            is_dev = isinstance(ids, torch.Tensor) and ids.is_cuda
        except ImportError:
            is_dev = False
        if is_dev:
            t = ids.contiguous()
            assert t.dtype in (torch.int64, torch.uint64), "vertex ids: 64-bit integers expected"
            capi.check(self._lib.dfx_basis_set_vertex_ids(self._h, t.data_ptr(), t.numel(), _stream()))
            return
        import numpy as np
        a = np.ascontiguousarray(np.asarray(ids), dtype=np.uint64)
        capi.check(self._lib.dfx_basis_set_vertex_ids_host(self._h, a.ctypes.data, a.size))

    def containerDescriptor(self):
        return _container_descriptor(self)

    def hasPermutedFaceDOFs(self):
        return bool(self._lib.dfx_basis_has_permuted_face_dofs(self._h))

    # GlobalBasis concept
    def dimension(self):
        return int(self._lib.dfx_basis_dimension(self._h))

    def size(self, prefix=()):
        p = (C.c_uint64 * max(1, len(prefix)))(*prefix)
        return int(self._lib.dfx_basis_size(self._h, p, len(prefix)))

    def maxNodeSize(self):
        return int(self._lib.dfx_basis_max_node_size(self._h))

    def maxMultiIndexSize(self):
        return int(self._lib.dfx_basis_max_multi_index_size(self._h))

    def minMultiIndexSize(self):
        return int(self._lib.dfx_basis_min_multi_index_size(self._h))

    def numElements(self):
        return int(self._lib.dfx_basis_num_elements(self._h))

    def numLeaves(self):
        return int(self._lib.dfx_basis_num_leaves(self._h))

    def numTreeNodes(self):
        return int(self._lib.dfx_basis_num_tree_nodes(self._h))

    def nodeInfo(self, node):
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        capi.check(self._lib.dfx_basis_node_info(self._h, node, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(local_offset=a.value, size=b.value, first_leaf=c.value, n_leaves=d.value)

    def child(self, node, *path):
        """tree node id reached from `node` along a child path (TypeTree::child)"""
        for c in path:
            node = int(self._lib.dfx_basis_node_child(self._h, node, c))
            if node < 0:
                raise capi.RangeError("no such child")
        return node

    def localIndexSizes(self):
        out = (C.c_uint8 * self.maxNodeSize())()
        capi.check(self._lib.dfx_basis_local_index_sizes(self._h, out))
        return np.frombuffer(out, dtype=np.uint8).copy()

    # batched LocalView::bind + index
    def indices(self, elem_begin=0, elem_end=None, dtype=None):
        """multi-indices of all local DOFs: tensor [ne, n_loc, maxMultiIndexSize]"""
        import torch
        ne_all = self.numElements()
        elem_end = ne_all if elem_end is None else elem_end
        _check_range(elem_begin, elem_end, ne_all)
        ne = elem_end - elem_begin
        stride = self.maxMultiIndexSize()
        dtype = dtype or torch.int32
        out = torch.empty((ne, self.maxNodeSize(), stride), dtype=dtype, device=f"cuda:{self.device}")
        fn = self._lib.dfx_indices_multi if dtype == torch.int32 else self._lib.dfx_indices_multi_u64
        capi.check(fn(self._h, elem_begin, elem_end, _ptr(out), stride, _stream()))
        return out

    def flatIndices(self, elem_begin=0, elem_end=None, dtype=None):
        """flat container positions of all local DOFs: tensor [ne, n_loc]"""
        import torch
        ne_all = self.numElements()
        elem_end = ne_all if elem_end is None else elem_end
        _check_range(elem_begin, elem_end, ne_all)
        ne = elem_end - elem_begin
        dtype = dtype or torch.int32
        out = torch.empty((ne, self.maxNodeSize()), dtype=dtype, device=f"cuda:{self.device}")
        fn = self._lib.dfx_indices_flat if dtype == torch.int32 else self._lib.dfx_indices_flat_u64
        capi.check(fn(self._h, elem_begin, elem_end, _ptr(out), _stream()))
        return out

    # the reference's Python front door (dune/python/functions/globalbasis.hh:141-252)
    def __len__(self):
        return self.dimension()

    def localView(self):
        return LocalView(self)

    def interpolate(self, coeff, f, mask=None):
        """basis.interpolate(dofVector, f): globalbasis.hh:199-210 / dune/python/functions/interpolate.hh:17-28"""
        return interpolate(self, coeff, f, mask)

    def asFunction(self, coeff):
        """basis.asFunction(dofVector) -> DiscreteFunction, globalbasis.hh:245-248"""
        return DiscreteGlobalBasisFunction(self, coeff)

    def newVector(self, fill=None):
        import torch
        v = torch.empty(self.dimension(), dtype=torch.float64, device=f"cuda:{self.device}")
        if fill is not None:
            v.fill_(fill)
        return v


class LocalView:
    """LocalViewWrapper of the reference's Python binding (globalbasis.hh:82-121, :179-192):
    bind(element) / unbind() / index(i) / len() / size() / maxSize() / tree().  An element is its
    position in elements(gridView) (x fastest) or its coordinate tuple.  index(i) is served from a
    window of the batched device index table (one device call per 16384 elements)."""

    WINDOW = 1 << 14

    def __init__(self, basis):
        self.basis = basis
        self._nloc = basis.maxNodeSize()
        self._lens = basis.localIndexSizes()
        self._e = None
        self._w0 = self._w1 = 0
        self._cache = None
        self._numbering = getattr(basis, "_numbering", 0)

    def bind(self, element):
        if self._numbering != getattr(self.basis, "_numbering", 0):
            # basis.update(gridView) / new vertex ids since the window was fetched: the numbering changed
            self._numbering = self.basis._numbering
            self._w0 = self._w1 = 0
            self._cache = None
            self._nloc = self.basis.maxNodeSize()
            self._lens = self.basis.localIndexSizes()
        g = self.basis.grid
        n = list(g.local_n) if g.local_n is not None else list(g.n)
        if not isinstance(element, (int, np.integer)):
            e, stride = 0, 1
            for j, c in enumerate(element):
                e += int(c) * stride
                stride *= n[j]
            element = e
        ne = self.basis.numElements()
        if not 0 <= element < ne:
            raise capi.RangeError("element outside the grid view")
        self._e = int(element)
        if not (self._w0 <= self._e < self._w1):
            self._w0 = self._e - self._e % self.WINDOW
            self._w1 = min(ne, self._w0 + self.WINDOW)
            import torch
            self._cache = self.basis.indices(self._w0, self._w1, dtype=torch.int64).cpu().numpy()

    def unbind(self):
        self._e = None

    def bound(self):
        return self._e is not None

    def element(self):
        return self._e

    def index(self, i):
        """global multi-index of local DOF i as a list of ints (LocalViewWrapper::index)"""
        if self._e is None:
            raise capi.DfxError("local view is not bound")
        return [int(v) for v in self._cache[self._e - self._w0, i, :self._lens[i]]]

    def __len__(self):
        return self._nloc

    def size(self):
        return self._nloc

    def maxSize(self):
        return self._nloc

    def tree(self):
        """root node info of the local ansatz tree: dict(local_offset, size, first_leaf, n_leaves)"""
        return self.basis.nodeInfo(0)


class ContainerDescriptor:
    """Run-time form of the descriptor types of containerdescriptors.hh:91-209 (Unknown, Value, Tuple,
    Array, Vector, UniformArray, UniformVector; FlatVector = UniformVector of Value): `size()` children,
    `cd[i]` the i-th child descriptor (the shared prototype for the uniform kinds), `type()` the
    reference type spelled out."""
    KINDS = {capi.CD_UNKNOWN: "Unknown", capi.CD_VALUE: "Value", capi.CD_TUPLE: "Tuple", capi.CD_ARRAY: "Array",
             capi.CD_VECTOR: "Vector", capi.CD_UNIFORM_ARRAY: "UniformArray", capi.CD_UNIFORM_VECTOR: "UniformVector"}

    def __init__(self, kind, size, children):
        self.kind, self._size, self.children = kind, size, children

    def size(self):
        return self._size

    def __len__(self):
        return self._size

    def __getitem__(self, i):
        if self.kind == capi.CD_VALUE:
            return self                                    # Value::operator[] returns a Value (:104-105)
        if self.kind == capi.CD_UNKNOWN:
            raise IndexError("Unknown container descriptor has no children")
        if not 0 <= i < self._size:
            raise IndexError("child index out of range")
        return self.children[0] if self.kind in (capi.CD_UNIFORM_ARRAY, capi.CD_UNIFORM_VECTOR) else self.children[i]

    def isFlatVector(self):
        return self.kind == capi.CD_UNIFORM_VECTOR and self.children[0].kind == capi.CD_VALUE

    def type(self):
        k = self.KINDS[self.kind]
        if self.kind in (capi.CD_UNKNOWN, capi.CD_VALUE):
            return k
        if self.kind == capi.CD_TUPLE:
            return "Tuple<" + ",".join(c.type() for c in self.children) + ">"
        inner = self.children[0].type()
        return f"{k}<{inner}>" if self.kind in (capi.CD_UNIFORM_VECTOR, capi.CD_VECTOR) else f"{k}<{inner},{self._size}>"

    def __repr__(self):
        return self.type()


def _container_descriptor(basis):
    """basis.containerDescriptor(), defaultglobalbasis.hh:180-186 (a SubspaceBasis forwards to its root,
    subspacebasis.hh:115-118)"""
    root, _ = _root_and_node(basis)
    n = C.c_int32()
    capi.check(root._lib.dfx_basis_container_descriptor(root._h, None, 0, C.byref(n), None, 0))
    arr = (capi.ContainerNode * n.value)()
    capi.check(root._lib.dfx_basis_container_descriptor(root._h, arr, n.value, C.byref(n), None, 0))
    pos = 0

    def build():
        nonlocal pos
        nd = arr[pos]
        pos += 1
        kids = [build() for _ in range(nd.n_stored)]
        return ContainerDescriptor(nd.kind, int(nd.size), kids)
    return build()


def makeBasis(grid, factory, device=0):
    """BasisFactory::makeBasis, defaultglobalbasis.hh:205-209"""
    return GlobalBasis(grid, factory, device)


class SubspaceBasis:
    """subspaceBasis(rootBasis, path...), subspacebasis.hh:27-157: same root indices,
    tree = child at the prefix path; dimension()/size() are the root's."""

    def __init__(self, root, *path):
        # a SubspaceBasis as root: no nesting, the tree paths are joined (subspacebasis.hh:130-133)
        if isinstance(root, SubspaceBasis):
            path = tuple(root.path) + tuple(path)
            root = root.root
        self.root = root
        self.path = tuple(path)
        self.node = root.child(0, *self.path)

    def rootBasis(self):
        return self.root

    def prefixPath(self):
        return self.path

    def dimension(self):
        return self.root.dimension()

    def containerDescriptor(self):
        return _container_descriptor(self)


def subspaceBasis(root, *path):
    return SubspaceBasis(root, *path)


def _root_and_node(basis):
    if isinstance(basis, SubspaceBasis):
        return basis.root, basis.node
    return basis, 0


def interpolate(basis, coeff, f, mask=None, separable=False):
    """interpolate(basis, coeff, f[, bitVector]) — interpolate.hh:204-302.

    coeff: float64 CUDA tensor of basis.dimension() entries (flat container order),
    f: a dfx_function from the registry (gaussian(), identity(n), ...) or a
       DiscreteGlobalBasisFunction over any basis on the same grid (a grid function: its local
       function is evaluated at the interpolation nodes, gridviewfunction.hh:68-75),
    mask: optional uint8/bool CUDA tensor in the same order,
    separable: opt into per-direction factor tables for registry functions that are products of 1-d
       factors (dfx_interpolate_separable: O(N) transcendentals, <= 3 ulp from the per-node result);
       the default evaluates f at every node like the reference."""
    root, node = _root_and_node(basis)
    if coeff.numel() != root.dimension():
        raise capi.RangeError("coefficient vector size does not match basis.dimension()")
    m = None
    if mask is not None:
        import torch
        m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    if isinstance(f, DiscreteGlobalBasisFunction):
        capi.check(root._lib.dfx_interpolate_dgbf(root._h, node, f.root._h, f.node, _ptr(f.coeff), _ptr(coeff), _ptr(m),
                                                  None, 0, _stream()))
        return coeff
    entry = root._lib.dfx_interpolate_separable if separable else root._lib.dfx_interpolate
    capi.check(entry(root._h, node, C.byref(f), _ptr(coeff), _ptr(m), _stream()))
    return coeff


def interpolateBatch(basis, coeff, parts, mask=None):
    """several interpolate calls on one vector as one call (dfx_interpolate_batch): parts = [(space, f), ...]
    with `space` the basis, a sub-space basis of it, or a tree-node id; equivalent to interpolate(space, coeff, f)
    for each part in order.  Two parts whose kernels are shorter than a launch gap share one launch."""
    root, _ = _root_and_node(basis)
    if coeff.numel() != root.dimension():
        raise capi.RangeError("coefficient vector size does not match basis.dimension()")
    nodes, fns = [], []
    for space, f in parts:
        if isinstance(space, int):
            nodes.append(space)
        else:
            r, node = _root_and_node(space)
            if r is not root:
                raise capi.DfxError("all parts must be sub-spaces of the same basis")
            nodes.append(node)
        fns.append(f)
    n = len(nodes)
    arr_n = (C.c_int32 * max(1, n))(*nodes)
    arr_f = (capi.Function * max(1, n))(*fns)
    m = None
    if mask is not None:
        import torch
        m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    capi.check(root._lib.dfx_interpolate_batch(root._h, n, arr_n, arr_f, _ptr(coeff), _ptr(m), _stream()))
    return coeff


def interpolationPoints(basis):
    """positions of all interpolation nodes and the leaf of every coefficient (form 2 of f)"""
    import torch
    n = basis.dimension()
    xyz = torch.empty((n, basis.grid.dim), dtype=torch.float64, device=f"cuda:{basis.device}")
    leaf = torch.empty(n, dtype=torch.int32, device=f"cuda:{basis.device}")
    capi.check(basis._lib.dfx_interpolation_points(basis._h, _ptr(xyz), _ptr(leaf), _stream()))
    return xyz, leaf


def interpolateFromValues(basis, coeff, values, mask=None):
    root, node = _root_and_node(basis)
    capi.check(root._lib.dfx_interpolate_from_values(root._h, node, _ptr(values), _ptr(coeff), _ptr(mask), _stream()))
    return coeff


def subEntityDOFs(basis, subEntityIndex, subEntityCodim):
    """subEntityDOFs(localView, subEntityIndex, subEntityCodim) — subentitydofs.hh:71-95, 182-190.
    Returns (contained local indices in the reference's order, dofIsContained flags); on a cube
    grid they are the same for every element.  An intersection is (indexInInside, 1)."""
    root, node = _root_and_node(basis)
    nloc = root.maxNodeSize()
    flags = (C.c_uint8 * nloc)()
    lst = (C.c_int32 * nloc)()
    n = C.c_int32()
    capi.check(root._lib.dfx_subentity_dofs(root._h, node, subEntityIndex, subEntityCodim, flags, lst, C.byref(n)))
    return [int(lst[i]) for i in range(n.value)], np.frombuffer(flags, dtype=np.uint8).astype(bool)


def boundaryDOFs(basis, mask=None):
    """forEachBoundaryDOF(basis, [&](auto&& index){ mask[index] = true; }) — boundarydofs.hh:110-128.
    mask: uint8/bool CUDA tensor of dimension() entries in flat container order (created zeroed
    when omitted); entries that are not boundary DOFs of the (sub-space) basis keep their value."""
    import torch
    root, node = _root_and_node(basis)
    if mask is None:
        mask = torch.zeros(root.dimension(), dtype=torch.uint8, device=f"cuda:{root.device}")
    if mask.numel() != root.dimension():
        raise capi.RangeError("mask size does not match basis.dimension()")
    m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    capi.check(root._lib.dfx_boundary_dofs(root._h, node, _ptr(m), _stream()))
    return mask


class DiscreteGlobalBasisFunction:
    """makeDiscreteGlobalBasisFunction<R>(basis, coeff) — discreteglobalbasisfunction.hh:281-478.

    evaluate(xhat) is localFunction(f) bound to every element in turn and evaluated
    at the reference-element points xhat [n_q, dim]:
       values    [ne, n_q, n_comp]
       jacobians [ne, n_q, n_comp, dim]   (derivative(localFunction), :495-677)"""

    def __init__(self, basis, coeff):
        self.basis = basis
        self.root, self.node = _root_and_node(basis)
        if coeff.numel() != self.root.dimension():
            raise capi.RangeError("coefficient vector size does not match basis.dimension()")
        self.coeff = coeff

    def nComp(self):
        return self.root.nodeInfo(self.node)["n_leaves"]

    def evaluate(self, xhat, values=True, jacobians=False, elem_begin=0, elem_end=None, out_values=None,
                 out_jacobians=None):
        import torch
        root = self.root
        xh = np.ascontiguousarray(np.asarray(xhat, dtype=np.float64).reshape(-1, root.grid.dim))
        nq = xh.shape[0]
        ne_all = root.numElements()
        elem_end = ne_all if elem_end is None else elem_end
        _check_range(elem_begin, elem_end, ne_all)
        ne = elem_end - elem_begin
        nc = self.nComp()
        dev = f"cuda:{root.device}"
        v = j = None
        if values:
            v = out_values if out_values is not None else torch.empty((ne, nq, nc), dtype=torch.float64, device=dev)
        if jacobians:
            j = out_jacobians if out_jacobians is not None else torch.empty((ne, nq, nc, root.grid.dim),
                                                                           dtype=torch.float64, device=dev)
        capi.check(root._lib.dfx_evaluate(root._h, self.node, _ptr(self.coeff),
                                          xh.ctypes.data_as(C.POINTER(C.c_double)), nq,
                                          elem_begin, elem_end, _ptr(v), _ptr(j), _stream()))
        if values and jacobians:
            return v, j
        return v if values else j


def _evaluate_global(self, x, values=True, jacobians=False, return_elements=False):
    """operator()(x) for world coordinates x [n, dim] (CUDA tensor or array):
    values [n, n_comp], jacobians [n, n_comp, dim]; NaN (and element -1) where no element of the
    local box contains the point — discreteglobalbasisfunction.hh:402-410."""
    import torch
    root = self.root
    dev = f"cuda:{root.device}"
    xs = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))
    xs = xs.to(device=dev, dtype=torch.float64).reshape(-1, root.grid.dim).contiguous()
    n, nc = xs.shape[0], self.nComp()
    v = torch.empty((n, nc), dtype=torch.float64, device=dev) if values else None
    j = torch.empty((n, nc, root.grid.dim), dtype=torch.float64, device=dev) if jacobians else None
    el = torch.empty(n, dtype=torch.int32, device=dev) if return_elements else None
    capi.check(root._lib.dfx_evaluate_global(root._h, self.node, _ptr(self.coeff), _ptr(xs), n, _ptr(v), _ptr(j), _ptr(el),
                                             _stream()))
    out = tuple(t for t in (v, j, el) if t is not None)
    return out[0] if len(out) == 1 else out


DiscreteGlobalBasisFunction.evaluateGlobal = _evaluate_global


def makeDiscreteGlobalBasisFunction(basis, coeff):
    return DiscreteGlobalBasisFunction(basis, coeff)


# ---- quadrature helpers (tensor Gauss-Legendre on [0,1]^d) --------------------
def gauss_points_1d(m):
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w


def tensor_gauss(dim, m, last_fastest=True):
    """Tensor Gauss rule with m points per direction.  dune-geometry's tensor-product
    rule runs the LAST coordinate fastest."""
    x1, w1 = gauss_points_1d(m)
    pts, wts = [], []
    idx = np.indices((m,) * dim).reshape(dim, -1).T
    for mi in idx:
        # np.indices varies the last axis fastest
        mi = mi if last_fastest else mi[::-1]
        pts.append([x1[i] for i in mi])
        wts.append(np.prod([w1[i] for i in mi]))
    return np.array(pts), np.array(wts)
