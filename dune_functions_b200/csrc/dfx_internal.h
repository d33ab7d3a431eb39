// dfx_internal.h — internal types shared by the host-side basis code and the kernels.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/dfx.h"

#if defined(__CUDACC__)
#define DFX_HD __host__ __device__ __forceinline__
#else
#define DFX_HD inline
#endif

namespace dfx {

typedef unsigned long long u64;

constexpr int kMaxLayouts = 4;      // distinct (kind, order) DOF lattices in one tree
constexpr int kMaxContinuousOrder = 17;  // LagrangeFaceDOFPermutation::maxOrder3d, lagrangebasis.hh:579
constexpr int kMaxDGOrder = 12;

// ---- device-side description of the grid box ------------------------------
struct DevGrid {
  int dim;
  int N[3];      // local box, elements per direction (1 beyond dim)
  int G[3];      // global grid
  int off[3];    // local_begin
  double lower[3];
  double h[3];   // (upper-lower)/G   — YaspGrid mesh width
};

// ---- DOF lattice of one (kind, order) on the local box ---------------------
// Continuous Lagrange: DOFs sit on the index-set components of the grid.  A
// component is the set of grid entities with a given "shift" mask s (bit j set =
// entity extends along direction j); MCMGMapper numbers codim 0 first ... vertices
// last, Yasp orders the components of a codim by ascending s, and each is
// lexicographic with x fastest (numbering assumptions).
struct DevLayout {
  int kind;        // DFX_NODE_LAGRANGE or DFX_NODE_LAGRANGE_DG
  int k;           // order
  int nloc;        // (k+1)^dim
  int ncomp;       // number of non-empty components
  u64 dimension;   // number of scalar DOFs
  u64 compOffset[8];   // scalar index of the first DOF of component s
  u64 compCount[8];    // DOFs in component s
  int blk[8];          // DOFs per entity of component s: (k-1)^popcount(s)
  int csize[8][3];     // entities per direction
  int order[8];        // components sorted by compOffset (first ncomp entries)
};

// ---- one leaf of the local ansatz tree -------------------------------------
// Every digit of the leaf's multi-index and its flat container position are
// affine in the leaf's scalar index j (Appendix B of SURVEY.md).
struct DevLeaf {
  int layout;
  int localOffset;
  int nloc;
  int ndig;
  u64 posA, posB;
  u64 digA[DFX_MAX_DIGITS], digB[DFX_MAX_DIGITS];
};

struct DevBasis {
  DevGrid grid;
  int nLeaves, nLayouts, nlocTotal, maxDigits;
  DevLayout layouts[kMaxLayouts];
  DevLeaf leaves[DFX_MAX_LEAVES];
  // ids of the grid's globalIdSet for the vertices of the local box, lexicographic (device memory,
  // owned by the handle); nullptr = the structured grid's own ids, monotone in lexicographic
  // position, for which every face orientation is the identity (lagrangebasis.hh:787, :615-634).
  // Last members on purpose: the offsets of everything the tuned kernels read stay what they were.
  const u64* vid;
  // orientation codes (orientCode) of the local box's edges and quadrilateral facets, one byte per entity:
  // component s (bit j set = direction j is free) starts at orientOff[s], entities lexicographic with
  // N_j (free) or N_j + 1 (fixed) per direction — the entity index of the DOF layout.  nullptr until built
  // (ensureOrientTable, after dfx_basis_set_vertex_ids).
  const unsigned char* orient;
  unsigned orientOff[8];
};

// scalar index of local node alpha of element e in a continuous/DG layout
DFX_HD u64 layoutIndex(const DevLayout& L, const DevGrid& g, const int e[3], const int a[3], int li) {
  if (L.kind == DFX_NODE_LAGRANGE_DG) {
    u64 ei = (u64)e[0] + (u64)g.N[0] * ((u64)e[1] + (u64)g.N[1] * (u64)e[2]);
    return ei * (u64)L.nloc + (u64)li;
  }
  const int k = L.k;
  unsigned s = 0;
  int c[3];
  u64 fr = 0, frs = 1;
  if (k == 0) {
    s = (1u << g.dim) - 1u;
    c[0] = e[0]; c[1] = e[1]; c[2] = e[2];
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (a[j] > 0 && a[j] < k) { s |= 1u << j; c[j] = e[j]; fr += (u64)(a[j] - 1) * frs; frs *= (u64)(k - 1); }
      else c[j] = e[j] + (a[j] == k ? 1 : 0);
    }
  }
  u64 ent = (u64)c[0] + (u64)L.csize[s][0] * ((u64)c[1] + (u64)L.csize[s][1] * (u64)c[2]);
  return L.compOffset[s] + ent * (u64)L.blk[s] + fr;
}

// ---- face-DOF permutation of continuous Lagrange k >= 3 ----------------------
// Experimental::LagrangeFaceDOFPermutation::permuteFaceDOF (lagrangebasis.hh:652-675) for cube
// grids whose vertex ids are not monotone.  `fr` is the index of a node inside its entity as the
// ELEMENT numbers it (LocalKey::index: lexicographic over the entity's free directions, base k-1);
// the result is the index with respect to the entity's globally unique orientation — and the other
// way round with INVERSE.  Edges (:232-235): flipped iff id(first vertex) > id(second).  Quadrilateral
// facets (:279-331, table :518-559): r counter-clockwise rotations bring the vertex with the smallest
// id to corner 0, then a reflection across the diagonal if that vertex's second neighbour has a
// smaller id than its first; the vertex with the smallest id is found directly (the reference derives
// it from the edge bits to save comparisons, ":291-297 ... is equivalent to").  The orientation depends
// on the entity only, so every element that contains it computes the same one.
// The orientation depends on the entity only, so it splits into a CODE per entity (edges: 1 = flipped; facets:
// rotations r in bits 0-1, reflection in bit 2; 0 = identity) and the permutation the code stands for.  The code
// table of the local box is built once per dfx_basis_set_vertex_ids (k_orient_codes), after which the kernels
// read one byte per entity instead of two or four 64-bit ids.
DFX_HD unsigned orientCode(const DevGrid& g, const u64* vid, unsigned s, const int c[3]) {
  const u64 v0s = (u64)(g.N[0] + 1), v1s = v0s * (u64)(g.N[1] + 1);
  const u64 base = (u64)c[0] + v0s * (u64)c[1] + v1s * (u64)c[2];
  const u64 st[3] = {1, v0s, v1s};
  const int nfree = ((s >> 0) & 1) + ((s >> 1) & 1) + ((s >> 2) & 1);
  if (nfree == 1 && g.dim > 1) {
    const int f = (s & 1u) ? 0 : ((s & 2u) ? 1 : 2);
    return vid[base] > vid[base + st[f]] ? 1u : 0u;
  }
  if (nfree == 2 && g.dim == 3) {
    const int f0 = (s & 1u) ? 0 : 1, f1 = (s & 4u) ? 2 : 1;
    const u64 id0 = vid[base], id1 = vid[base + st[f0]], id2 = vid[base + st[f1]], id3 = vid[base + st[f0] + st[f1]];
    unsigned r, refl;
    if (id0 < id1 && id0 < id2 && id0 < id3) { r = 0; refl = id2 < id1; }
    else if (id1 < id2 && id1 < id3) { r = 3; refl = id0 < id3; }
    else if (id2 < id3) { r = 1; refl = id3 < id0; }
    else { r = 2; refl = id1 < id2; }
    return r | (refl << 2);
  }
  return 0u;
}

// nfree: free directions of the entity (1 edge, 2 quadrilateral facet); anything else is the identity
template <bool INVERSE>
DFX_HD constexpr unsigned applyOrientCode(unsigned code, int nfree, int k, unsigned fr) {
  if (nfree == 1) return code ? (unsigned)(k - 2) - fr : fr;
  if (nfree == 2) {
    unsigned r = code & 3u;
    const unsigned refl = code >> 2;
    const unsigned m = (unsigned)(k - 1);
    unsigned c0 = fr % m, c1 = fr / m, t = 0;
    if (INVERSE) {
      if (refl) { t = c0; c0 = c1; c1 = t; }
      r = (4u - r) & 3u;
    }
    if (r == 1) { t = c0; c0 = m - 1 - c1; c1 = t; }            // swap, then mirror c0
    else if (r == 2) { c0 = m - 1 - c0; c1 = m - 1 - c1; }
    else if (r == 3) { t = c0; c0 = c1; c1 = m - 1 - t; }       // mirror c0, then swap
    if (!INVERSE && refl) { t = c0; c0 = c1; c1 = t; }
    return c0 + m * c1;
  }
  return fr;
}

template <bool INVERSE>
DFX_HD unsigned orientFaceDOF(const DevGrid& g, const u64* vid, int k, unsigned s, const int c[3], unsigned fr) {
  const int nfree = ((s >> 0) & 1) + ((s >> 1) & 1) + ((s >> 2) & 1);
  if (!((nfree == 1 && g.dim > 1) || (nfree == 2 && g.dim == 3))) return fr;
  return applyOrientCode<INVERSE>(orientCode(g, vid, s, c), nfree, k, fr);
}

// layoutIndex with the face-DOF permutation applied (PreBasis::indices, lagrangebasis.hh:863-869)
DFX_HD u64 layoutIndexOriented(const DevLayout& L, const DevGrid& g, const u64* vid, const int e[3], const int a[3], int li) {
  if (vid == nullptr || L.kind != DFX_NODE_LAGRANGE || L.k < 3 || g.dim < 2) return layoutIndex(L, g, e, a, li);
  const int k = L.k;
  unsigned s = 0, fr = 0, frs = 1;
  int c[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (a[j] > 0 && a[j] < k) { s |= 1u << j; c[j] = e[j]; fr += (unsigned)(a[j] - 1) * frs; frs *= (unsigned)(k - 1); }
    else c[j] = e[j] + (a[j] == k ? 1 : 0);
  }
  fr = orientFaceDOF<false>(g, vid, k, s, c, fr);
  const u64 ent = (u64)c[0] + (u64)L.csize[s][0] * ((u64)c[1] + (u64)L.csize[s][1] * (u64)c[2]);
  return L.compOffset[s] + ent * (u64)L.blk[s] + (u64)fr;
}

// The element that owns scalar DOF t of layout L in interpolate(), and the DOF's lattice multi-index in
// it: the LAST element (in grid-view iteration order of the GLOBAL grid) that contains the node — its
// write is the one that survives in the reference loop (interpolate.hh:254-259 visits the elements in
// order and :166-171 overwrites).  ie = GLOBAL element coordinates, a = alpha in {0..k}^dim.
DFX_HD void dofOwnerNode(const DevBasis& B, const DevLayout& L, u64 t, int ie[3], int a[3]) {
  const DevGrid& g = B.grid;
  const int k = L.k;
  ie[0] = ie[1] = ie[2] = 0;
  a[0] = a[1] = a[2] = 0;
  if (L.kind == DFX_NODE_LAGRANGE_DG) {
    const u64 e = t / (u64)L.nloc;
    int i = (int)(t - e * (u64)L.nloc);
    const int n1 = k + 1;
    a[0] = i % n1; i /= n1;
    a[1] = i % n1; i /= n1;
    a[2] = i;
    ie[0] = (int)(e % (u64)g.N[0]) + g.off[0];
    const u64 r = e / (u64)g.N[0];
    ie[1] = (int)(r % (u64)g.N[1]) + g.off[1];
    ie[2] = (int)(r / (u64)g.N[1]) + g.off[2];
    return;
  }
  // component of t
  int s = L.order[0];
  for (int q = 1; q < L.ncomp; ++q)
    if (t >= L.compOffset[L.order[q]]) s = L.order[q];
  const u64 r = t - L.compOffset[s];
  const u64 ent = r / (u64)L.blk[s];
  unsigned fr = (unsigned)(r - ent * (u64)L.blk[s]);
  int c[3];
  c[0] = (int)(ent % (u64)L.csize[s][0]);
  const u64 r2 = ent / (u64)L.csize[s][0];
  c[1] = (int)(r2 % (u64)L.csize[s][1]);
  c[2] = (int)(r2 / (u64)L.csize[s][1]);
  // t counts the entity's nodes in its globally unique orientation; the element that writes the node
  // numbers them in its own (lagrangebasis.hh:868, inverted)
  if (B.vid != nullptr && k > 2 && g.dim > 1) fr = orientFaceDOF<true>(g, B.vid, k, (unsigned)s, c, fr);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j >= g.dim) continue;
    if ((s >> j) & 1) {            // node strictly inside the entity along j
      ie[j] = g.off[j] + c[j];
      if (k > 1) { a[j] = (int)(fr % (unsigned)(k - 1)) + 1; fr /= (unsigned)(k - 1); }
    } else {                       // node on a lattice plane: owner = upper neighbour if any
      const int gc = g.off[j] + c[j];
      if (gc < g.G[j]) { ie[j] = gc; a[j] = 0; } else { ie[j] = g.G[j] - 1; a[j] = k; }
    }
  }
}

// ---- per local index, folded once on the host for the index-table kernel (csrc/dfx_interp_fast.cu) ----
// flat position of local index li in element (ec0, ec1, ec2): P0 + S * ent with ent = ec0 + cs0 (ec1 + cs1 ec2);
// digit p of its multi-index: dB[p] + dA[p] * ent  (identity orientations)
struct IdxLi {
  u64 P0;
  unsigned S, cs0, cs1, ndig;
  u64 dB[DFX_MAX_DIGITS], dA[DFX_MAX_DIGITS];
};

// ---- host-side basis --------------------------------------------------------
struct HostNode {           // expanded local ansatz tree, preorder
  int kind, order, ims;
  int localOffset, size, firstLeaf, nLeaves;
  std::vector<int> children;
};

struct PreNode;             // pre-basis tree (power holds ONE child), see dfx_basis.cu

struct dfx_basis_impl {
  dfx_grid_desc gridDesc;
  int device;
  DevBasis dev;
  std::shared_ptr<PreNode> pre;
  std::vector<HostNode> nodes;
  std::vector<dfx_tree_node> treeDesc;   // the descriptor the handle was created from
  u64 dimension;
  int minDigits, maxDigits;
  // device scratch owned by the handle
  std::vector<unsigned> localTab;                       // per local index: leaf | alpha (see k_indices_fast)
  void* dLocalTab = nullptr;                            // uploaded on first use
  std::vector<IdxLi> idxTab; void* dIdxTab = nullptr;   // per local index: folded constants of the index table
  void* dDev = nullptr;                                 // device copy of `dev`
  void* dVid = nullptr;                                 // vertex ids (dev.vid points here)
  void* dOrient = nullptr; size_t dOrientBytes = 0; bool orientValid = false;   // orientation codes (dev.orient)
  // evaluation with permuted face DOFs: the same tree with every leaf discontinuous, and the
  // element-local coefficients gathered for it (csrc/dfx_kernels.cu, k_gather_local)
  dfx_basis_impl* companion = nullptr; bool companionTried = false;
  void* dLocalCoeff = nullptr; size_t dLocalCoeffBytes = 0;
  std::vector<u64> vidHost; bool vidPending = false;    // ids given from host memory, uploaded at the next device call
  bool oriented = false;                                // ids given AND some continuous leaf of order >= 3 in 2-d/3-d
  void* dSep = nullptr; size_t dSepBytes = 0;           // per-direction tables of a separable f
  void* dShape = nullptr; size_t dShapeBytes = 0;       // shape tables of the last evaluate
  void* dNodal = nullptr; size_t dNodalBytes = 0;       // nodal values of a grid function (dfx_interpolate_dgbf)
  void* dJinv = nullptr;                                // diag(J^-1) per direction and element coordinate
  void* dDense = nullptr; size_t dDenseBytes = 0;       // dense table of the DMMA path
  std::vector<double> denseXhat; int dDenseNode = -1, dDenseCols = 0, dDenseVal = 0;
  std::vector<double> lastXhat; int lastShapeNode = -1;
  void* pinned = nullptr; size_t pinnedBytes = 0;       // staging for *_host entry points
  void* dScratch[3] = {nullptr, nullptr, nullptr}; size_t dScratchBytes[3] = {0, 0, 0};
  void* stream = nullptr;                               // private stream of the *_host entry points
  void* evStream[2] = {nullptr, nullptr};
};

// error plumbing
void setError(const std::string& msg);
int fail(int code, const std::string& msg);
const char* lastErrorCStr();

// host basis construction (dfx_basis.cu)
int buildBasis(const dfx_grid_desc* g, const dfx_tree_node* t, int n, int device, dfx_basis_impl** out);
u64 preSize(const dfx_basis_impl* b, const u64* prefix, int len);
void containerDescriptor(const dfx_basis_impl* b, std::vector<dfx_container_node>& nodes, std::string& type);

// 1-d Lagrange shape functions on nodes m/k, as LagrangeCubeLocalBasis evaluates them
double lagrangeP(int k, int i, double x);
double lagrangeDP(int k, int i, double x);

// peer-memory halo link (dfx_kernels.cu): block layout [buf 0 | buf 1 | arrival | consumed | error | timeout | counters]
struct dfx_halo_link_impl {
  int device = 0;
  u64 n = 0;                       // doubles per plane buffer
  double* local = nullptr;         // this rank's block (receives from the upper neighbour)
  double* lower = nullptr;         // lower neighbour's block (we push into it)
  double* upper = nullptr;         // upper neighbour's block (we acknowledge into it)
  bool lowerIpc = false, upperIpc = false;
};
int haloLinkBytes(u64 n, size_t& bytes);
size_t haloLinkWordOffset(u64 n, int which);     // byte offset of control word `which` (2 = error, 3 = timeout in clock cycles)

}  // namespace dfx

struct dfx_basis : dfx::dfx_basis_impl {};
struct dfx_halo_link : dfx::dfx_halo_link_impl {};
