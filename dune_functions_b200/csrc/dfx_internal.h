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
  void* dDev = nullptr;                                 // device copy of `dev`
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

// 1-d Lagrange shape functions on nodes m/k, as LagrangeCubeLocalBasis evaluates them
double lagrangeP(int k, int i, double x);
double lagrangeDP(int k, int i, double x);

// peer-memory halo link (dfx_kernels.cu): block layout [buf 0 | buf 1 | arrival | consumed | error]
struct dfx_halo_link_impl {
  int device = 0;
  u64 n = 0;                       // doubles per plane buffer
  double* local = nullptr;         // this rank's block (receives from the upper neighbour)
  double* lower = nullptr;         // lower neighbour's block (we push into it)
  double* upper = nullptr;         // upper neighbour's block (we acknowledge into it)
  bool lowerIpc = false, upperIpc = false;
};
int haloLinkBytes(u64 n, size_t& bytes);

}  // namespace dfx

struct dfx_basis : dfx::dfx_basis_impl {};
struct dfx_halo_link : dfx::dfx_halo_link_impl {};
