// dfx_operators.cu — the element-loop CALLERS next to the hot path:
// what examples/poisson-pq2.cc does with a global basis, done on the device.
//
//   k_pattern<FILL>     getOccupationPattern (poisson-pq2.cc:137-174) + MatrixIndexSet::exportIdx:
//                       CSR pattern of all pairs of DOFs that share an element, in closed form per
//                       row (the union of the row's elements is a lattice box), columns sorted.
//   k_laplace_rows      assembleLaplaceMatrix (:176-255) with getLocalMatrix (:35-87): every matrix
//                       row is summed by ONE warp over the elements that contain the row's node, in
//                       ascending element order — the reference's accumulation order, no atomics.
//   k_volume_points /   getVolumeTerm (:91-134): f at the quadrature points of every element, then the
//   k_volume_rows       same owner-computes row gather.
//   k_csr_mv            BCRSMatrix::mv / mmv (y = A x, y -= A x), one warp per row, fixed reduction tree.
//   k_dirichlet         symmetric incorporation of Dirichlet values (:356-386).
//   k_laplace_local /   the same stiffness matrix applied WITHOUT assembling it (y = A x): per element
//   k_laplace_gather    gather, sum-factorised gradients at (k+1)^d Gauss points, scale, transposed
//                       contraction -> element-local results [local index][element]; then every DOF
//                       sums its <= 2^d contributions in ascending element order (deterministic,
//                       no atomics).  560 B per Q2 element instead of 6 KB for the assembled SpMV.
//
// Rows and columns are flat container positions (include/dfx.h), so the pattern works for any
// tree; the Laplace assembly is for a single-leaf (scalar) basis like the example.
#include <algorithm>
#include <cmath>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "dfx_device.cuh"
#include "dfx_launch.h"
#include "dfx_sumfactor.cuh"

namespace dfx {

namespace {

constexpr int kRowWarps = 4;             // warps (= rows in flight) per block
constexpr int kMaxRowLen = 2048;         // longest row the shared-memory staging holds (pattern, 32-bit keys)
constexpr int kMaxLaplaceRow = 1024;     // longest row of the Laplace assembler (FP64 accumulators)

// the node of scalar DOF t of a continuous / element-local layout: entity coordinates c, in-entity
// lattice offsets r (0 = pinned to the lattice plane c_j), shift mask s
struct RowNode { int c[3]; int r[3]; int s; int e; int li; };

__device__ __forceinline__ RowNode locateNode(const DevLayout& L, const DevGrid& g, u64 t) {
  RowNode nd;
  const int k = L.k;
  if (L.kind == DFX_NODE_LAGRANGE_DG || k == 0) {
    const u64 e = t / (u64)L.nloc;
    nd.li = (int)(t - e * (u64)L.nloc);
    nd.c[0] = (int)(e % (u64)g.N[0]);
    const u64 q = e / (u64)g.N[0];
    nd.c[1] = (int)(q % (u64)g.N[1]);
    nd.c[2] = (int)(q / (u64)g.N[1]);
    nd.r[0] = nd.r[1] = nd.r[2] = 0;
    nd.s = 7; nd.e = 1;
    return nd;
  }
  int s = L.order[0];
  for (int q = 1; q < L.ncomp; ++q)
    if (t >= L.compOffset[L.order[q]]) s = L.order[q];
  const u64 rr = t - L.compOffset[s];
  const u64 ent = rr / (u64)L.blk[s];
  unsigned fr = (unsigned)(rr - ent * (u64)L.blk[s]);
  nd.c[0] = (int)(ent % (u64)L.csize[s][0]);
  const u64 q2 = ent / (u64)L.csize[s][0];
  nd.c[1] = (int)(q2 % (u64)L.csize[s][1]);
  nd.c[2] = (int)(q2 / (u64)L.csize[s][1]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    nd.r[j] = 0;
    if ((s >> j) & 1) { nd.r[j] = (int)(fr % (unsigned)(k - 1)) + 1; fr /= (unsigned)(k - 1); }
  }
  nd.s = s; nd.e = 0; nd.li = 0;
  return nd;
}

// elements that contain the node: per direction [lo, hi] (local box coordinates)
__device__ __forceinline__ void elementBox(const DevGrid& g, const RowNode& nd, int lo[3], int hi[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j >= g.dim || nd.e || ((nd.s >> j) & 1)) { lo[j] = hi[j] = nd.c[j]; continue; }
    lo[j] = nd.c[j] > 0 ? nd.c[j] - 1 : 0;
    hi[j] = nd.c[j] < g.N[j] ? nd.c[j] : g.N[j] - 1;
  }
}

// number of DOFs of leaf layout L inside the element box, and the v-th of them as a scalar index
__device__ __forceinline__ unsigned boxCount(const DevLayout& L, const DevGrid& g, const int lo[3], const int hi[3]) {
  unsigned n = 1;
  const bool cell = L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0;
  for (int j = 0; j < g.dim; ++j) n *= cell ? (unsigned)(hi[j] - lo[j] + 1) : (unsigned)(L.k * (hi[j] - lo[j] + 1) + 1);
  return cell ? n * (unsigned)L.nloc : n;
}
__device__ __forceinline__ u64 boxIndex(const DevLayout& L, const DevGrid& g, const int lo[3], const int hi[3], unsigned v) {
  const int k = L.k;
  if (L.kind == DFX_NODE_LAGRANGE_DG || k == 0) {
    const unsigned li = v % (unsigned)L.nloc; v /= (unsigned)L.nloc;
    int e[3] = {0, 0, 0};
    for (int j = 0; j < g.dim; ++j) { const unsigned w = (unsigned)(hi[j] - lo[j] + 1); e[j] = lo[j] + (int)(v % w); v /= w; }
    int a[3];
    localAlpha(k, (int)li, a);
    return layoutIndex(L, g, e, a, (int)li);
  }
  int c[3] = {0, 0, 0}, a[3] = {0, 0, 0};
  for (int j = 0; j < g.dim; ++j) {
    const unsigned w = (unsigned)(k * (hi[j] - lo[j] + 1) + 1);
    const int X = k * lo[j] + (int)(v % w); v /= w;
    c[j] = X / k; a[j] = X - c[j] * k;       // a = 0: pinned to lattice plane c (layoutIndex: e + [a == k] = c)
  }
  return layoutIndex(L, g, c, a, 0);
}

// ---- a single continuous leaf: the columns of a row in ASCENDING index order, in closed form ----
// The scalar index is compOffset[s'] + blk[s'] * lex(entity) + in-entity offset with the components
// ordered by compOffset, so the sorted columns are: for every component s' (in L.order) the
// entities of the element box in lexicographic order (x fastest; W_j = ne_j entities along an
// extended direction, ne_j + 1 lattice planes along a pinned one), each with its blk in-entity nodes.
struct BoxRank {
  unsigned prefix[9];      // prefix[q] = number of columns in components L.order[0..q)
  unsigned W[8][3];        // entities of the box per direction, per component (indexed by position q)
};
__device__ __forceinline__ void boxRankInit(const DevLayout& L, const DevGrid& g, const int lo[3], const int hi[3], BoxRank& R) {
  unsigned run = 0;
  for (int q = 0; q < L.ncomp; ++q) {
    const int s = L.order[q];
    unsigned n = (unsigned)L.blk[s];
    for (int j = 0; j < 3; ++j) {
      const unsigned ne = j < g.dim ? (unsigned)(hi[j] - lo[j] + 1) : 1u;
      R.W[q][j] = j >= g.dim ? 1u : (((s >> j) & 1) ? ne : ne + 1u);
      n *= R.W[q][j];
    }
    R.prefix[q] = run;
    run += n;
  }
  R.prefix[L.ncomp] = run;
}
// the v-th smallest column of the row
__device__ __forceinline__ u64 boxColumn(const DevLayout& L, const int lo[3], const BoxRank& R, unsigned v) {
  int q = 0;
  while (q + 1 < L.ncomp && v >= R.prefix[q + 1]) ++q;
  const int s = L.order[q];
  unsigned r = v - R.prefix[q];
  const unsigned blk = (unsigned)L.blk[s];
  const unsigned fr = r % blk; r /= blk;
  const unsigned e0 = r % R.W[q][0]; r /= R.W[q][0];
  const unsigned e1 = r % R.W[q][1]; r /= R.W[q][1];
  const u64 ent = (u64)(lo[0] + (int)e0) + (u64)L.csize[s][0] * ((u64)(lo[1] + (int)e1) + (u64)L.csize[s][1] * (u64)(lo[2] + (int)r));
  return L.compOffset[s] + ent * (u64)blk + fr;
}
// rank of lattice node alpha of element e (inside the box) among the row's columns
__device__ __forceinline__ unsigned boxRankOf(const DevLayout& L, const int lo[3], const BoxRank& R, const int e[3], const int a[3]) {
  const int k = L.k;
  int s = 0, c[3];
  unsigned fr = 0, frs = 1;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (a[j] > 0 && a[j] < k) { s |= 1 << j; c[j] = e[j]; fr += (unsigned)(a[j] - 1) * frs; frs *= (unsigned)(k - 1); }
    else c[j] = e[j] + (a[j] == k ? 1 : 0);
  }
  int q = 0;
  while (q + 1 < L.ncomp && L.order[q] != s) ++q;
  const unsigned ent = (unsigned)(c[0] - lo[0]) + R.W[q][0] * ((unsigned)(c[1] - lo[1]) + R.W[q][1] * (unsigned)(c[2] - lo[2]));
  return R.prefix[q] + ent * (unsigned)L.blk[s] + fr;
}

// in-warp bitonic sort of n <= kMaxRowLen keys in shared memory (padded with 0xffffffff)
__device__ __forceinline__ void warpSort(unsigned* key, int n, int lane) {
  int p2 = 32;
  while (p2 < n) p2 <<= 1;
  for (int i = n + lane; i < p2; i += 32) key[i] = 0xffffffffu;
  __syncwarp();
  for (int k = 2; k <= p2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < p2; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned a = key[i], b = key[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { key[i] = b; key[ixj] = a; }
        }
      }
      __syncwarp();
    }
}

// One warp per (scalar DOF t of layout layoutId, leaf on that layout) = one matrix row.
// FILL = false: row lengths into len[row]; FILL = true: sorted columns at col[rowPtr[row]...].
template <bool FILL>
__global__ void __launch_bounds__(kRowWarps * 32)
k_pattern(const DevBasis* __restrict__ Bp, int layoutId, u64* __restrict__ len, const u64* __restrict__ rowPtr,
          uint32_t* __restrict__ col) {
  __shared__ unsigned keys[kRowWarps][FILL ? kMaxRowLen : 1];
  const DevBasis& B = *Bp;
  const DevGrid& g = B.grid;
  const DevLayout& L = B.layouts[layoutId];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u64 t = (u64)blockIdx.x * kRowWarps + warp;
  if (t >= L.dimension) return;
  const RowNode nd = locateNode(L, g, t);
  int lo[3], hi[3];
  elementBox(g, nd, lo, hi);
  // the row's columns: every leaf's DOFs in the element box (the same set for every leaf of this node)
  unsigned total = 0;
  for (int l = 0; l < B.nLeaves; ++l) total += boxCount(B.layouts[B.leaves[l].layout], g, lo, hi);
  if (!FILL) {
    if (lane == 0)
      for (int l = 0; l < B.nLeaves; ++l)
        if (B.leaves[l].layout == layoutId) len[B.leaves[l].posA * t + B.leaves[l].posB] = total;
    return;
  }
  if (B.nLeaves == 1 && L.kind == DFX_NODE_LAGRANGE && L.k >= 1 && B.leaves[0].posA == 1) {
    // single continuous leaf: emit the columns in rank order, no sort
    BoxRank R;
    boxRankInit(L, g, lo, hi, R);
    const u64 row = t + B.leaves[0].posB;
    uint32_t* dst = col + rowPtr[row];
    for (unsigned v = lane; v < total; v += 32) dst[v] = (uint32_t)(boxColumn(L, lo, R, v) + B.leaves[0].posB);
    return;
  }
  unsigned* key = keys[warp];
  unsigned base = 0;
  for (int l = 0; l < B.nLeaves; ++l) {
    const DevLeaf& leaf = B.leaves[l];
    const DevLayout& LL = B.layouts[leaf.layout];
    const unsigned n = boxCount(LL, g, lo, hi);
    for (unsigned v = lane; v < n; v += 32) key[base + v] = (unsigned)(leaf.posA * boxIndex(LL, g, lo, hi, v) + leaf.posB);
    base += n;
  }
  __syncwarp();
  warpSort(key, (int)total, lane);
  for (int l = 0; l < B.nLeaves; ++l) {
    if (B.leaves[l].layout != layoutId) continue;
    const u64 row = B.leaves[l].posA * t + B.leaves[l].posB;
    uint32_t* dst = col + rowPtr[row];
    for (unsigned v = lane; v < total; v += 32) dst[v] = key[v];
  }
}

// position of column c in the sorted row segment [p0, p1)
__device__ __forceinline__ u64 findColumn(const uint32_t* __restrict__ col, u64 p0, u64 p1, uint32_t c) {
  while (p0 < p1) {
    const u64 m = (p0 + p1) >> 1;
    if (col[m] < c) p0 = m + 1; else p1 = m;
  }
  return p0;
}

struct LaplaceArgs {
  int nloc, nq;
  u64 posA, posB;
};

// Stiffness rows.  K[d][i][j] = sum_q w_q d_d phi_i(x_q) d_d phi_j(x_q) on the reference cube
// (quadrature order 2 (dim k - 1) like the example), so that the element matrix is
// sum_d |det J| (J^-1_dd)^2 K[d] for an axis-aligned element: the example's loop over quadrature
// points with the element-independent part summed first.
__global__ void __launch_bounds__(kRowWarps * 32)
k_laplace_rows(const DevBasis* __restrict__ Bp, const __grid_constant__ LaplaceArgs A, const double* __restrict__ K,
               const u64* __restrict__ rowPtr, const uint32_t* __restrict__ col, double* __restrict__ values) {
  __shared__ double acc[kRowWarps][kMaxLaplaceRow];
  const DevBasis& B = *Bp;
  const DevGrid& g = B.grid;
  const DevLayout& L = B.layouts[B.leaves[0].layout];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u64 t = (u64)blockIdx.x * kRowWarps + warp;
  if (t >= L.dimension) return;
  const RowNode nd = locateNode(L, g, t);
  int lo[3], hi[3];
  elementBox(g, nd, lo, hi);
  const u64 row = A.posA * t + A.posB;
  const u64 p0 = rowPtr[row], p1 = rowPtr[row + 1];
  const int rowLen = (int)(p1 - p0);
  double* a = acc[warp];
  for (int v = lane; v < rowLen; v += 32) a[v] = 0.0;
  __syncwarp();
  const int k = L.k, n = A.nloc;
  const bool ranked = L.kind == DFX_NODE_LAGRANGE && k >= 1;
  BoxRank BR;
  if (ranked) boxRankInit(L, g, lo, hi, BR);
  // ascending element index = the order in which the reference's element loop adds to this row
  for (int e2 = lo[2]; e2 <= hi[2]; ++e2)
    for (int e1 = lo[1]; e1 <= hi[1]; ++e1)
      for (int e0 = lo[0]; e0 <= hi[0]; ++e0) {
        const int e[3] = {e0, e1, e2};
        // local index of the row's node in this element
        int iLoc;
        if (nd.e) iLoc = nd.li;
        else {
          int al[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) al[j] = ((nd.s >> j) & 1) ? nd.r[j] : (j < g.dim && e[j] < nd.c[j] ? k : 0);
          iLoc = al[0] + (k + 1) * (al[1] + (k + 1) * al[2]);
        }
        double coef[3] = {0, 0, 0};
        {
          double detJ = 1.0, ji[3] = {1, 1, 1};
          for (int j = 0; j < g.dim; ++j) {
            const double l0 = gridCoordinate(g, j, g.off[j] + e[j]), u0 = gridCoordinate(g, j, g.off[j] + e[j] + 1);
            detJ *= u0 - l0; ji[j] = 1.0 / (u0 - l0);
          }
          for (int j = 0; j < g.dim; ++j) coef[j] = detJ * ji[j] * ji[j];
        }
        for (int jl = lane; jl < n; jl += 32) {
          int aj[3];
          localAlpha(k, jl, aj);
          u64 pos;
          if (ranked) pos = p0 + boxRankOf(L, lo, BR, e, aj);          // closed-form position in the sorted row
          else pos = findColumn(col, p0, p1, (uint32_t)(A.posA * layoutIndex(L, g, e, aj, jl) + A.posB));
          double s = 0.0;
          for (int d = 0; d < g.dim; ++d) s += coef[d] * K[((size_t)d * n + iLoc) * n + jl];
          a[pos - p0] += s;                 // distinct jl -> distinct columns: no two lanes meet
        }
        __syncwarp();
      }
  for (int v = lane; v < rowLen; v += 32) values[p0 + v] = a[v];
}

// Stokes rows for a Taylor-Hood tree composite(power<dim>(Q2), Q1) (examples/stokes-taylorhood.cc
// :49-160, :252-297).  Reference-cube integrals at the example's quadrature order:
//   K[d][i][j] = sum_q w_q d_d phi_i d_d phi_j   (Q2 x Q2),   G[d][i][j] = sum_q w_q d_d phi_i psi_j   (Q2 x Q1)
// so that for an axis-aligned element   A_{v_k v_k} = sum_d |det J| (J^-1_dd)^2 K[d],
// A_{v_k p} = A_{p v_k}^T = -|det J| J^-1_kk G[k],   everything else 0.  One warp per matrix row,
// elements in ascending order, columns located in the sorted row by binary search.
struct StokesArgs { int dim, nv, np; };

__global__ void __launch_bounds__(kRowWarps * 32)
k_stokes_rows(const DevBasis* __restrict__ Bp, const __grid_constant__ StokesArgs A, int rowLeaf,
              const double* __restrict__ K, const double* __restrict__ G, const u64* __restrict__ rowPtr,
              const uint32_t* __restrict__ col, double* __restrict__ values) {
  __shared__ double acc[kRowWarps][kMaxLaplaceRow];
  const DevBasis& B = *Bp;
  const DevGrid& g = B.grid;
  const DevLeaf& rleaf = B.leaves[rowLeaf];
  const DevLayout& L = B.layouts[rleaf.layout];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u64 t = (u64)blockIdx.x * kRowWarps + warp;
  if (t >= L.dimension) return;
  const RowNode nd = locateNode(L, g, t);
  int lo[3], hi[3];
  elementBox(g, nd, lo, hi);
  const u64 row = rleaf.posA * t + rleaf.posB;
  const u64 p0 = rowPtr[row], p1 = rowPtr[row + 1];
  const int rowLen = (int)(p1 - p0);
  double* a = acc[warp];
  for (int v = lane; v < rowLen; v += 32) a[v] = 0.0;
  __syncwarp();
  const int dim = A.dim, nv = A.nv, np = A.np, k = L.k;
  const bool rowIsPressure = rowLeaf == dim;
  const DevLayout& LV = B.layouts[B.leaves[0].layout];
  const DevLayout& LP = B.layouts[B.leaves[dim].layout];
  for (int e2 = lo[2]; e2 <= hi[2]; ++e2)
    for (int e1 = lo[1]; e1 <= hi[1]; ++e1)
      for (int e0 = lo[0]; e0 <= hi[0]; ++e0) {
        const int e[3] = {e0, e1, e2};
        int al[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) al[j] = ((nd.s >> j) & 1) ? nd.r[j] : (j < g.dim && e[j] < nd.c[j] ? k : 0);
        const int iLoc = al[0] + (k + 1) * (al[1] + (k + 1) * al[2]);
        double detJ = 1.0, ji[3] = {1, 1, 1};
        for (int j = 0; j < dim; ++j) {
          const double l0 = gridCoordinate(g, j, g.off[j] + e[j]), u0 = gridCoordinate(g, j, g.off[j] + e[j] + 1);
          detJ *= u0 - l0; ji[j] = 1.0 / (u0 - l0);
        }
        // local columns: [v_0 (nv) | ... | v_{dim-1} (nv) | p (np)]
        const int nloc = dim * nv + np;
        for (int jl = lane; jl < nloc; jl += 32) {
          const int cl = jl < dim * nv ? jl / nv : dim;              // column leaf
          const int jc = jl - (cl < dim ? cl * nv : dim * nv);      // local index inside that leaf
          double s = 0.0;
          bool coupled = true;
          if (!rowIsPressure && cl < dim) {
            if (cl != rowLeaf) coupled = false;
            else for (int d = 0; d < dim; ++d) s += detJ * ji[d] * ji[d] * K[((size_t)d * nv + iLoc) * nv + jc];
          } else if (!rowIsPressure) {
            s = -detJ * ji[rowLeaf] * G[((size_t)rowLeaf * nv + iLoc) * np + jc];
          } else if (cl < dim) {
            s = -detJ * ji[cl] * G[((size_t)cl * nv + jc) * np + iLoc];
          } else coupled = false;
          if (!coupled) continue;                                    // structural zero: stays 0
          const DevLeaf& cleaf = B.leaves[cl];
          const DevLayout& LC = cl < dim ? LV : LP;
          int aj[3];
          localAlpha(LC.k, jc, aj);
          const uint32_t c = (uint32_t)(cleaf.posA * layoutIndex(LC, g, e, aj, jc) + cleaf.posB);
          a[findColumn(col, p0, p1, c) - p0] += s;
        }
        __syncwarp();
      }
  for (int v = lane; v < rowLen; v += 32) values[p0 + v] = a[v];
}

// f at the quadrature points: pts[e][q] = f(x_q(e)) * w_q * |det J(e)|
template <class F>
__global__ void __launch_bounds__(256)
k_volume_points(const __grid_constant__ DevGrid g, int nq, const double* __restrict__ xq /*[nq][dim]*/,
                const double* __restrict__ wq, const F f, u64 total, double* __restrict__ pts) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const u64 e = t / (u64)nq;
  const int q = (int)(t - e * (u64)nq);
  int ec[3];
  ec[0] = (int)(e % (u64)g.N[0]);
  const u64 r = e / (u64)g.N[0];
  ec[1] = (int)(r % (u64)g.N[1]);
  ec[2] = (int)(r / (u64)g.N[1]);
  double x[3] = {0, 0, 0}, detJ = 1.0;
  for (int j = 0; j < g.dim; ++j) {
    const double l0 = gridCoordinate(g, j, g.off[j] + ec[j]), u0 = gridCoordinate(g, j, g.off[j] + ec[j] + 1);
    x[j] = __dadd_rn(l0, __dmul_rn(xq[q * g.dim + j], __dsub_rn(u0, l0)));
    detJ *= u0 - l0;
  }
  const double cm = f.common(x[0], x[1], x[2]);
  pts[t] = f.comp(0, cm, x[0], x[1], x[2]) * wq[q] * detJ;
}

// rhs rows: sum over the node's elements (ascending) of sum_q phi_iLoc(x_q) pts[e][q]
__global__ void __launch_bounds__(kRowWarps * 32)
k_volume_rows(const DevBasis* __restrict__ Bp, const __grid_constant__ LaplaceArgs A, const double* __restrict__ phi /*[nq][nloc]*/,
              const double* __restrict__ pts, double* __restrict__ rhs) {
  const DevBasis& B = *Bp;
  const DevGrid& g = B.grid;
  const DevLayout& L = B.layouts[B.leaves[0].layout];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u64 t = (u64)blockIdx.x * kRowWarps + warp;
  if (t >= L.dimension) return;
  const RowNode nd = locateNode(L, g, t);
  int lo[3], hi[3];
  elementBox(g, nd, lo, hi);
  const int k = L.k;
  double total = 0.0;
  for (int e2 = lo[2]; e2 <= hi[2]; ++e2)
    for (int e1 = lo[1]; e1 <= hi[1]; ++e1)
      for (int e0 = lo[0]; e0 <= hi[0]; ++e0) {
        int iLoc;
        if (nd.e) iLoc = nd.li;
        else {
          const int e[3] = {e0, e1, e2};
          int al[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) al[j] = ((nd.s >> j) & 1) ? nd.r[j] : (j < g.dim && e[j] < nd.c[j] ? k : 0);
          iLoc = al[0] + (k + 1) * (al[1] + (k + 1) * al[2]);
        }
        const u64 ei = (u64)e0 + (u64)g.N[0] * ((u64)e1 + (u64)g.N[1] * (u64)e2);
        double s = 0.0;
        for (int q = lane; q < A.nq; q += 32) s += phi[(size_t)q * A.nloc + iLoc] * pts[ei * (u64)A.nq + q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        total += s;
      }
  if (lane == 0) rhs[A.posA * t + A.posB] = total;
}

// y = A x (SUB = false) or y -= A x (SUB = true).  LPR lanes share one row (CSR-vector): LPR is
// picked from the mean row length so that short rows (Q1: 27 entries) do not idle most of a warp;
// the reduction tree inside the lane group is fixed, so results do not depend on the launch.
template <bool SUB, int LPR>
__global__ void __launch_bounds__(256)
k_csr_mv(u64 n, const u64* __restrict__ rowPtr, const uint32_t* __restrict__ col, const double* __restrict__ values,
         const double* __restrict__ x, double* __restrict__ y) {
  const u64 row = ((u64)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int lane = threadIdx.x & (LPR - 1);
  const bool active = row < n;
  double s = 0.0;
  if (active) {
    const u64 p0 = rowPtr[row], p1 = rowPtr[row + 1];
    for (u64 p = p0 + lane; p < p1; p += LPR) s += values[p] * __ldg(x + col[p]);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (active && lane == 0) y[row] = SUB ? y[row] - s : s;
}

__global__ void __launch_bounds__(256)
k_dirichlet(u64 n, const u64* __restrict__ rowPtr, const uint32_t* __restrict__ col, double* __restrict__ values,
            const uint8_t* __restrict__ dirichlet, const double* __restrict__ x, double* __restrict__ rhs) {
  const u64 row = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const u64 p0 = rowPtr[row], p1 = rowPtr[row + 1];
  if (dirichlet[row]) {
    if (lane == 0) rhs[row] = x[row];
    for (u64 p = p0 + lane; p < p1; p += 32) values[p] = col[p] == row ? 1.0 : 0.0;
  } else {
    for (u64 p = p0 + lane; p < p1; p += 32) if (dirichlet[col[p]]) values[p] = 0.0;
  }
}

// ---- matrix-free Laplace ---------------------------------------------------------------------
template <int N1>
struct MfTables {
  double A[3][N1][N1];    // A[j][q][a] = l_a(x_q)            (values at the Gauss points)
  double D[3][N1][N1];    // D[j][q][a] = l_a'(x_q)
  double At[3][N1][N1];   // transposes: At[j][a][q]
  double Dt[3][N1][N1];
  double C[N1][N1];       // collocation derivative on the Gauss points: C[q][r] = l~_r'(x_q)
  double w[N1];           // 1-d Gauss weights on [0,1]
};

struct MfArgs {
  u64 ne, posA, posB;
  FastDiv divN0, divN1;
};

// One thread per element: y_loc = A^e x_loc, written as local[i * ne + e] (coalesced across the
// elements of a warp for every local index i).
template <int DIM, int K>
__global__ void __launch_bounds__(64)
k_laplace_local(const __grid_constant__ DevGrid g, const __grid_constant__ DevLayout L,
                const __grid_constant__ MfTables<K + 1> T, const __grid_constant__ MfArgs R,
                const double* __restrict__ x, double* __restrict__ local) {
  constexpr int N1 = K + 1;
  constexpr int NLOC = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  const u64 el = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= R.ne) return;
  int ec[3];
  {
    unsigned q, r;
    R.divN0.divmod((unsigned)el, q, r);
    ec[0] = (int)r;
    unsigned q2;
    R.divN1.divmod(q, q2, r);
    ec[1] = (int)r; ec[2] = (int)q2;
  }
  double c[NLOC];
#pragma unroll
  for (int a2 = 0; a2 < (DIM == 3 ? N1 : 1); ++a2)
#pragma unroll
    for (int a1 = 0; a1 < N1; ++a1)
#pragma unroll
      for (int a0 = 0; a0 < N1; ++a0) {
        const int s = ((a0 > 0 && a0 < K) ? 1 : 0) | ((a1 > 0 && a1 < K) ? 2 : 0) | ((DIM == 3 && a2 > 0 && a2 < K) ? 4 : 0);
        const int c0 = ec[0] + (a0 == K ? 1 : 0);
        const int c1 = ec[1] + (a1 == K ? 1 : 0);
        const int c2 = ec[2] + ((DIM == 3 && a2 == K) ? 1 : 0);
        const unsigned ent = (unsigned)c0 + (unsigned)L.csize[s][0] * ((unsigned)c1 + (unsigned)L.csize[s][1] * (unsigned)c2);
        const u64 p = R.posA * (L.compOffset[s] + (u64)ent) + R.posB;       // K <= 2: one node per entity
        c[a0 + N1 * (a1 + N1 * a2)] = __ldg(x + p);
      }
  // |det J| (J^-1_dd)^2 of this element
  double coef[3] = {0, 0, 0};
  {
    double detJ = 1.0, ji[3] = {1, 1, 1};
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
      const double l0 = gridCoordinate(g, j, g.off[j] + ec[j]), u0 = gridCoordinate(g, j, g.off[j] + ec[j] + 1);
      detJ *= u0 - l0; ji[j] = 1.0 / (u0 - l0);
    }
#pragma unroll
    for (int j = 0; j < DIM; ++j) coef[j] = detJ * ji[j] * ji[j];
  }
  // u = values at the (K+1)^DIM Gauss points; on that grid u is a polynomial of degree K per
  // direction, so d/dx_d at the points is the collocation derivative C along the grid lines:
  //   y = A^T [ sum_d C_d^T ( coef_d w (C_d u) ) ]      (4 tensor contractions + 2 DIM line passes
  // instead of 2 DIM tensor contractions: 972 instead of 1458 FMAs for Q2 in 3-d)
  double u[NLOC], z[NLOC];
  sumFactor<DIM, N1, N1>(T.A[0], T.A[1], T.A[2], c, u);
#pragma unroll
  for (int i = 0; i < NLOC; ++i) z[i] = 0.0;
  constexpr int NZ = DIM == 3 ? N1 : 1;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const int st = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);          // stride of direction d in the point index
#pragma unroll
    for (int l2 = 0; l2 < NZ; ++l2)
#pragma unroll
      for (int l1 = 0; l1 < N1; ++l1) {
        // the line through (l1, l2) in the two other directions
        const int o1 = d == 0 ? N1 : 1, o2 = d == 2 ? N1 : N1 * N1;
        const int base = l1 * o1 + (DIM == 3 ? l2 * o2 : 0);
        const double wl = coef[d] * T.w[l1] * (DIM == 3 ? T.w[l2] : 1.0);
        double g[N1];
#pragma unroll
        for (int q = 0; q < N1; ++q) {
          double sgr = T.C[q][0] * u[base];
#pragma unroll
          for (int r = 1; r < N1; ++r) sgr = fma(T.C[q][r], u[base + r * st], sgr);
          g[q] = sgr * (wl * T.w[q]);
        }
#pragma unroll
        for (int r = 0; r < N1; ++r) {
          double acc = z[base + r * st];
#pragma unroll
          for (int q = 0; q < N1; ++q) acc = fma(T.C[q][r], g[q], acc);
          z[base + r * st] = acc;
        }
      }
  }
  double y[NLOC];
  sumFactor<DIM, N1, N1>(T.At[0], T.At[1], T.At[2], z, y);
#pragma unroll
  for (int i = 0; i < NLOC; ++i) local[(u64)i * R.ne + el] = y[i];
}

// One thread per DOF of ONE index-set component (all per-component constants are kernel
// arguments, 32-bit magic divisions only): sum of the element-local results of the elements that
// contain the node, ascending element order; Dirichlet rows return x (identity rows of the
// modified matrix).  Threads of a warp = consecutive entities in x: every read of `local` is
// one contiguous run per (element offset, local index).
struct GatherArgs {
  u64 ne, posA, posB, compOffset;
  unsigned count;              // DOFs of this component
  int s;                       // shift mask: bit j set = the entity extends along j (node inside the element)
  int k, dim;
  int N[3];
  FastDiv divC0, divC1;        // entity coordinates from the in-component index
};

template <int DIM, int K, int S>
__global__ void __launch_bounds__(256)
k_laplace_gather(const __grid_constant__ GatherArgs G, const double* __restrict__ local, const double* __restrict__ x,
                 const uint8_t* __restrict__ mask, double* __restrict__ y) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G.count) return;
  const u64 p = G.compOffset + (u64)t;               // flat scalar basis: position = scalar index
  if (mask != nullptr && mask[p]) { y[p] = x[p]; return; }
  unsigned c0, c1, c2, q;
  G.divC0.divmod(t, q, c0);
  G.divC1.divmod(q, c2, c1);
  constexpr int k1 = K + 1;
  constexpr bool P0 = !(S & 1), P1 = !((S >> 1) & 1), P2 = DIM == 3 && !((S >> 2) & 1);   // pinned directions
  const unsigned N0 = (unsigned)G.N[0], N1 = (unsigned)G.N[1], N2 = (unsigned)G.N[2];
  double s = 0.0;
  // elements e = c - delta, delta_j in {0,1} in the pinned directions; delta = 1 is the lower
  // element (ascending element order = delta descending); local indices are compile-time
#pragma unroll
  for (int d2 = P2 ? 1 : 0; d2 >= 0; --d2)
#pragma unroll
    for (int d1 = P1 ? 1 : 0; d1 >= 0; --d1)
#pragma unroll
      for (int d0 = P0 ? 1 : 0; d0 >= 0; --d0) {
        const int a0 = P0 ? (d0 ? K : 0) : 1, a1 = P1 ? (d1 ? K : 0) : 1, a2 = DIM == 3 ? (P2 ? (d2 ? K : 0) : 1) : 0;
        const int iLoc = a0 + k1 * (a1 + k1 * a2);
        const bool ok = (!P0 || (d0 ? c0 > 0 : c0 < N0)) && (!P1 || (d1 ? c1 > 0 : c1 < N1)) && (!P2 || (d2 ? c2 > 0 : c2 < N2));
        if (ok) {
          const unsigned ei = (c0 - d0) + N0 * ((c1 - d1) + N1 * (c2 - d2));
          s += local[(u64)iLoc * G.ne + ei];
        }
      }
  y[p] = s;
}

// x with the Dirichlet entries zeroed (the columns the modified matrix drops)
__global__ void __launch_bounds__(256)
k_mask_copy(u64 n, const double* __restrict__ x, const uint8_t* __restrict__ mask, double* __restrict__ out) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = mask[t] ? 0.0 : x[t];
}

// ---- host: Gauss rule and reference-element tables --------------------------------------------
void gaussLegendre01(int m, std::vector<double>& x, std::vector<double>& w) {
  x.resize(m); w.resize(m);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < m; ++i) {
    long double z = cosl(pi * (i + 0.75L) / (m + 0.5L)), pp = 1;
    for (int it = 0; it < 100; ++it) {
      long double p1 = 1, p2 = 0;
      for (int j = 1; j <= m; ++j) { const long double p3 = p2; p2 = p1; p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j; }
      pp = m * (z * p1 - p2) / (z * z - 1);
      const long double dz = p1 / pp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    x[m - 1 - i] = (double)(0.5L * (1 + z));
    w[m - 1 - i] = (double)(1 / ((1 - z * z) * pp * pp));
  }
}

// tensor rule of `order` on the reference cube, last coordinate fastest (dune-geometry)
void tensorRule(int dim, int order, std::vector<double>& pos, std::vector<double>& wt) {
  const int m = (order < 0 ? 0 : order) / 2 + 1;
  std::vector<double> x, w;
  gaussLegendre01(m, x, w);
  int n = 1;
  for (int j = 0; j < dim; ++j) n *= m;
  pos.resize((size_t)n * dim); wt.resize(n);
  for (int p = 0; p < n; ++p) {
    int r = p; double ww = 1;
    for (int j = dim - 1; j >= 0; --j) { const int i = r % m; r /= m; pos[(size_t)p * dim + j] = x[i]; ww *= w[i]; }
    wt[p] = ww;
  }
}

void shapeAt(int dim, int k, const double* x, std::vector<double>& val, std::vector<double>& grad) {
  int n = 1;
  for (int j = 0; j < dim; ++j) n *= k + 1;
  val.assign(n, 1.0); grad.assign((size_t)n * dim, 1.0);
  for (int i = 0; i < n; ++i) {
    int a[3] = {0, 0, 0};
    { int ii = i; for (int j = 0; j < dim; ++j) { a[j] = ii % (k + 1); ii /= (k + 1); } }
    auto P = [&](int j) { return k == 0 ? 1.0 : (k == 1 ? (a[j] ? x[j] : 1 - x[j]) : lagrangeP(k, a[j], x[j])); };
    auto DP = [&](int j) { return k == 0 ? 0.0 : (k == 1 ? (a[j] ? 1.0 : -1.0) : lagrangeDP(k, a[j], x[j])); };
    for (int j = 0; j < dim; ++j) val[i] *= P(j);
    for (int d = 0; d < dim; ++d)
      for (int j = 0; j < dim; ++j) grad[(size_t)i * dim + d] *= j == d ? DP(j) : P(j);
  }
}

}  // namespace

int matrixPattern(dfx_basis* b, u64* rowPtr, uint32_t* col, u64* nnzHost, cudaStream_t st) {
  const DevBasis& D = b->dev;
  if (b->dimension >= 0xffffffffull) return fail(DFX_ERR_RANGE, "column indices exceed 32 bits");
  // longest possible row: every leaf's DOFs of a 2^dim element box
  {
    u64 longest = 0;
    for (int l = 0; l < D.nLeaves; ++l) {
      const DevLayout& L = D.layouts[D.leaves[l].layout];
      u64 n = 1;
      const bool cell = L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0;
      for (int j = 0; j < D.grid.dim; ++j) n *= cell ? 2 : (u64)(2 * L.k + 1);
      longest += cell ? n * (u64)L.nloc : n;
    }
    if (longest > (u64)kMaxRowLen) return fail(DFX_ERR_NOT_IMPLEMENTED, "matrix rows longer than the staging buffer (order too high)");
  }
  const u64 n = b->dimension;
  if (!col) {
    // row lengths, then an exclusive scan in place; rowPtr[n] = nnz
    DFX_CUDA(cudaMemsetAsync(rowPtr + n, 0, sizeof(u64), st));
    for (int q = 0; q < D.nLayouts; ++q) {
      const u64 rows = D.layouts[q].dimension;
      const u64 blocks = (rows + kRowWarps - 1) / kRowWarps;
      if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "basis too large for one launch");
      k_pattern<false><<<(unsigned)blocks, kRowWarps * 32, 0, st>>>((const DevBasis*)b->dDev, q, rowPtr, nullptr, nullptr);
      int rc = postLaunch("k_pattern");
      if (rc) return rc;
    }
    size_t tmpBytes = 0;
    if (n + 1 > 0x7fffffffull) return fail(DFX_ERR_RANGE, "too many rows for the scan");
    DFX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, rowPtr, rowPtr, (int)(n + 1), st));
    void* tmp = nullptr;
    DFX_CUDA(cudaMallocAsync(&tmp, tmpBytes, st));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, rowPtr, rowPtr, (int)(n + 1), st);
    cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) return cudaFail(e, "cub::DeviceScan::ExclusiveSum");
    g_launchCount.fetch_add(1, std::memory_order_relaxed);
  } else {
    for (int q = 0; q < D.nLayouts; ++q) {
      const u64 rows = D.layouts[q].dimension;
      const u64 blocks = (rows + kRowWarps - 1) / kRowWarps;
      k_pattern<true><<<(unsigned)blocks, kRowWarps * 32, 0, st>>>((const DevBasis*)b->dDev, q, nullptr, rowPtr, col);
      int rc = postLaunch("k_pattern");
      if (rc) return rc;
    }
  }
  if (nnzHost) {
    DFX_CUDA(cudaMemcpyAsync(nnzHost, rowPtr + n, sizeof(u64), cudaMemcpyDeviceToHost, st));
    DFX_CUDA(cudaStreamSynchronize(st));
  }
  return DFX_OK;
}

int assembleLaplace(dfx_basis* b, const u64* rowPtr, const uint32_t* col, double* values, const dfx_function* f,
                    double* rhs, cudaStream_t st) {
  const DevBasis& D = b->dev;
  if (D.nLeaves != 1) return fail(DFX_ERR_NOT_IMPLEMENTED, "the Laplace assembler is for a single-leaf (scalar) basis");
  const DevLayout& L = D.layouts[D.leaves[0].layout];
  const int dim = D.grid.dim, k = L.k, n = L.nloc;
  {
    u64 longest = 1;
    const bool cell = L.kind == DFX_NODE_LAGRANGE_DG || k == 0;
    for (int j = 0; j < dim; ++j) longest *= cell ? (u64)(k + 1) : (u64)(2 * k + 1);
    if (longest > (u64)kMaxLaplaceRow) return fail(DFX_ERR_NOT_IMPLEMENTED, "matrix rows longer than the staging buffer (order too high)");
  }
  LaplaceArgs A;
  A.nloc = n; A.posA = D.leaves[0].posA; A.posB = D.leaves[0].posB; A.nq = 0;
  const u64 rows = L.dimension;
  const u64 blocks = (rows + kRowWarps - 1) / kRowWarps;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "basis too large for one launch");
  std::vector<double> pos, wt, val, grad;
  if (values) {
    // K[d][i][j] with the example's rule: order 2 (dim * order - 1)
    tensorRule(dim, 2 * (dim * k - 1), pos, wt);
    std::vector<double> K((size_t)dim * n * n, 0.0);
    for (size_t q = 0; q < wt.size(); ++q) {
      shapeAt(dim, k, &pos[q * dim], val, grad);
      for (int d = 0; d < dim; ++d)
        for (int i = 0; i < n; ++i) {
          const double gi = grad[(size_t)i * dim + d] * wt[q];
          for (int j = 0; j < n; ++j) K[((size_t)d * n + i) * n + j] += gi * grad[(size_t)j * dim + d];
        }
    }
    double* dK = nullptr;
    DFX_CUDA(cudaMallocAsync(&dK, K.size() * sizeof(double), st));
    DFX_CUDA(cudaMemcpyAsync(dK, K.data(), K.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    DFX_CUDA(cudaStreamSynchronize(st));       // K leaves scope
    k_laplace_rows<<<(unsigned)blocks, kRowWarps * 32, 0, st>>>((const DevBasis*)b->dDev, A, dK, rowPtr, col, values);
    int rc = postLaunch("k_laplace_rows");
    cudaFreeAsync(dK, st);
    if (rc) return rc;
  }
  if (rhs) {
    if (!f) return fail(DFX_ERR_INVALID, "null volume term");
    if (f->n_comp != 1) return fail(DFX_ERR_RANGE, "the volume term must be scalar");
    tensorRule(dim, dim * k, pos, wt);
    const int nq = (int)wt.size();
    A.nq = nq;
    std::vector<double> phi((size_t)nq * n);
    for (int q = 0; q < nq; ++q) {
      shapeAt(dim, k, &pos[(size_t)q * dim], val, grad);
      for (int i = 0; i < n; ++i) phi[(size_t)q * n + i] = val[i];
    }
    const u64 ne = (u64)D.grid.N[0] * D.grid.N[1] * D.grid.N[2];
    const u64 total = ne * (u64)nq;
    double *dTab = nullptr, *dPts = nullptr;
    const size_t tabN = (size_t)nq * dim + nq + phi.size();
    DFX_CUDA(cudaMallocAsync(&dTab, tabN * sizeof(double), st));
    DFX_CUDA(cudaMallocAsync(&dPts, total * sizeof(double), st));
    std::vector<double> tab;
    tab.insert(tab.end(), pos.begin(), pos.end());
    tab.insert(tab.end(), wt.begin(), wt.end());
    tab.insert(tab.end(), phi.begin(), phi.end());
    DFX_CUDA(cudaMemcpyAsync(dTab, tab.data(), tabN * sizeof(double), cudaMemcpyHostToDevice, st));
    DFX_CUDA(cudaStreamSynchronize(st));
    RegistryFn fn; fn.f = *f; fn.dim = dim;
    const u64 pb = (total + 255) / 256;
    if (pb > 0x7fffffffull) return fail(DFX_ERR_RANGE, "grid too large for one launch");
    k_volume_points<RegistryFn><<<(unsigned)pb, 256, 0, st>>>(D.grid, nq, dTab, dTab + (size_t)nq * dim, fn, total, dPts);
    int rc = postLaunch("k_volume_points");
    if (!rc) {
      k_volume_rows<<<(unsigned)blocks, kRowWarps * 32, 0, st>>>((const DevBasis*)b->dDev, A, dTab + (size_t)nq * dim + nq, dPts, rhs);
      rc = postLaunch("k_volume_rows");
    }
    cudaFreeAsync(dTab, st); cudaFreeAsync(dPts, st);
    if (rc) return rc;
  }
  return DFX_OK;
}

template <int DIM, int K>
static int launchLaplaceLocal(dfx_basis* b, const MfArgs& R, const double* x, double* local, cudaStream_t st) {
  constexpr int N1 = K + 1;
  MfTables<N1> T;
  std::vector<double> gx, gw;
  gaussLegendre01(N1, gx, gw);
  for (int j = 0; j < 3; ++j)
    for (int q = 0; q < N1; ++q)
      for (int a = 0; a < N1; ++a) {
        const double v = K == 1 ? (a ? gx[q] : 1 - gx[q]) : lagrangeP(K, a, gx[q]);
        const double d = K == 1 ? (a ? 1.0 : -1.0) : lagrangeDP(K, a, gx[q]);
        T.A[j][q][a] = v; T.At[j][a][q] = v; T.D[j][q][a] = d; T.Dt[j][a][q] = d;
      }
  for (int q = 0; q < N1; ++q) T.w[q] = gw[q];
  for (int q = 0; q < N1; ++q) {
    double diag = 0.0;
    for (int r = 0; r < N1; ++r) {
      if (r == q) continue;
      double wq = 1.0, wr = 1.0;
      for (int m = 0; m < N1; ++m) { if (m != q) wq *= gx[q] - gx[m]; if (m != r) wr *= gx[r] - gx[m]; }
      T.C[q][r] = (wq / wr) / (gx[q] - gx[r]);
      diag += 1.0 / (gx[q] - gx[r]);
    }
    T.C[q][q] = diag;
  }
  const u64 blocks = (R.ne + 63) / 64;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "grid too large for one launch");
  k_laplace_local<DIM, K><<<(unsigned)blocks, 64, 0, st>>>(b->dev.grid, b->dev.layouts[b->dev.leaves[0].layout], T, R, x, local);
  return postLaunch("k_laplace_local");
}

// y = A x with the Laplace stiffness matrix of assembleLaplace, never assembled.  `scratch` is
// caller-provided: n_loc * numElements doubles for the element-local results plus dimension()
// doubles for the masked copy of x.
int applyLaplace(dfx_basis* b, const double* x, double* y, const uint8_t* dirichlet, double* scratch, cudaStream_t st) {
  const DevBasis& D = b->dev;
  if (D.nLeaves != 1) return fail(DFX_ERR_NOT_IMPLEMENTED, "the matrix-free Laplace operator is for a single-leaf (scalar) basis");
  const DevLayout& L = D.layouts[D.leaves[0].layout];
  if (L.kind != DFX_NODE_LAGRANGE || L.k < 1 || L.k > 2 || D.grid.dim < 2)
    return fail(DFX_ERR_NOT_IMPLEMENTED, "matrix-free Laplace: continuous Lagrange order 1 or 2 in 2-d / 3-d");
  if (D.leaves[0].posA != 1 || D.leaves[0].posB != 0) return fail(DFX_ERR_NOT_IMPLEMENTED, "matrix-free Laplace: flat scalar basis");
  MfArgs R;
  R.ne = (u64)D.grid.N[0] * D.grid.N[1] * D.grid.N[2];
  if (R.ne > 0xffffffffull || L.dimension > 0xffffffffull) return fail(DFX_ERR_RANGE, "more than 2^32 elements in one box");
  R.posA = 1; R.posB = 0;
  R.divN0 = FastDiv((unsigned)D.grid.N[0]); R.divN1 = FastDiv((unsigned)D.grid.N[1]);
  double* local = scratch;
  const double* xin = x;
  if (dirichlet) {
    double* xm = scratch + (size_t)L.nloc * R.ne;
    const u64 n = L.dimension;
    k_mask_copy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, x, dirichlet, xm);
    int rc0 = postLaunch("k_mask_copy");
    if (rc0) return rc0;
    xin = xm;
  }
  int rc = DFX_ERR_NOT_IMPLEMENTED;
  if (D.grid.dim == 3 && L.k == 2) rc = launchLaplaceLocal<3, 2>(b, R, xin, local, st);
  if (D.grid.dim == 3 && L.k == 1) rc = launchLaplaceLocal<3, 1>(b, R, xin, local, st);
  if (D.grid.dim == 2 && L.k == 2) rc = launchLaplaceLocal<2, 2>(b, R, xin, local, st);
  if (D.grid.dim == 2 && L.k == 1) rc = launchLaplaceLocal<2, 1>(b, R, xin, local, st);
  if (rc) return rc;
  for (int q = 0; q < L.ncomp; ++q) {
    const int s = L.order[q];
    GatherArgs G;
    G.ne = R.ne; G.posA = 1; G.posB = 0; G.compOffset = L.compOffset[s];
    G.count = (unsigned)L.compCount[s];
    G.s = s; G.k = L.k; G.dim = D.grid.dim;
    for (int j = 0; j < 3; ++j) G.N[j] = D.grid.N[j];
    G.divC0 = FastDiv((unsigned)L.csize[s][0]); G.divC1 = FastDiv((unsigned)L.csize[s][1]);
    if (G.count == 0) continue;
    const unsigned gb = (G.count + 255) / 256;
    bool launched = false;
#define DFX_GATHER(DIM_, K_, S_)                                                            \
  if (!launched && D.grid.dim == DIM_ && L.k == K_ && s == S_) {                            \
    k_laplace_gather<DIM_, K_, S_><<<gb, 256, 0, st>>>(G, local, x, dirichlet, y);          \
    launched = true;                                                                        \
  }
    DFX_GATHER(3, 2, 0) DFX_GATHER(3, 2, 1) DFX_GATHER(3, 2, 2) DFX_GATHER(3, 2, 3)
    DFX_GATHER(3, 2, 4) DFX_GATHER(3, 2, 5) DFX_GATHER(3, 2, 6) DFX_GATHER(3, 2, 7)
    DFX_GATHER(2, 2, 0) DFX_GATHER(2, 2, 1) DFX_GATHER(2, 2, 2) DFX_GATHER(2, 2, 3)
    DFX_GATHER(3, 1, 0) DFX_GATHER(2, 1, 0)
#undef DFX_GATHER
    if (!launched) return fail(DFX_ERR_NOT_IMPLEMENTED, "matrix-free Laplace: unexpected index-set component");
    rc = postLaunch("k_laplace_gather");
    if (rc) return rc;
  }
  return DFX_OK;
}

int assembleStokes(dfx_basis* b, const u64* rowPtr, const uint32_t* col, double* values, cudaStream_t st) {
  const DevBasis& D = b->dev;
  const int dim = D.grid.dim;
  bool th = D.nLeaves == dim + 1 && dim >= 2;
  for (int l = 0; th && l < dim; ++l) {
    const DevLayout& L = D.layouts[D.leaves[l].layout];
    th = L.kind == DFX_NODE_LAGRANGE && L.k == 2 && D.leaves[l].layout == D.leaves[0].layout;
  }
  if (th) { const DevLayout& L = D.layouts[D.leaves[dim].layout]; th = L.kind == DFX_NODE_LAGRANGE && L.k == 1; }
  if (!th) return fail(DFX_ERR_NOT_IMPLEMENTED, "the Stokes assembler needs a Taylor-Hood tree composite(power<dim>(lagrange<2>), lagrange<1>)");
  const DevLayout& LV = D.layouts[D.leaves[0].layout];
  const DevLayout& LP = D.layouts[D.leaves[dim].layout];
  const int nv = LV.nloc, np = LP.nloc;
  {
    u64 longest = 0, v = 1, p = 1;
    for (int j = 0; j < dim; ++j) { v *= 5; p *= 3; }
    longest = (u64)dim * v + p;
    if (longest > (u64)kMaxLaplaceRow) return fail(DFX_ERR_NOT_IMPLEMENTED, "matrix rows longer than the staging buffer");
  }
  std::vector<double> pos, wt, valV, gradV, valP, gradP;
  tensorRule(dim, 2 * (dim * 2 - 1), pos, wt);
  std::vector<double> K((size_t)dim * nv * nv, 0.0), G((size_t)dim * nv * np, 0.0);
  for (size_t q = 0; q < wt.size(); ++q) {
    shapeAt(dim, 2, &pos[q * dim], valV, gradV);
    shapeAt(dim, 1, &pos[q * dim], valP, gradP);
    for (int d = 0; d < dim; ++d)
      for (int i = 0; i < nv; ++i) {
        const double gi = gradV[(size_t)i * dim + d] * wt[q];
        for (int j = 0; j < nv; ++j) K[((size_t)d * nv + i) * nv + j] += gi * gradV[(size_t)j * dim + d];
        for (int j = 0; j < np; ++j) G[((size_t)d * nv + i) * np + j] += gi * valP[j];
      }
  }
  double* dTab = nullptr;
  DFX_CUDA(cudaMallocAsync(&dTab, (K.size() + G.size()) * sizeof(double), st));
  DFX_CUDA(cudaMemcpyAsync(dTab, K.data(), K.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  DFX_CUDA(cudaMemcpyAsync(dTab + K.size(), G.data(), G.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  DFX_CUDA(cudaStreamSynchronize(st));
  StokesArgs A; A.dim = dim; A.nv = nv; A.np = np;
  int rc = DFX_OK;
  for (int l = 0; l <= dim && !rc; ++l) {
    const u64 rows = D.layouts[D.leaves[l].layout].dimension;
    const u64 blocks = (rows + kRowWarps - 1) / kRowWarps;
    if (blocks > 0x7fffffffull) { rc = fail(DFX_ERR_RANGE, "basis too large for one launch"); break; }
    k_stokes_rows<<<(unsigned)blocks, kRowWarps * 32, 0, st>>>((const DevBasis*)b->dDev, A, l, dTab, dTab + K.size(), rowPtr, col, values);
    rc = postLaunch("k_stokes_rows");
  }
  cudaFreeAsync(dTab, st);
  return rc;
}

int csrMv(u64 n, const u64* rowPtr, const uint32_t* col, const double* values, const double* x, double* y, bool subtract,
          cudaStream_t st, int meanRow) {
  if (n == 0) return DFX_OK;
  const int lpr = meanRow > 48 ? 32 : (meanRow > 24 ? 16 : (meanRow > 12 ? 8 : 4));
  const u64 blocks = (n * (u64)lpr + 255) / 256;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "matrix too large for one launch");
#define DFX_MV(LPR_)                                                                                   \
  if (lpr == LPR_) {                                                                                   \
    if (subtract) k_csr_mv<true, LPR_><<<(unsigned)blocks, 256, 0, st>>>(n, rowPtr, col, values, x, y); \
    else k_csr_mv<false, LPR_><<<(unsigned)blocks, 256, 0, st>>>(n, rowPtr, col, values, x, y);         \
  }
  DFX_MV(4) DFX_MV(8) DFX_MV(16) DFX_MV(32)
#undef DFX_MV
  return postLaunch("k_csr_mv");
}

int dirichletRows(u64 n, const u64* rowPtr, const uint32_t* col, double* values, const uint8_t* dirichlet, const double* x,
                  double* rhs, cudaStream_t st, int meanRow) {
  if (n == 0) return DFX_OK;
  int rc = csrMv(n, rowPtr, col, values, x, rhs, true, st, meanRow);       // stiffnessMatrix.mmv(x, rhs)
  if (rc) return rc;
  const u64 blocks = (n * 32 + 255) / 256;
  k_dirichlet<<<(unsigned)blocks, 256, 0, st>>>(n, rowPtr, col, values, dirichlet, x, rhs);
  return postLaunch("k_dirichlet");
}

}  // namespace dfx
