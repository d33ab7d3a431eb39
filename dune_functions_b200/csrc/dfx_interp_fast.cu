// dfx_interp_fast.cu — structured fast paths for interpolate() and the index table.
//
//   k_interp_rows      owner-writes interpolation of continuous Lagrange leaves.  The DOF
//                      lattice has (k N_x + 1) x (k N_y + 1) x (k N_z + 1) points; a block takes
//                      32 lattice rows (Y, Z fixed per row) and first puts the per-row, per-leaf
//                      flat-position bases into shared memory.  Each thread then owns fixed
//                      x-entities (their x-direction factors / coordinates stay in registers)
//                      and sweeps the rows: per row it writes the vertex-type node and the k-1
//                      edge-type nodes right of it, so consecutive threads write consecutive
//                      addresses in both index-set components the row touches.
//   k_interp_cell      element-local DOFs (LagrangeDG, Lagrange<0>): a warp sweeps consecutive
//                      elements, each lane keeps fixed local indices; contiguous 256-byte stores.
//   k_sep_tables       per-direction factor tables of a separable f (exp(-a|x|^2) =
//                      prod_j exp(-a x_j^2), prod_j sin(w x_j), x_j, const): O(N) instead of
//                      O(N^3) transcendental evaluations.
//   k_indices_fast     element -> global index table with 32-bit magic-number division
//                      and a per-local-index lookup table.
//
// Replaces the interpolate element loop (interpolate.hh:254-259 with :151-173) and
// DefaultLocalView::bind -> PreBasis::indices (defaultlocalview.hh:95-101).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "dfx_device.cuh"
#include "dfx_eval_fast.h"
#include "dfx_launch.h"

namespace dfx {

struct RowArgs {
  int layoutId, firstLeaf, nLeaves;
  int LX[3];               // table length per direction
  int tabOff[3];           // separable mode: offset of direction j's table
  FastDiv divK;            // lattice coordinate -> (entity, node inside)
  FastDiv divLY;           // row index -> (Z, Y)
  int rowsPerBlock;        // lattice rows per block (<= kRowsPerBlock)
  int tabN;                // table length of one range entry (separable mode)
  int vectorTables;        // separable mode: one table set per range entry (vector-valued f)
  int interleaved;         // the active leaves are the components of an interleaved power node (position = C j + c)
  int nActive;             // leaves of the subtree that live on this lattice
  FastDiv divC;            // lean kernel: thread -> (x-entity, leaf) for an interleaved group of nActive leaves
  int leafId[DFX_MAX_LEAVES];     // their ids, range entries and in-entity position stride
  int leafComp[DFX_MAX_LEAVES];
  unsigned leafPosA[DFX_MAX_LEAVES];
};

// one direction of a lattice point: entity coordinate, in-entity node, owner element, alpha
struct Dir { int c, inner, ie, a; };

__device__ __forceinline__ Dir latticeDir(const DevGrid& g, const FastDiv& divK, int k, int j, int X) {
  Dir d;
  unsigned q, r;
  divK.divmod((unsigned)X, q, r);
  d.c = (int)q;
  d.inner = (int)r;                       // 0: on a lattice plane, else node r of the entity
  if (r != 0) { d.ie = g.off[j] + d.c; d.a = d.inner; }
  else {
    // shared by the elements on both sides: the later one in iteration order (the upper
    // neighbour, if the grid has one) writes last — interpolate.hh:254-259
    const int gc = g.off[j] + d.c;
    if (gc < g.G[j]) { d.ie = gc; d.a = 0; } else { d.ie = g.G[j] - 1; d.a = k; }
  }
  return d;
}

__device__ __forceinline__ double nodeCoord(const DevGrid& g, int k, int j, int ie, int a) {
  const double xhat = k == 0 ? 0.5 : __ddiv_rn(__dmul_rn(1.0, (double)a), (double)k);
  return elementGlobal(g, j, ie, xhat);
}

// factor of range entry c of a separable registry function in direction j:
// y_c(x) = prod_j sepFactor(f, c, j, x_j)
__device__ __forceinline__ double sepFactor(const dfx_function& f, int c, int j, double x) {
  switch (f.id) {
    case DFX_FN_GAUSSIAN: return exp(__dmul_rn(-f.p[0], __dmul_rn(x, x)));
    case DFX_FN_SINPROD: return sin(__dmul_rn(f.p[0], x));
    case DFX_FN_COORD: return j == (int)f.p[0] ? x : 1.0;
    case DFX_FN_CONSTANT: return j == 0 ? f.p[c] : 1.0;
    case DFX_FN_IDENTITY: return j == c ? x : 1.0;                     // y_c = x_c (0 for c >= dim, see the table kernel)
    case DFX_FN_GAUSSIAN_VEC: {                                         // y_c = x_c exp(-a |x|^2)
      const double e = exp(__dmul_rn(-f.p[0], __dmul_rn(x, x)));
      return j == c ? __dmul_rn(x, e) : e; }
  }
  return 0.0;
}

__global__ void __launch_bounds__(256)
k_sep_tables(const __grid_constant__ DevGrid g, const __grid_constant__ dfx_function f, const __grid_constant__ RowArgs A,
             int k, int cellMode, double* __restrict__ tab) {
  const int tt = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = A.tabOff[2] + A.LX[2];                 // table length of one range entry
  if (tt >= n * f.n_comp) return;
  const int c = tt / n, t = tt - c * n;
  const int j = t >= A.tabOff[2] ? 2 : (t >= A.tabOff[1] ? 1 : 0);
  if (j >= g.dim) { tab[tt] = 1.0; return; }
  // IDENTITY has no coordinate for range entries beyond the dimension: they are 0
  if (f.id == DFX_FN_IDENTITY && c >= g.dim) { tab[tt] = 0.0; return; }
  const int X = t - A.tabOff[j];
  int ie, a;
  if (cellMode) { ie = g.off[j] + X / (k + 1); a = X % (k + 1); }
  else { const Dir d = latticeDir(g, A.divK, k, j, X); ie = d.ie; a = d.a; }
  tab[tt] = sepFactor(f, c, j, nodeCoord(g, k, j, ie, a));
}

// per lattice row (Y, Z fixed) and active leaf: flat position of x-entity 0's vertex-type
// node / first edge-type node and the position stride between consecutive x-entities
struct RowLeaf { u64 baseV, baseE; unsigned strideV, strideE, posA, comp; double yz; };   // yz: MODE 1, t_y t_z of the leaf's range entry
struct RowCommon { double u1, u2; };   // MODE 0: node coordinates x_1, x_2;  MODE 1: factors t_y, t_z
constexpr int kRowsPerBlock = 32;      // most lattice rows one block sweeps (A.rowsPerBlock <= this)
constexpr int kMaxHoist = 4;           // orders 1..4 keep the x-direction data in registers

// MODE 0: f evaluated at every node; MODE 1: product of per-direction tables.
// KT: compile-time order (0 = run-time order, nothing hoisted).  ONE: the subtree has a
// single leaf on this lattice (the scalar case), its bases sit in slot 0.
template <class F, int MODE, int KT, bool ONE>
__global__ void __launch_bounds__(256)
k_interp_rows(const __grid_constant__ DevBasis B, const __grid_constant__ RowArgs A, const F f,
              const double* __restrict__ tab, double* __restrict__ coeff, const uint8_t* __restrict__ mask) {
  __shared__ RowCommon rowc[kRowsPerBlock];
  __shared__ RowLeaf rowl[ONE ? 1 : DFX_MAX_LEAVES][kRowsPerBlock];
  const DevGrid& g = B.grid;
  const DevLayout& L = B.layouts[A.layoutId];
  const int k = KT > 0 ? KT : L.k;            // >= 1
  const unsigned nRows = (unsigned)A.LX[1] * (unsigned)A.LX[2];
  const unsigned row0 = blockIdx.x * (unsigned)A.rowsPerBlock;
  const unsigned nr = min((unsigned)A.rowsPerBlock, nRows - row0);
  const int nAct = ONE ? 1 : A.nActive;
  // MODE 1 with a single active leaf: the tables of that leaf's range entry
  const double* tab1 = (MODE == 1 && ONE && A.vectorTables) ? tab + (size_t)A.leafComp[0] * A.tabN : tab;
  // ---- per-row setup by the first nr*nAct threads ----
  for (unsigned w = threadIdx.x; w < nr * (unsigned)nAct; w += blockDim.x) {
    const unsigned rI = w / (unsigned)nAct, lI = w - rI * (unsigned)nAct;
    unsigned Z, Y;
    A.divLY.divmod(row0 + rI, Z, Y);
    const Dir dy = latticeDir(g, A.divK, k, 1, (int)Y);
    const Dir dz = latticeDir(g, A.divK, k, 2, (int)Z);
    unsigned syz = 0;
    u64 fyz = 0, fs = 1;
    if (g.dim > 1 && dy.inner) { syz |= 2u; fyz += (u64)(dy.inner - 1) * fs; fs *= (u64)(k - 1); }
    if (g.dim > 2 && dz.inner) { syz |= 4u; fyz += (u64)(dz.inner - 1) * fs; }
    const unsigned sV = syz, sE = syz | 1u;
    const u64 rowV = L.compOffset[sV] + (u64)L.blk[sV] * ((u64)L.csize[sV][0] * ((u64)dy.c + (u64)L.csize[sV][1] * (u64)dz.c)) + fyz;
    const u64 rowE = L.compOffset[sE] + (u64)L.blk[sE] * ((u64)L.csize[sE][0] * ((u64)dy.c + (u64)L.csize[sE][1] * (u64)dz.c)) + fyz * (u64)(k - 1);
    const DevLeaf& leaf = B.leaves[A.leafId[lI]];
    RowLeaf rl;
    rl.baseV = leaf.posA * rowV + leaf.posB;
    rl.baseE = leaf.posA * rowE + leaf.posB;
    rl.strideV = (unsigned)(leaf.posA * (u64)L.blk[sV]);
    rl.strideE = (unsigned)(leaf.posA * (u64)L.blk[sE]);
    rl.posA = (unsigned)leaf.posA; rl.comp = (unsigned)A.leafComp[lI];
    rl.yz = 0.0;
    if (MODE == 1) {
      const double* tc = tab + (size_t)(A.vectorTables ? A.leafComp[lI] : 0) * A.tabN;
      rl.yz = __dmul_rn(tc[A.tabOff[1] + (int)Y], tc[A.tabOff[2] + (int)Z]);
    }
    rowl[lI][rI] = rl;
    if (lI == 0) {
      RowCommon rc;
      if (MODE == 0) {
        rc.u1 = g.dim > 1 ? nodeCoord(g, k, 1, dy.ie, dy.a) : 0.0;
        rc.u2 = g.dim > 2 ? nodeCoord(g, k, 2, dz.ie, dz.a) : 0.0;
      } else {
        rc.u1 = __dmul_rn(tab1[A.tabOff[1] + (int)Y], tab1[A.tabOff[2] + (int)Z]);   // t_y * t_z
        rc.u2 = 0.0;
      }
      rowc[rI] = rc;
    }
  }
  __syncthreads();
  const bool scalar = MODE == 1 || f.ncomp() == 1;
  // x-direction datum of node a of x-entity c0: coordinate (MODE 0) or table factor (MODE 1)
  auto xdata = [&](unsigned c0, int a) -> double {
    if (MODE == 1) return tab1[A.tabOff[0] + k * (int)c0 + a];
    int ie = g.off[0] + (int)c0, aa = a;
    if (a == 0 && ie >= g.G[0]) { ie = g.G[0] - 1; aa = k; }    // last lattice plane: owner on the left
    return nodeCoord(g, k, 0, ie, aa);
  };
  // write node a of x-entity c0 in row rI for every active leaf
  auto emit = [&](unsigned c0, unsigned rI, int a, double xa, const RowCommon& rc) {
    const double cm = MODE == 0 ? f.common(xa, rc.u1, rc.u2) : __dmul_rn(xa, rc.u1);
    for (int lI = 0; lI < (ONE ? 1 : nAct); ++lI) {
      const RowLeaf rl = rowl[lI][rI];
      const u64 p = a == 0 ? rl.baseV + (u64)rl.strideV * (u64)c0
                           : rl.baseE + (u64)rl.strideE * (u64)c0 + (u64)rl.posA * (u64)(a - 1);
      if (mask == nullptr || mask[p]) {
        if (MODE == 1 && !ONE && A.vectorTables)        // range entry of this leaf: its own x-table and row factor
          coeff[p] = __dmul_rn(tab[(size_t)rl.comp * A.tabN + A.tabOff[0] + k * (int)c0 + a], rl.yz);
        else
          coeff[p] = MODE == 1 ? cm : f.comp(scalar ? 0 : (int)rl.comp, cm, xa, rc.u1, rc.u2);
      }
    }
  };
  // ---- table mode on an interleaved group (power node with FlatInterleaved / BlockedInterleaved
  // merging: position = C j + c): the leaves are spread over the threads, thread u owns (x-entity
  // u / C, leaf u % C), so the lanes of a warp write consecutive positions instead of C strided runs ----
  if (MODE == 1 && !ONE && A.interleaved) {
    const unsigned C = (unsigned)nAct;
    for (unsigned u = threadIdx.x; u < ((unsigned)g.N[0] + 1u) * C; u += blockDim.x) {
      const unsigned c0 = u / C, lI = u - c0 * C;
      const bool closing = c0 == (unsigned)g.N[0];                 // the last lattice plane: vertex-type node only
      const unsigned comp = (unsigned)(A.vectorTables ? A.leafComp[lI] : 0);
      const double* tx = tab + (size_t)comp * A.tabN + A.tabOff[0] + k * (int)c0;
      for (unsigned rI = 0; rI < nr; ++rI) {
        const RowLeaf rl = rowl[lI][rI];
        for (int a = 0; a < (closing ? 1 : k); ++a) {
          const u64 p = a == 0 ? rl.baseV + (u64)rl.strideV * (u64)c0
                               : rl.baseE + (u64)rl.strideE * (u64)c0 + (u64)rl.posA * (u64)(a - 1);
          if (mask == nullptr || mask[p]) coeff[p] = __dmul_rn(tx[a], rl.yz);
        }
      }
    }
    return;
  }
  // ---- each thread owns x-entities c0 = tid, tid + blockDim, ... < N_x and sweeps the rows ----
  for (unsigned c0 = threadIdx.x; c0 < (unsigned)g.N[0]; c0 += blockDim.x) {
    // nodes X = k c0 + a, a = 0 (vertex-type) .. k-1 (edge-type); their x-data stay in registers
    double hx[KT > 0 ? KT : 1];
    if (KT > 0) {
#pragma unroll
      for (int a = 0; a < KT; ++a) hx[a] = xdata(c0, a);
    }
#pragma unroll 4
    for (unsigned rI = 0; rI < nr; ++rI) {
      const RowCommon rc = rowc[rI];
      if (KT > 0) {
#pragma unroll
        for (int a = 0; a < KT; ++a) emit(c0, rI, a, hx[a], rc);
      } else {
        for (int a = 0; a < k; ++a) emit(c0, rI, a, xdata(c0, a), rc);
      }
    }
  }
  // ---- the closing vertex plane x-entity N_x: one node per row ----
  if (threadIdx.x < nr) {
    const unsigned rI = threadIdx.x;
    emit((unsigned)g.N[0], rI, 0, xdata((unsigned)g.N[0], 0), rowc[rI]);
  }
}

// ---- per-node evaluation, scalar case: ONE leaf on the lattice, no mask ------------------------------
// The same sweep as k_interp_rows<F, 0, KT, true> with everything that is not the function taken out of
// the node loop: a row's record holds the ADDRESS of x-entity 0's vertex-type / first edge-type node and
// the byte stride between x-entities, so a node's store address is one 32 x 32 + 64-bit multiply-add;
// reference coordinates alpha / k are compile-time constants; the functor sits in the constant bank.
// With exp from expTab a Gaussian node is 14 FP64 issues + about 12 others, below the HBM time of the
// 8 bytes it writes.
// alpha / k for alpha in 0 .. KT as the compiler folds it (IEEE division at compile time: the bits of __ddiv_rn)
template <int KT>
__device__ __forceinline__ double xhatOf(int a) {
  double r = 0.0;
#pragma unroll
  for (int i = 1; i <= KT; ++i) if (a == i) r = (double)i / (double)KT;
  return r;
}
struct LeanRow { double* pV; double* pE; unsigned sV8, sE8; double u1, u2; unsigned a8, pad; double q1, q2; };   // q1, q2: u1^2, u2^2 (F::kFromSquares)
int g_rowsLean = -1;    // lattice rows per block of the lean kernel; -1: leanRows() (tuning knob DFX_OP_TUNE_ROWS)

// Many rows per block amortise the per-thread x-data (measured 1 .. 32 rows: 256^3 0.43 -> 0.19 ms, 512^3 flat
// from 4 rows on), but a small lattice must still give every SM several waves of blocks (Taylor-Hood 128^3
// has 66 049 + 16 641 rows: 32 rows per block are 2.2 waves with a 0.2-full tail; measured 2 .. 32 rows:
// 0.107 / 0.090 (4 rows) / 0.097 / 0.124 ms): about 128 blocks per SM.
// ... and the rows the resident blocks write at one time should stay a compact band of every index-set component:
// at 512^3 (8.2 KB per lattice row) 32 rows per block put 310 MB in flight and the
// kernel takes 1.82 ms, 8 rows (78 MB) 1.71 ms, 4 rows 1.76 ms (`profiles/r02_interp_row_hoist_ab.jsonl`); at
// 256^3 the optimum is flat from 16 rows on.  rowBytes: bytes one lattice row writes.
static int leanRows(u64 nRows, u64 rowBytes) {
  if (g_rowsLean > 0) return std::min(g_rowsLean, kRowsPerBlock);
  const u64 bySize = std::max<u64>(4, std::min<u64>(kRowsPerBlock, nRows / (148 * 128)));
  const u64 byBand = std::max<u64>(4, (80ull << 20) / (148ull * 8 * std::max<u64>(1, rowBytes)));
  return (int)std::min(bySize, byBand);
}

// Body shared by the single and the two-part kernel.  The active leaves are either ONE leaf (any position
// stride) or the C components of an interleaved power node (position = C j + c, FlatInterleaved /
// BlockedInterleaved): thread u owns (x-entity u / C, leaf u % C), so consecutive lanes write consecutive
// addresses whatever C is.
template <class F, int KT>
__device__ __forceinline__ void leanRowsBody(const DevBasis& B, const RowArgs& A, const F& f, double* __restrict__ coeff,
                                             unsigned blockId, LeanRow* rows) {
  const DevGrid& g = B.grid;
  const DevLayout& L = B.layouts[A.layoutId];
  constexpr int k = KT;
  const unsigned nRows = (unsigned)A.LX[1] * (unsigned)A.LX[2];
  const unsigned row0 = blockId * (unsigned)A.rowsPerBlock;
  const unsigned nr = min((unsigned)A.rowsPerBlock, nRows - row0);
  if (threadIdx.x < nr) {
    const unsigned rI = threadIdx.x;
    unsigned Z, Y;
    A.divLY.divmod(row0 + rI, Z, Y);
    const Dir dy = latticeDir(g, A.divK, k, 1, (int)Y);
    const Dir dz = latticeDir(g, A.divK, k, 2, (int)Z);
    unsigned syz = 0;
    u64 fyz = 0, fs = 1;
    if (g.dim > 1 && dy.inner) { syz |= 2u; fyz += (u64)(dy.inner - 1) * fs; fs *= (u64)(k - 1); }
    if (g.dim > 2 && dz.inner) { syz |= 4u; fyz += (u64)(dz.inner - 1) * fs; }
    const unsigned sV = syz, sE = syz | 1u;
    const u64 rowV = L.compOffset[sV] + (u64)L.blk[sV] * ((u64)L.csize[sV][0] * ((u64)dy.c + (u64)L.csize[sV][1] * (u64)dz.c)) + fyz;
    const u64 rowE = L.compOffset[sE] + (u64)L.blk[sE] * ((u64)L.csize[sE][0] * ((u64)dy.c + (u64)L.csize[sE][1] * (u64)dz.c)) + fyz * (u64)(k - 1);
    const DevLeaf& leaf = B.leaves[A.leafId[0]];
    LeanRow r;
    r.pV = coeff + (leaf.posA * rowV + leaf.posB);
    r.pE = coeff + (leaf.posA * rowE + leaf.posB);
    r.sV8 = (unsigned)(leaf.posA * (u64)L.blk[sV]) * 8u;
    r.sE8 = (unsigned)(leaf.posA * (u64)L.blk[sE]) * 8u;
    r.a8 = (unsigned)leaf.posA * 8u; r.pad = 0;
    r.u1 = g.dim > 1 ? elementGlobal(g, 1, dy.ie, xhatOf<KT>(dy.a)) : 0.0;
    r.u2 = g.dim > 2 ? elementGlobal(g, 2, dz.ie, xhatOf<KT>(dz.a)) : 0.0;
    r.q1 = __dmul_rn(r.u1, r.u1); r.q2 = __dmul_rn(r.u2, r.u2);
    rows[rI] = r;
  }
  const unsigned C = (unsigned)A.nActive;                         // 1, or the components of an interleaved group
  const bool scalar = f.ncomp() == 1;
  const unsigned lim = (unsigned)g.N[0] * C;
  unsigned u = threadIdx.x, c0 = 0, lI = 0;
  double hx[KT], hq[KT];            // hq: squares of hx (F::kFromSquares)
  // x-data of x-entity c0: nodes X = k c0 + a, a = 0 vertex-type, 1 .. k-1 edge-type; owner element = c0 itself
  auto xprep = [&](unsigned uu) {
    A.divC.divmod(uu, c0, lI);
    const double lo = gridCoordinate(g, 0, g.off[0] + (int)c0), up = gridCoordinate(g, 0, g.off[0] + (int)c0 + 1);
    const double d = __dsub_rn(up, lo);
#pragma unroll
    for (int a = 0; a < KT; ++a) {
      hx[a] = __dadd_rn(lo, __dmul_rn((double)a / (double)KT, d));   // alpha / k folded at compile time
      hq[a] = __dmul_rn(hx[a], hx[a]);
    }
  };
  if (u < lim) xprep(u);            // before the barrier: overlaps the row set-up of the first warp
  __syncthreads();
  while (u < lim) {
    const int comp = scalar ? 0 : A.leafComp[lI];
    const unsigned off8 = 8u * lI;                                 // leaf lI sits lI entries behind leaf 0
#pragma unroll 4
    for (unsigned rI = 0; rI < nr; ++rI) {
      const LeanRow r = rows[rI];
#pragma unroll
      for (int a = 0; a < KT; ++a) {
        const double cm = F::kFromSquares ? f.commonSq(hq[a], r.q1, r.q2) : f.common(hx[a], r.u1, r.u2);
        const double v = f.comp(comp, cm, hx[a], r.u1, r.u2);
        char* p = a == 0 ? reinterpret_cast<char*>(r.pV) + (size_t)r.sV8 * c0
                         : reinterpret_cast<char*>(r.pE) + (size_t)r.sE8 * c0 + (size_t)r.a8 * (unsigned)(a - 1);
        *reinterpret_cast<double*>(p + off8) = v;
      }
    }
    u += blockDim.x;
    if (u < lim) xprep(u);
  }
  // the closing vertex plane x-entity N_x: one node per row and leaf; its owner is the element right of it if
  // the grid has one, else the last element with alpha = k (interpolate.hh:254-259, last writer)
  for (unsigned w = threadIdx.x; w < nr * C; w += blockDim.x) {
    unsigned rI, lI;
    A.divC.divmod(w, rI, lI);
    const LeanRow r = rows[rI];
    int ie = g.off[0] + g.N[0], aa = 0;
    if (ie >= g.G[0]) { ie = g.G[0] - 1; aa = k; }
    const double xc = elementGlobal(g, 0, ie, xhatOf<KT>(aa));
    const double v = f.comp(scalar ? 0 : A.leafComp[lI], f.common(xc, r.u1, r.u2), xc, r.u1, r.u2);
    *reinterpret_cast<double*>(reinterpret_cast<char*>(r.pV) + (size_t)r.sV8 * (unsigned)g.N[0] + 8u * lI) = v;
  }
}

template <class F, int KT>
__global__ void __launch_bounds__(256)
k_interp_rows_lean(const __grid_constant__ DevBasis B, const __grid_constant__ RowArgs A, const __grid_constant__ F f,
                   double* __restrict__ coeff) {
  __shared__ LeanRow rows[kRowsPerBlock];
  leanRowsBody<F, KT>(B, A, f, coeff, blockIdx.x, rows);
}

// Two interpolate calls in ONE launch (dfx_interpolate_batch): the blocks [0, blocksA) sweep part A's lattice,
// the rest part B's — e.g. the Taylor-Hood velocity (Q2 lattice, x -> x) and pressure (Q1 lattice,
// exp(-|x|^2)) of examples/interpolation.cc:89-96, whose kernels are each shorter than a launch gap.
template <class FA, int KTA, class FB, int KTB>
__global__ void __launch_bounds__(256)
k_interp_rows_lean2(const __grid_constant__ DevBasis B, const __grid_constant__ RowArgs A, const __grid_constant__ FA fa,
                    const __grid_constant__ RowArgs A2, const __grid_constant__ FB fb, unsigned blocksA,
                    double* __restrict__ coeff) {
  __shared__ LeanRow rows[kRowsPerBlock];
  if (blockIdx.x < blocksA) leanRowsBody<FA, KTA>(B, A, fa, coeff, blockIdx.x, rows);
  else leanRowsBody<FB, KTB>(B, A2, fb, coeff, blockIdx.x - blocksA, rows);
}

// ---- per-node evaluation for continuous orders >= 3: sweep RUNS of the numbering, not lattice rows ----------
// For order k >= 3 an edge / facet / element carries (k-1)^m interior DOFs that the reference numbers entity by
// entity (dofLayout, lagrangebasis.hh:747-762), so a lattice row touches 16 of every 64 bytes of the interior
// component and 4 different rows complete a sector.  In every index-set component the entities of one
// (c1, c2) row are consecutive, so their DOFs form ONE contiguous run of csize[s][0] * blk[s] entries: a block
// takes a few such runs of one component, thread (tx, ty) owns in-entity DOF fr = tx % blk — its lattice offsets
// (a0, a1, a2) are constants — and x-entities tx / blk, + TX / blk, ...; the x-coordinate is computed once per
// (thread, x-entity) and swept over the block's runs, whose y / z node coordinates sit in shared memory.  Every
// warp store is 256 contiguous bytes.
struct RunArgs {
  int layoutId, leafId, comp;
  int runsPerBlock[8];          // per component (index = shift mask s)
  unsigned blockStart[9];       // first block of component order[q], q = 0 .. ncomp
  FastDiv divBlk[8], divCs1[8];
};
constexpr int kRunRows = 32;
template <int K> struct RunRow { double* base; double y[K - 1], z[K - 1]; unsigned obase, pad; double yq[K - 1], zq[K - 1]; };   // yq, zq: squares (F::kFromSquares)   // obase: ORIENT, entity index of the run's first entity

// ORIENT (dfx_basis_set_vertex_ids): slot t of an edge / facet block holds the node the entity's globally unique
// orientation numbers t; the element numbers it applyOrientCode<INVERSE>(code, t) (lagrangebasis.hh:863-869 read
// backwards, as dofOwnerNode does), so the thread reads the entity's code byte and picks its node's offsets per entity.
template <class F, int K, int TX, int TY, bool ORIENT>
__global__ void __launch_bounds__(TX * TY)
k_interp_runs(const __grid_constant__ DevBasis B, const __grid_constant__ RunArgs A, const __grid_constant__ F f,
              double* __restrict__ coeff) {
  __shared__ RunRow<K> rows[kRunRows];
  const DevGrid& g = B.grid;
  const DevLayout& L = B.layouts[A.layoutId];
  const DevLeaf& leaf = B.leaves[A.leafId];
  // component of this block
  int q = 0;
  while (q + 1 < L.ncomp && blockIdx.x >= A.blockStart[q + 1]) ++q;
  const int s = L.order[q];
  const unsigned blk = (unsigned)L.blk[s];
  const unsigned cs0 = (unsigned)L.csize[s][0], cs1 = (unsigned)L.csize[s][1], cs2 = (unsigned)L.csize[s][2];
  const unsigned nRuns = cs1 * cs2;
  const unsigned R = (unsigned)A.runsPerBlock[s];
  const unsigned row0 = (blockIdx.x - A.blockStart[q]) * R;
  const unsigned nr = min(R, nRuns - row0);
  const unsigned tid = threadIdx.y * TX + threadIdx.x;
  // coordinate of node `a` (1 .. K-1) of entity c along a free direction / of lattice plane c along a pinned one
  auto freeCoord = [&](int j, unsigned c, int a) { return elementGlobal(g, j, g.off[j] + (int)c, xhatOf<K>(a)); };
  auto planeCoord = [&](int j, unsigned c) {
    // shared by the elements on both sides: the later one in iteration order writes last (interpolate.hh:254-259)
    const int gc = g.off[j] + (int)c;
    return gc < g.G[j] ? elementGlobal(g, j, gc, 0.0) : elementGlobal(g, j, g.G[j] - 1, 1.0);
  };
  if (tid < nr) {
    unsigned c1, c2;
    A.divCs1[s].divmod(row0 + tid, c2, c1);
    RunRow<K> r;
    const u64 j0 = L.compOffset[s] + (u64)blk * ((u64)cs0 * ((u64)c1 + (u64)cs1 * (u64)c2));
    r.base = coeff + (leaf.posA * j0 + leaf.posB);
    r.obase = cs0 * (c1 + cs1 * c2); r.pad = 0;
#pragma unroll
    for (int a = 1; a < K; ++a) {
      r.y[a - 1] = g.dim > 1 ? ((s & 2) ? freeCoord(1, c1, a) : planeCoord(1, c1)) : 0.0;
      r.z[a - 1] = g.dim > 2 ? ((s & 4) ? freeCoord(2, c2, a) : planeCoord(2, c2)) : 0.0;
      r.yq[a - 1] = __dmul_rn(r.y[a - 1], r.y[a - 1]); r.zq[a - 1] = __dmul_rn(r.z[a - 1], r.z[a - 1]);
    }
    rows[tid] = r;
  }
  // this thread's in-entity DOF: digits of fr over the free directions, x first (LocalKey::index order)
  unsigned fr = threadIdx.x % blk;
  int a0 = 0, yi = 0, zi = 0;
  if (s & 1) { a0 = 1 + (int)(fr % (K - 1)); fr /= (K - 1); }
  if (s & 2) { yi = (int)(fr % (K - 1)); fr /= (K - 1); }
  if (s & 4) { zi = (int)(fr % (K - 1)); }
  const unsigned stride8 = (unsigned)leaf.posA * 8u;
  const int comp = f.ncomp() == 1 ? 0 : A.comp;
  const unsigned len = cs0 * blk;
  const unsigned mainLen = len - len % TX;       // whole passes of TX entries; the rest (a few entries of the closing
                                                 // lattice plane when x is pinned) goes item by item below
  __syncthreads();
  // ORIENT: edges (one free direction, 2-d and 3-d) and quadrilateral facets (two free directions, 3-d)
  const int nfree = (s & 1) + ((s >> 1) & 1) + ((s >> 2) & 1);
  const bool permuted = ORIENT && B.orient != nullptr && ((nfree == 1 && g.dim > 1) || (nfree == 2 && g.dim == 3));
  const unsigned char* __restrict__ otab = permuted ? B.orient + B.orientOff[s] : nullptr;
  const unsigned slot = threadIdx.x % blk;
  if (permuted) {
    for (unsigned u = threadIdx.x; u < mainLen; u += TX) {
      const unsigned c0 = A.divBlk[s].div(u);
      double xs[K - 1];                            // x of the node for every in-entity offset along x (x free), else one value
#pragma unroll
      for (int a = 1; a < K; ++a) xs[a - 1] = (s & 1) ? freeCoord(0, c0, a) : planeCoord(0, c0);
      const size_t off = (size_t)stride8 * u;
      for (unsigned r = threadIdx.y; r < nr; r += TY) {
        const RunRow<K>& row = rows[r];
        unsigned fl = applyOrientCode<true>((unsigned)__ldg(otab + row.obase + c0), nfree, K, slot);
        int b0 = 0, yj = 0, zj = 0;
        if (s & 1) { b0 = (int)(fl % (K - 1)); fl /= (K - 1); }
        if (s & 2) { yj = (int)(fl % (K - 1)); fl /= (K - 1); }
        if (s & 4) { zj = (int)(fl % (K - 1)); }
        double x = xs[0];
#pragma unroll
        for (int a = 1; a < K - 1; ++a) x = b0 == a ? xs[a] : x;
        const double y = row.y[yj], z = row.z[zj];
        *reinterpret_cast<double*>(reinterpret_cast<char*>(row.base) + off) = f.comp(comp, f.common(x, y, z), x, y, z);
      }
    }
  } else {
    for (unsigned u = threadIdx.x; u < mainLen; u += TX) {
      const unsigned c0 = A.divBlk[s].div(u);
      const double x = (s & 1) ? freeCoord(0, c0, a0) : planeCoord(0, c0);
      const double xq = __dmul_rn(x, x);
      const size_t off = (size_t)stride8 * u;
#pragma unroll 4
      for (unsigned r = threadIdx.y; r < nr; r += TY) {
        const RunRow<K>& row = rows[r];
        const double y = row.y[yi], z = row.z[zi];
        const double cm = F::kFromSquares ? f.commonSq(xq, row.yq[yi], row.zq[zi]) : f.common(x, y, z);
        *reinterpret_cast<double*>(reinterpret_cast<char*>(row.base) + off) = f.comp(comp, cm, x, y, z);
      }
    }
  }
  // ---- the last len % TX entries of every run: one (run, entry) item per thread ----
  const unsigned rem = len - mainLen;
  for (unsigned t = tid; t < rem * nr; t += TX * TY) {
    const unsigned r = t / rem, u = mainLen + (t - r * rem);
    const unsigned c0 = A.divBlk[s].div(u);
    unsigned fu = u - c0 * blk;
    if (permuted) fu = applyOrientCode<true>((unsigned)__ldg(otab + rows[r].obase + c0), nfree, K, fu);
    int b0 = 0, yj = 0, zj = 0;
    if (s & 1) { b0 = 1 + (int)(fu % (K - 1)); fu /= (K - 1); }
    if (s & 2) { yj = (int)(fu % (K - 1)); fu /= (K - 1); }
    if (s & 4) { zj = (int)(fu % (K - 1)); }
    const double x = (s & 1) ? freeCoord(0, c0, b0) : planeCoord(0, c0);
    const RunRow<K>& row = rows[r];
    const double y = row.y[yj], z = row.z[zj];
    *reinterpret_cast<double*>(reinterpret_cast<char*>(row.base) + (size_t)stride8 * u) = f.comp(comp, f.common(x, y, z), x, y, z);
  }
}

struct CellArgs {
  int k, dim, nloc;              // direct copies: no indexed-constant loads in the kernel
  int off[3];
  FastDiv divN0, divN1, divK1;
  u64 ne;                        // elements of the local box
  int tabOff[3];
  int nActive;                   // leaves of the subtree on this lattice: position map + range entry
  unsigned long long posA[DFX_MAX_LEAVES], posB[DFX_MAX_LEAVES];
  int comp[DFX_MAX_LEAVES];
};

constexpr int kCellTask = 32;     // consecutive elements swept by one warp
constexpr int kCellWarps = 8;

// Element-local DOFs (LagrangeDG, Lagrange<0>): DOF index = element * nloc + i, so one element
// is a contiguous run of nloc doubles.  A warp sweeps kCellTask consecutive elements and
// writes each of them completely before moving on (lane i writes local indices i, i + 32,
// ...), so a warp's stores walk through memory strictly forward in 256-byte runs.  The
// per-lane data of up to NP local indices (lattice multi-index, reference coordinate, the
// y/z factors or coordinates of the current element row) stay in registers; NP = 0 is the
// run-time variant for elements with more than 32 * 8 DOFs (one local index at a time).
// SIMPLE: one leaf on the lattice with unit position stride and no mask — the store address
// is a running pointer (the scalar case of every BASELINE config).
template <class F, int MODE, int NP, bool SIMPLE>
__global__ void __launch_bounds__(kCellWarps * 32)
k_interp_cell(const __grid_constant__ DevGrid g, const __grid_constant__ CellArgs A, const F f,
              const double* __restrict__ tab, double* __restrict__ coeff, const uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const u64 task = (u64)blockIdx.x * kCellWarps + (threadIdx.x >> 5);
  const u64 eBegin = task * kCellTask;
  if (eBegin >= A.ne) return;
  const unsigned nEl = (unsigned)((A.ne - eBegin) < (u64)kCellTask ? (A.ne - eBegin) : (u64)kCellTask);
  const int k = A.k, k1 = k + 1, nloc = A.nloc;
  const unsigned N0 = (unsigned)g.N[0], N1g = (unsigned)g.N[1];
  unsigned s0, s1, s2;                          // coordinates of the first element
  {
    unsigned q;
    A.divN0.divmod((unsigned)eBegin, q, s0);
    A.divN1.divmod(q, s2, s1);
  }
  const bool scalar = MODE == 1 || f.ncomp() == 1;
  // one DOF: common part cm, node (x0, x1, x2), scalar DOF index t -> every active leaf
  auto emit = [&](u64 t, double cm, double x0, double x1, double x2) {
    if (A.nActive == 1) {
      const u64 p = A.posA[0] * t + A.posB[0];
      if (mask == nullptr || mask[p]) coeff[p] = MODE == 1 ? cm : f.comp(scalar ? 0 : A.comp[0], cm, x0, x1, x2);
    } else {
      for (int l = 0; l < A.nActive; ++l) {
        const u64 p = A.posA[l] * t + A.posB[l];
        if (mask == nullptr || mask[p]) coeff[p] = MODE == 1 ? cm : f.comp(scalar ? 0 : A.comp[l], cm, x0, x1, x2);
      }
    }
  };
  if (NP > 0) {
    constexpr int NR = NP > 0 ? NP : 1;
    unsigned a0[NR], a1[NR], a2[NR];
    double xh0[NR];                             // MODE 0: reference coordinate alpha_0 / k of the node
    double u1[NR], u2[NR];                      // MODE 1: (t_y t_z, -); MODE 0: node coordinates x_1, x_2 of the row
#pragma unroll
    for (int t = 0; t < NP; ++t) {
      const unsigned i = (unsigned)(lane + 32 * t);
      unsigned q;
      A.divK1.divmod(i < (unsigned)nloc ? i : 0u, q, a0[t]);
      A.divK1.divmod(q, a2[t], a1[t]);
      xh0[t] = k == 0 ? 0.5 : __ddiv_rn(__dmul_rn(1.0, (double)a0[t]), (double)k);
      u1[t] = 1.0; u2[t] = 0.0;
    }
    unsigned ec0 = s0, ec1 = s1, ec2 = s2;
    bool newRow = true;
    for (unsigned n = 0; n < nEl; ++n) {
      if (newRow) {
#pragma unroll
        for (int t = 0; t < NP; ++t) {
          if (MODE == 1) {
            u1[t] = __dmul_rn(A.dim > 1 ? tab[A.tabOff[1] + (int)(ec1 * k1 + a1[t])] : 1.0,
                              A.dim > 2 ? tab[A.tabOff[2] + (int)(ec2 * k1 + a2[t])] : 1.0);
          } else {
            u1[t] = A.dim > 1 ? nodeCoord(g, k, 1, g.off[1] + (int)ec1, (int)a1[t]) : 0.0;
            u2[t] = A.dim > 2 ? nodeCoord(g, k, 2, g.off[2] + (int)ec2, (int)a2[t]) : 0.0;
          }
        }
        newRow = false;
      }
      const u64 tBase = (eBegin + n) * (u64)nloc + (u64)lane;
      if (MODE == 1 && SIMPLE) {
        const double* tx = tab + A.tabOff[0] + (int)(ec0 * k1);
        double* po = coeff + A.posB[0] + tBase;
#pragma unroll
        for (int t = 0; t < NP; ++t)
          if (lane + 32 * t < nloc) po[32 * t] = __dmul_rn(tx[a0[t]], u1[t]);
      } else if (MODE == 1) {
        const double* tx = tab + A.tabOff[0] + (int)(ec0 * k1);
#pragma unroll
        for (int t = 0; t < NP; ++t)
          if (lane + 32 * t < nloc) emit(tBase + 32 * t, __dmul_rn(tx[a0[t]], u1[t]), 0.0, 0.0, 0.0);
      } else {
        // elementGlobal(g, 0, ie, xhat) with the element's end points hoisted out of the pass loop
        const double lo = gridCoordinate(g, 0, g.off[0] + (int)ec0), up = gridCoordinate(g, 0, g.off[0] + (int)ec0 + 1);
        const double d = __dsub_rn(up, lo);
        if (SIMPLE) {
          double* po = coeff + A.posB[0] + tBase;
          const int c0 = scalar ? 0 : A.comp[0];
#pragma unroll
          for (int t = 0; t < NP; ++t)
            if (lane + 32 * t < nloc) {
              const double x0 = __dadd_rn(lo, __dmul_rn(xh0[t], d));
              po[32 * t] = f.comp(c0, f.common(x0, u1[t], u2[t]), x0, u1[t], u2[t]);
            }
        } else {
#pragma unroll
          for (int t = 0; t < NP; ++t)
            if (lane + 32 * t < nloc) {
              const double x0 = __dadd_rn(lo, __dmul_rn(xh0[t], d));
              emit(tBase + 32 * t, f.common(x0, u1[t], u2[t]), x0, u1[t], u2[t]);
            }
        }
      }
      if (++ec0 >= N0) { ec0 = 0; newRow = true; if (++ec1 >= N1g) { ec1 = 0; ++ec2; } }
    }
    return;
  }
  for (int i = lane; i < nloc; i += 32) {
    unsigned a0, a1, a2, q;
    A.divK1.divmod((unsigned)i, q, a0);
    A.divK1.divmod(q, a2, a1);
    unsigned ec0 = s0, ec1 = s1, ec2 = s2;
    bool newRow = true;
    double yz = 1.0, x1 = 0.0, x2 = 0.0;      // MODE 1: t_y t_z of the row;  MODE 0: node coordinates
    for (unsigned n = 0; n < nEl; ++n) {
      if (newRow) {
        if (MODE == 1) {
          yz = __dmul_rn(A.dim > 1 ? tab[A.tabOff[1] + (int)(ec1 * k1 + a1)] : 1.0,
                         A.dim > 2 ? tab[A.tabOff[2] + (int)(ec2 * k1 + a2)] : 1.0);
        } else {
          x1 = A.dim > 1 ? nodeCoord(g, k, 1, g.off[1] + (int)ec1, (int)a1) : 0.0;
          x2 = A.dim > 2 ? nodeCoord(g, k, 2, g.off[2] + (int)ec2, (int)a2) : 0.0;
        }
        newRow = false;
      }
      double cm, x0 = 0.0;
      if (MODE == 1) cm = __dmul_rn(tab[A.tabOff[0] + (int)(ec0 * k1 + a0)], yz);
      else { x0 = nodeCoord(g, k, 0, g.off[0] + (int)ec0, (int)a0); cm = f.common(x0, x1, x2); }
      emit((eBegin + n) * (u64)nloc + (u64)i, cm, x0, x1, x2);
      if (++ec0 >= N0) { ec0 = 0; newRow = true; if (++ec1 >= N1g) { ec1 = 0; ++ec2; } }
    }
  }
}

// |argument of exp(-a |x|^2)| <= 700 over the whole grid box: expTabT<false> applies
static bool gaussianBounded(const DevGrid& g, const dfx_function& fn) {
  double r2 = 0;
  for (int j = 0; j < g.dim; ++j) {
    const double lo = g.lower[j], up = g.lower[j] + g.h[j] * g.G[j];
    r2 += std::max(lo * lo, up * up);
  }
  return std::abs(fn.p[0]) * r2 * 1.001 <= 700.0;
}
// per-node kernels with a compile-time function id (orders 1 and 2 / up to 8 passes of 32 local indices)
template <int ID>
static bool launchRowsNode(int k, bool one, bool leanShape, unsigned grid, unsigned bs, cudaStream_t st, const DevBasis& D, RowArgs A,
                           const dfx_function& fn, double* coeff, const uint8_t* mask) {
  RegistryFnT<ID> f; f.f = fn; f.dim = D.grid.dim;
  // LEAN functor: |argument of exp| bounded by 700 over the whole grid box
  const bool bounded = ID != DFX_FN_GAUSSIAN || gaussianBounded(D.grid, fn);
  if (leanShape && bounded) {
    // one leaf or an interleaved group, unmasked: the lean kernel
    RegistryFnT<ID, true> fl; fl.f = fn; fl.dim = D.grid.dim;
    const u64 nRows = (u64)A.LX[1] * A.LX[2];
    A.rowsPerBlock = leanRows(nRows, (u64)A.LX[0] * (u64)A.nActive * sizeof(double));
    const unsigned gridL = (unsigned)((nRows + A.rowsPerBlock - 1) / A.rowsPerBlock);
    if (k == 1) k_interp_rows_lean<RegistryFnT<ID, true>, 1><<<gridL, bs, 0, st>>>(D, A, fl, coeff);
    else if (k == 2) k_interp_rows_lean<RegistryFnT<ID, true>, 2><<<gridL, bs, 0, st>>>(D, A, fl, coeff);
    else if (k == 3) k_interp_rows_lean<RegistryFnT<ID, true>, 3><<<gridL, bs, 0, st>>>(D, A, fl, coeff);
    else k_interp_rows_lean<RegistryFnT<ID, true>, 4><<<gridL, bs, 0, st>>>(D, A, fl, coeff);
    return true;
  }
  if (k > 2) return false;            // orders 3, 4 outside the lean shape: the run-time-id kernels of the caller
  if (k == 1) {
    if (one) k_interp_rows<RegistryFnT<ID>, 0, 1, true><<<grid, bs, 0, st>>>(D, A, f, nullptr, coeff, mask);
    else k_interp_rows<RegistryFnT<ID>, 0, 1, false><<<grid, bs, 0, st>>>(D, A, f, nullptr, coeff, mask);
  } else {
    if (one) k_interp_rows<RegistryFnT<ID>, 0, 2, true><<<grid, bs, 0, st>>>(D, A, f, nullptr, coeff, mask);
    else k_interp_rows<RegistryFnT<ID>, 0, 2, false><<<grid, bs, 0, st>>>(D, A, f, nullptr, coeff, mask);
  }
  return true;
}
template <int ID>
static void launchCellNode(int passes, unsigned blocks, cudaStream_t st, const DevGrid& g, const CellArgs& C, const dfx_function& fn,
                           double* coeff, const uint8_t* mask) {
  const bool lean = C.nActive == 1 && C.posA[0] == 1 && mask == nullptr && (ID != DFX_FN_GAUSSIAN || gaussianBounded(g, fn));
  if (lean) {
    // one leaf, unit stride, no mask: running store pointer, zero-padded coordinates, bounded exp argument
    RegistryFnT<ID, true> f; f.f = fn; f.dim = g.dim;
    if (passes <= 1) k_interp_cell<RegistryFnT<ID, true>, 0, 1, true><<<blocks, kCellWarps * 32, 0, st>>>(g, C, f, nullptr, coeff, mask);
    else if (passes <= 2) k_interp_cell<RegistryFnT<ID, true>, 0, 2, true><<<blocks, kCellWarps * 32, 0, st>>>(g, C, f, nullptr, coeff, mask);
    else if (passes <= 4) k_interp_cell<RegistryFnT<ID, true>, 0, 4, true><<<blocks, kCellWarps * 32, 0, st>>>(g, C, f, nullptr, coeff, mask);
    else k_interp_cell<RegistryFnT<ID, true>, 0, 8, true><<<blocks, kCellWarps * 32, 0, st>>>(g, C, f, nullptr, coeff, mask);
    return;
  }
  RegistryFnT<ID> f; f.f = fn; f.dim = g.dim;
  if (passes <= 1) k_interp_cell<RegistryFnT<ID>, 0, 1, false><<<blocks, kCellWarps * 32, 0, st>>>(g, C, f, nullptr, coeff, mask);
  else if (passes <= 2) k_interp_cell<RegistryFnT<ID>, 0, 2, false><<<blocks, kCellWarps * 32, 0, st>>>(g, C, f, nullptr, coeff, mask);
  else if (passes <= 4) k_interp_cell<RegistryFnT<ID>, 0, 4, false><<<blocks, kCellWarps * 32, 0, st>>>(g, C, f, nullptr, coeff, mask);
  else k_interp_cell<RegistryFnT<ID>, 0, 8, false><<<blocks, kCellWarps * 32, 0, st>>>(g, C, f, nullptr, coeff, mask);
}

// continuous orders 3 .. 5, one leaf, no mask: the run sweep (k_interp_runs)
template <int ID>
static bool launchRunsNode(const DevBasis& D, int q, int leafId, int comp, const dfx_function& fn, double* coeff, cudaStream_t st,
                           bool orient = false) {
  const DevLayout& L = D.layouts[q];
  const int k = L.k;
  if (k < 3 || k > 5) return false;
  if ((fn.id == DFX_FN_GAUSSIAN || fn.id == DFX_FN_GAUSSIAN_VEC) && !gaussianBounded(D.grid, fn)) return false;
  RunArgs A;
  A.layoutId = q; A.leafId = leafId; A.comp = comp;
  u64 blocks = 0;
  for (int s = 0; s < 8; ++s) { A.runsPerBlock[s] = 1; A.divBlk[s] = FastDiv(1); A.divCs1[s] = FastDiv(1); }
  for (int qq = 0; qq < L.ncomp; ++qq) {
    const int s = L.order[qq];
    const u64 len = (u64)L.csize[s][0] * (u64)L.blk[s];
    const u64 runs = (u64)L.csize[s][1] * (u64)L.csize[s][2];
    // runs per block: consecutive runs of a component are adjacent in memory, so a block's runs are ONE contiguous chunk —
    // 32 K entries of it (was 8 K: the row set-up and its barrier weighed 15-20 % of a block; Q3 256^3 0.853 -> 0.639 ms,
    // 128^3 0.112 -> 0.102 ms; 64 K entries / 64 runs: 0.654 / 0.101 ms)
    const int R = (int)std::max<u64>(1, std::min<u64>(kRunRows, 32768 / std::max<u64>(1, len)));
    A.runsPerBlock[s] = R;
    A.blockStart[qq] = (unsigned)blocks;
    blocks += (runs + R - 1) / R;
    A.divBlk[s] = FastDiv((unsigned)L.blk[s]);
    A.divCs1[s] = FastDiv((unsigned)std::max(1, L.csize[s][1]));
    if (len > 0xffffffffull || runs > 0xffffffffull) return false;
  }
  for (int qq = L.ncomp; qq <= 8; ++qq) A.blockStart[qq] = (unsigned)blocks;
  if (blocks == 0 || blocks > 0x7fffffffull) return false;
  RegistryFnT<ID, true> f; f.f = fn; f.dim = D.grid.dim;
  if (orient) {
    if (k == 3) k_interp_runs<RegistryFnT<ID, true>, 3, 128, 2, true><<<(unsigned)blocks, dim3(128, 2), 0, st>>>(D, A, f, coeff);
    else if (k == 4) k_interp_runs<RegistryFnT<ID, true>, 4, 108, 2, true><<<(unsigned)blocks, dim3(108, 2), 0, st>>>(D, A, f, coeff);
    else k_interp_runs<RegistryFnT<ID, true>, 5, 128, 2, true><<<(unsigned)blocks, dim3(128, 2), 0, st>>>(D, A, f, coeff);
    return true;
  }
  if (k == 3) k_interp_runs<RegistryFnT<ID, true>, 3, 128, 2, false><<<(unsigned)blocks, dim3(128, 2), 0, st>>>(D, A, f, coeff);
  else if (k == 4) k_interp_runs<RegistryFnT<ID, true>, 4, 108, 2, false><<<(unsigned)blocks, dim3(108, 2), 0, st>>>(D, A, f, coeff);
  else k_interp_runs<RegistryFnT<ID, true>, 5, 128, 2, false><<<(unsigned)blocks, dim3(128, 2), 0, st>>>(D, A, f, coeff);
  return true;
}

// Handles with permuted face DOFs (dfx_basis_set_vertex_ids): ONE continuous leaf of order 3 .. 5, no mask — the run
// sweep with one orientation byte per entity.  -1: not this configuration (the caller takes the general kernel).
int interpolateRunsOriented(dfx_basis* b, int leafId, const dfx_function& fn, double* coeff, cudaStream_t st) {
  const DevBasis& D0 = b->dev;
  const int q = D0.leaves[leafId].layout;
  const DevLayout& L = D0.layouts[q];
  if (L.kind != DFX_NODE_LAGRANGE || L.k < 3 || L.k > 5 || fn.n_comp != 1) return -1;
  if (b->dimension > 0xffffffffull) return -1;
  if (ensureOrientTable(b, st) != DFX_OK) return -1;
  for (unsigned s = 1; s < 7; ++s)                                  // the table numbers entities like the DOF layout
    for (int j = 0; j < D0.grid.dim; ++j)
      if ((int)(D0.grid.N[j] + (((s >> j) & 1u) ? 0 : 1)) != L.csize[s][j] && L.compCount[s] > 0) return -1;
  const DevBasis& D = b->dev;                                       // with the table published
  bool launched = false;
  switch (fn.id) {
    case DFX_FN_GAUSSIAN: launched = launchRunsNode<DFX_FN_GAUSSIAN>(D, q, leafId, 0, fn, coeff, st, true); break;
    case DFX_FN_QUADRATIC: launched = launchRunsNode<DFX_FN_QUADRATIC>(D, q, leafId, 0, fn, coeff, st, true); break;
    default: launched = launchRunsNode<-1>(D, q, leafId, 0, fn, coeff, st, true); break;
  }
  if (!launched) return -1;
  return postLaunch("k_interp_runs");
}

static bool separable(const dfx_function& f) {
  if (f.n_comp == 1) return f.id == DFX_FN_GAUSSIAN || f.id == DFX_FN_SINPROD || f.id == DFX_FN_COORD || f.id == DFX_FN_CONSTANT;
  // vector-valued: every range entry is a product of per-direction factors (one table set each)
  return f.id == DFX_FN_IDENTITY || f.id == DFX_FN_GAUSSIAN_VEC || f.id == DFX_FN_CONSTANT;
}

// wantSep: the caller opted into the per-direction factor tables (dfx_interpolate_separable); otherwise
// f is evaluated at every node, as interpolate.hh:151-173 does
int interpolateFast(dfx_basis* b, int firstLeaf, int nLeaves, const dfx_function& fn, double* coeff,
                    const uint8_t* mask, bool wantSep, cudaStream_t st) {
  const DevBasis& D = b->dev;
  const DevGrid& g = D.grid;
  RegistryFn f; f.f = fn; f.dim = g.dim;
  const bool sep = wantSep && separable(fn);
  for (int q = 0; q < D.nLayouts; ++q) {
    bool used = false;
    for (int l = 0; l < nLeaves; ++l) used = used || D.leaves[firstLeaf + l].layout == q;
    if (!used) continue;
    const DevLayout& L = D.layouts[q];
    const int k = L.k;
    const bool cell = L.kind == DFX_NODE_LAGRANGE_DG || k == 0;
    RowArgs A;
    A.layoutId = q; A.firstLeaf = firstLeaf; A.nLeaves = nLeaves;
    for (int j = 0; j < 3; ++j) A.LX[j] = j < g.dim ? (cell ? g.N[j] * (k + 1) : k * g.N[j] + 1) : 1;
    A.tabOff[0] = 0; A.tabOff[1] = A.LX[0]; A.tabOff[2] = A.LX[0] + A.LX[1];
    A.divK = FastDiv((unsigned)(k > 0 ? k : 1));
    const int tabN = A.LX[0] + A.LX[1] + A.LX[2];
    A.tabN = tabN; A.vectorTables = fn.n_comp > 1 ? 1 : 0;
    double* tab = nullptr;
    // vector-valued tables are wired into the lattice-row kernel only; element-local layouts (DG)
    // evaluate such an f at every node
    const bool sepAll = sep;
    const bool sep = sepAll && (fn.n_comp == 1 || !cell);
    if (sep) {
      const size_t need = (size_t)tabN * fn.n_comp * sizeof(double);
      if (b->dSepBytes < need) {
        if (b->dSep) cudaFree(b->dSep);
        b->dSep = nullptr; b->dSepBytes = 0;
        DFX_CUDA(cudaMalloc(&b->dSep, need));
        b->dSepBytes = need;
      }
      tab = (double*)b->dSep;
      k_sep_tables<<<(tabN * fn.n_comp + 255) / 256, 256, 0, st>>>(g, fn, A, k, cell ? 1 : 0, tab);
      int rc = postLaunch("k_sep_tables");
      if (rc) return rc;
    }
    if (!cell) {
      const u64 nRows = (u64)A.LX[1] * A.LX[2];
      if (nRows > 0xffffffffull) return fail(DFX_ERR_RANGE, "lattice too large for one launch");
      A.divLY = FastDiv((unsigned)A.LX[1]);
      A.nActive = 0;
      for (int l = 0; l < nLeaves; ++l)
        if (D.leaves[firstLeaf + l].layout == q) {
          A.leafId[A.nActive] = firstLeaf + l; A.leafComp[A.nActive] = l;
          A.leafPosA[A.nActive] = (unsigned)D.leaves[firstLeaf + l].posA;
          ++A.nActive;
        }
      // rows per block: the table-driven kernel streams best with few rows per block (the blocks in
      // flight then cover a compact band of every index-set component: 0.19 -> 0.175 ms at 256^3, 1.67 ->
      // 1.27 ms at 512^3), the kernel that evaluates f at every node amortises its per-thread x-data over many
      A.rowsPerBlock = sep ? (L.dimension >= 50000000ull ? 4 : 16) : kRowsPerBlock;    // L2-resident vectors like more rows
      // interleaved group: C leaves with position stride C and consecutive offsets
      A.interleaved = 0;
      A.divC = FastDiv((unsigned)std::max(1, A.nActive));
      if (A.nActive > 1) {
        bool il = true;
        for (int q2 = 0; q2 < A.nActive; ++q2) {
          const DevLeaf& lf = D.leaves[A.leafId[q2]];
          il = il && lf.posA == (u64)A.nActive && lf.posB == D.leaves[A.leafId[0]].posB + (u64)q2;
        }
        A.interleaved = il ? 1 : 0;
      }
      const unsigned grid = (unsigned)((nRows + A.rowsPerBlock - 1) / A.rowsPerBlock);
      // threads per block: the row's x-entities split evenly over the fewest <= 256-wide chunks
      // (the generic per-node kernel walks x-entities, only the table kernel and the lean kernel spread the leaves)
      const bool leanShape = !sep && k <= 4 && mask == nullptr && (A.nActive == 1 || A.interleaved);
      const unsigned nEnt = (unsigned)g.N[0] * (((sep || leanShape) && A.interleaved) ? (unsigned)A.nActive : 1u), chunks = (nEnt + 255) / 256;
      const unsigned bs = std::max(32u, (((nEnt + chunks - 1) / chunks + 31) / 32) * 32);
      // per-node evaluation with the function id known at compile time for the functions the BASELINE
      // configs interpolate (no switch in the node loop); everything else dispatches on f.id at run time
      bool launched = false;
      if (!sep && k >= 3 && k <= 5 && A.nActive == 1 && mask == nullptr) {
        // orders 3 .. 5: sweep the contiguous runs of the numbering (entity-blocked interior DOFs)
        switch (fn.id) {
          case DFX_FN_GAUSSIAN: launched = launchRunsNode<DFX_FN_GAUSSIAN>(D, q, A.leafId[0], A.leafComp[0], fn, coeff, st); break;
          case DFX_FN_QUADRATIC: launched = launchRunsNode<DFX_FN_QUADRATIC>(D, q, A.leafId[0], A.leafComp[0], fn, coeff, st); break;
          default: launched = launchRunsNode<-1>(D, q, A.leafId[0], A.leafComp[0], fn, coeff, st); break;
        }
        if (launched) { int rc = postLaunch("k_interp_runs"); if (rc) return rc; continue; }
      }
      if (!sep && (k <= 2 || leanShape)) {
        switch (fn.id) {
          case DFX_FN_GAUSSIAN: launched = launchRowsNode<DFX_FN_GAUSSIAN>(k, A.nActive == 1, leanShape, grid, bs, st, D, A, fn, coeff, mask); break;
          case DFX_FN_QUADRATIC: launched = launchRowsNode<DFX_FN_QUADRATIC>(k, A.nActive == 1, leanShape, grid, bs, st, D, A, fn, coeff, mask); break;
          case DFX_FN_IDENTITY: launched = launchRowsNode<DFX_FN_IDENTITY>(k, A.nActive == 1, leanShape, grid, bs, st, D, A, fn, coeff, mask); break;
          default: break;
        }
      }
#define DFX_ROWS2(KT_, ONE_)                                                                                 \
  { if (sep) k_interp_rows<RegistryFn, 1, KT_, ONE_><<<grid, bs, 0, st>>>(D, A, f, tab, coeff, mask);      \
    else k_interp_rows<RegistryFn, 0, KT_, ONE_><<<grid, bs, 0, st>>>(D, A, f, tab, coeff, mask); }
#define DFX_ROWS(KT_) { if (A.nActive == 1) DFX_ROWS2(KT_, true) else DFX_ROWS2(KT_, false) }
      if (!launched)
        switch (k) {
          case 1: DFX_ROWS(1) break;
          case 2: DFX_ROWS(2) break;
          case 3: DFX_ROWS(3) break;
          case 4: DFX_ROWS(4) break;
          default: DFX_ROWS(0)
        }
#undef DFX_ROWS2
#undef DFX_ROWS
      int rc = postLaunch("k_interp_rows");
      if (rc) return rc;
    } else {
      CellArgs C;
      C.k = k; C.dim = g.dim; C.nloc = L.nloc;
      for (int j = 0; j < 3; ++j) C.off[j] = g.off[j];
      C.nActive = 0;
      for (int l = 0; l < nLeaves; ++l)
        if (D.leaves[firstLeaf + l].layout == q) {
          C.posA[C.nActive] = D.leaves[firstLeaf + l].posA; C.posB[C.nActive] = D.leaves[firstLeaf + l].posB;
          C.comp[C.nActive] = l; ++C.nActive;
        }
      C.divN0 = FastDiv((unsigned)g.N[0]); C.divN1 = FastDiv((unsigned)g.N[1]);
      C.divK1 = FastDiv((unsigned)(k + 1));
      C.ne = (u64)g.N[0] * g.N[1] * g.N[2];
      for (int j = 0; j < 3; ++j) C.tabOff[j] = A.tabOff[j];
      const u64 tasks = (C.ne + kCellTask - 1) / kCellTask;
      const u64 blocks = (tasks + kCellWarps - 1) / kCellWarps;
      if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "basis too large for one launch");
      const int passes = (L.nloc + 31) / 32;
      const bool simple = sep && C.nActive == 1 && C.posA[0] == 1 && mask == nullptr;
      bool launched = false;
      if (!sep && passes <= 8) {
        switch (fn.id) {
          case DFX_FN_GAUSSIAN: launchCellNode<DFX_FN_GAUSSIAN>(passes, (unsigned)blocks, st, g, C, fn, coeff, mask); launched = true; break;
          case DFX_FN_QUADRATIC: launchCellNode<DFX_FN_QUADRATIC>(passes, (unsigned)blocks, st, g, C, fn, coeff, mask); launched = true; break;
          default: break;
        }
      }
#define DFX_CELL(NP_)                                                                                              \
  {                                                                                                                \
    if (simple) k_interp_cell<RegistryFn, 1, NP_, true><<<(unsigned)blocks, kCellWarps * 32, 0, st>>>(g, C, f, tab, coeff, mask);       \
    else if (sep) k_interp_cell<RegistryFn, 1, NP_, false><<<(unsigned)blocks, kCellWarps * 32, 0, st>>>(g, C, f, tab, coeff, mask);   \
    else k_interp_cell<RegistryFn, 0, NP_, false><<<(unsigned)blocks, kCellWarps * 32, 0, st>>>(g, C, f, tab, coeff, mask);             \
  }
      if (launched) {}
      else if (passes <= 1) DFX_CELL(1)
      else if (passes <= 2) DFX_CELL(2)
      else if (passes <= 4) DFX_CELL(4)
      else if (passes <= 8) DFX_CELL(8)
      else DFX_CELL(0)
#undef DFX_CELL
      int rc = postLaunch("k_interp_cell");
      if (rc) return rc;
    }
  }
  return DFX_OK;
}

// ---------------------------------------------------------------------------
// dfx_interpolate_batch: several (subtree, f) pairs on one coefficient vector.  Two parts that are each one
// lean lattice sweep (continuous order 1 / 2, one leaf or an interleaved group, no mask) run as ONE launch.
static bool leanPlan(const dfx_basis* b, int firstLeaf, int nLeaves, const dfx_function& fn, RowArgs& A, int& k, unsigned& bs,
                     unsigned& grid) {
  const DevBasis& D = b->dev;
  const DevGrid& g = D.grid;
  if (b->oriented || b->dimension > 0xffffffffull) return false;
  const int q = D.leaves[firstLeaf].layout;
  for (int l = 1; l < nLeaves; ++l) if (D.leaves[firstLeaf + l].layout != q) return false;
  const DevLayout& L = D.layouts[q];
  k = L.k;
  if (L.kind != DFX_NODE_LAGRANGE || k < 1 || k > 2) return false;
  if ((fn.id == DFX_FN_GAUSSIAN || fn.id == DFX_FN_GAUSSIAN_VEC) && !gaussianBounded(g, fn)) return false;
  A.layoutId = q; A.firstLeaf = firstLeaf; A.nLeaves = nLeaves;
  for (int j = 0; j < 3; ++j) A.LX[j] = j < g.dim ? k * g.N[j] + 1 : 1;
  A.tabOff[0] = 0; A.tabOff[1] = A.LX[0]; A.tabOff[2] = A.LX[0] + A.LX[1];
  A.divK = FastDiv((unsigned)k);
  A.tabN = 0; A.vectorTables = 0;
  const u64 nRows = (u64)A.LX[1] * A.LX[2];
  if (nRows > 0x7fffffffull) return false;
  A.divLY = FastDiv((unsigned)A.LX[1]);
  A.nActive = nLeaves;
  bool il = true;
  for (int l = 0; l < nLeaves; ++l) {
    const DevLeaf& lf = D.leaves[firstLeaf + l];
    A.leafId[l] = firstLeaf + l; A.leafComp[l] = l; A.leafPosA[l] = (unsigned)lf.posA;
    il = il && lf.posA == (u64)nLeaves && lf.posB == D.leaves[firstLeaf].posB + (u64)l;
  }
  if (nLeaves > 1 && !il) return false;
  A.interleaved = nLeaves > 1 ? 1 : 0;
  A.divC = FastDiv((unsigned)nLeaves);
  A.rowsPerBlock = leanRows(nRows, (u64)A.LX[0] * (u64)nLeaves * sizeof(double));
  grid = (unsigned)((nRows + A.rowsPerBlock - 1) / A.rowsPerBlock);
  const unsigned nEnt = (unsigned)g.N[0] * (unsigned)nLeaves, chunks = (nEnt + 255) / 256;
  bs = std::max(32u, (((nEnt + chunks - 1) / chunks + 31) / 32) * 32);
  return true;
}

template <int IDA, int IDB>
static bool launchLean2(int kA, int kB, unsigned gridA, unsigned gridB, unsigned bs, cudaStream_t st, const DevBasis& D,
                        const RowArgs& A, const dfx_function& fa, const RowArgs& A2, const dfx_function& fb, double* coeff) {
  RegistryFnT<IDA, true> FA; FA.f = fa; FA.dim = D.grid.dim;
  RegistryFnT<IDB, true> FB; FB.f = fb; FB.dim = D.grid.dim;
  const unsigned grid = gridA + gridB;
#define DFX_L2(KA_, KB_)                                                                                         \
  if (kA == KA_ && kB == KB_) {                                                                                  \
    k_interp_rows_lean2<RegistryFnT<IDA, true>, KA_, RegistryFnT<IDB, true>, KB_><<<grid, bs, 0, st>>>(D, A, FA, A2, FB, gridA, coeff); \
    return true;                                                                                                 \
  }
  DFX_L2(2, 1) DFX_L2(1, 1) DFX_L2(2, 2) DFX_L2(1, 2)
#undef DFX_L2
  return false;
}

// -1: not two lean sweeps (the caller runs the parts one after the other)
int interpolateBatch2(dfx_basis* b, int flA, int nlA, const dfx_function& fa, int flB, int nlB, const dfx_function& fb,
                      double* coeff, cudaStream_t st) {
  RowArgs A, A2;
  int kA, kB;
  unsigned bsA, bsB, gridA, gridB;
  if (!leanPlan(b, flA, nlA, fa, A, kA, bsA, gridA) || !leanPlan(b, flB, nlB, fb, A2, kB, bsB, gridB)) return -1;
  if ((u64)gridA + gridB > 0x7fffffffull) return -1;
  const unsigned bs = std::max(bsA, bsB);
  bool ok;
  if (fa.id == DFX_FN_IDENTITY && fb.id == DFX_FN_GAUSSIAN)
    ok = launchLean2<DFX_FN_IDENTITY, DFX_FN_GAUSSIAN>(kA, kB, gridA, gridB, bs, st, b->dev, A, fa, A2, fb, coeff);
  else
    ok = launchLean2<-1, -1>(kA, kB, gridA, gridB, bs, st, b->dev, A, fa, A2, fb, coeff);
  if (!ok) return -1;
  return postLaunch("k_interp_rows_lean2");
}

// ---------------------------------------------------------------------------
// Index table.  A thread keeps ONE local index li for its whole life, so everything that
// depends on li only — leaf, lattice multi-index, index-set component, affine position map —
// is folded once into (P0, S, cs0, cs1) and the per-element work is
//     ent = ec0 + cs0 (ec1 + cs1 ec2);   out = P0 + S * ent
// (for LagrangeDG cs = N, so ent is the element index).  Threads tid -> (element sub-slot,
// li) with li fastest: the block's stores form one contiguous run of the table.
struct IdxArgs {
  u64 e0, ne;
  FastDiv divN0, divN1;
  int stride;       // digits per multi-index (MULTI)
  int lpb, m, iters;   // local indices per pass, elements in flight, iterations per block
  int liBegin, liEnd;  // local index range (GATHER: the subtree's; otherwise the whole tree)
};
// GATHER: where local index li of element e goes in the companion basis' coefficient layout,
// position = posA[leaf] * (e * nloc_leaf + i) + posB[leaf]  (csrc/dfx_capi.cu, orientedCompanion)
struct GatherTab { u64 posA[DFX_MAX_LEAVES], posB[DFX_MAX_LEAVES]; };

// ORIENT (dfx_basis_set_vertex_ids): the thread's node sits in an edge / quadrilateral facet whose DOFs
// the reference renumbers by the entity's orientation (lagrangebasis.hh:863-869).  Entity, component
// and the unpermuted in-entity index fr stay fixed for the thread; per element the permuted index
// fr' = orientFaceDOF(ids of the entity's vertices) replaces it: out = P0 + S ent + posA (fr' - fr).
// GATHER (needs !MULTI): instead of writing the table, dst[companion position] = src[position]
// (LocalFunction::bind, discreteglobalbasisfunction.hh:124-148).
template <class T, bool MULTI, bool ORIENT, bool GATHER>
__global__ void __launch_bounds__(256)
k_indices_fast(const DevBasis* __restrict__ Bp, const __grid_constant__ IdxArgs A,
               const unsigned* __restrict__ localTab, T* __restrict__ out,
               const double* __restrict__ src, double* __restrict__ dst, const __grid_constant__ GatherTab G) {
  const DevGrid& g = Bp->grid;
  const int nloc = Bp->nlocTotal;
  const int sub = threadIdx.x / A.lpb, lane = threadIdx.x - sub * A.lpb;
  const u64 blkEl0 = (u64)blockIdx.x * (u64)(A.m * A.iters);
  for (int li = (GATHER ? A.liBegin : 0) + lane; li < (GATHER ? A.liEnd : nloc); li += A.lpb) {
    const unsigned packed = __ldg(localTab + li);
    const DevLeaf& leaf = Bp->leaves[packed & 31u];
    const DevLayout& L = Bp->layouts[leaf.layout];
    const int a[3] = {(int)((packed >> 5) & 31u), (int)((packed >> 10) & 31u), (int)((packed >> 15) & 31u)};
    const int zero[3] = {0, 0, 0};
    // scalar index at element (0,0,0) and its stride per entity
    const u64 jbase = layoutIndex(L, g, zero, a, li - leaf.localOffset);
    unsigned s = 0;
    if (L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0) s = (1u << g.dim) - 1u;
    else
      for (int j = 0; j < 3; ++j) if (a[j] > 0 && a[j] < L.k) s |= 1u << j;
    const unsigned cs0 = (unsigned)L.csize[s][0], cs1 = (unsigned)L.csize[s][1];
    const u64 blk = (u64)L.blk[s];
    // all per-entry arithmetic runs in the output type: for 32-bit tables the true values fit
    // (checked by the caller), so wrapping 32-bit multiply-adds give the same bits as 64-bit ones
    const T P0 = (T)(leaf.posA * jbase + leaf.posB);
    const T S = (T)(leaf.posA * blk);
    T dA[DFX_MAX_DIGITS], dB[DFX_MAX_DIGITS], dF[DFX_MAX_DIGITS];
    if (MULTI) {
#pragma unroll
      for (int p = 0; p < DFX_MAX_DIGITS; ++p) {
        dA[p] = p < leaf.ndig ? (T)(leaf.digA[p] * blk) : (T)0;
        dB[p] = p < leaf.ndig ? (T)(leaf.digA[p] * jbase + leaf.digB[p]) : (T)0;
        if (ORIENT) dF[p] = p < leaf.ndig ? (T)leaf.digA[p] : (T)0;
      }
    }
    // ORIENT: is this node's in-entity index subject to the permutation, and what is it unpermuted
    bool perm = false;
    unsigned fr = 0;
    int sh[3] = {0, 0, 0};
    const T PA = (T)leaf.posA;
    const u64* __restrict__ vid = ORIENT ? Bp->vid : nullptr;
    const int kOrd = L.k;
    int nfree = 0;
    // the entity's orientation code: one byte from the handle's table (built once per set of ids by
    // k_orient_codes; entity index = the DOF layout's), else derived from the 2 / 4 vertex ids on the spot
    const unsigned char* __restrict__ otab = nullptr;
    if (ORIENT && vid != nullptr && L.kind == DFX_NODE_LAGRANGE && kOrd > 2 && g.dim > 1) {
      nfree = (int)((s & 1u) + ((s >> 1) & 1u) + ((s >> 2) & 1u));
      perm = nfree == 1 || (nfree == 2 && g.dim == 3);
      unsigned frs = 1;
      for (int j = 0; j < 3; ++j) {
        if ((s >> j) & 1u) { fr += (unsigned)(a[j] - 1) * frs; frs *= (unsigned)(kOrd - 1); }
        else sh[j] = a[j] == kOrd ? 1 : 0;
      }
      if (perm && Bp->orient != nullptr) otab = Bp->orient + Bp->orientOff[s] + (unsigned)sh[0] + cs0 * ((unsigned)sh[1] + cs1 * (unsigned)sh[2]);
    }
    auto permuted = [&](unsigned entNow, unsigned e0_, unsigned e1_, unsigned e2_) -> unsigned {
      if (otab != nullptr) return applyOrientCode<false>((unsigned)__ldg(otab + entNow), nfree, kOrd, fr);
      const int c[3] = {(int)e0_ + sh[0], (int)e1_ + sh[1], (int)e2_ + sh[2]};
      return orientFaceDOF<false>(g, vid, kOrd, s, c, fr);
    };
    // walk this thread's elements el = blkEl0 + sub, + m, + 2m, ...: coordinates advance with carry,
    // the value and the output pointer advance by constant steps between carries
    const u64 el = blkEl0 + (u64)sub;
    if (el >= A.ne) continue;
    unsigned ec0, ec1, ec2;
    {
      unsigned q;
      A.divN0.divmod((unsigned)(A.e0 + el), q, ec0);
      A.divN1.divmod(q, ec2, ec1);
    }
    const unsigned N0 = (unsigned)g.N[0], N1g = (unsigned)g.N[1];
    const unsigned m = (unsigned)A.m;
    const u64 left = (A.ne - el + m - 1) / m;
    const int nIt = left < (u64)A.iters ? (int)left : A.iters;
    unsigned ent = ec0 + cs0 * (ec1 + cs1 * ec2);
    if (MULTI) {
      T* po = out + (el * (u64)nloc + (u64)li) * (u64)A.stride;
      const size_t step = (size_t)m * nloc * A.stride;
      for (int it = 0; it < nIt; ++it) {
        T df = (T)0;
        if (ORIENT && perm) df = (T)permuted(ent, ec0, ec1, ec2) - (T)fr;
#pragma unroll
        for (int p = 0; p < DFX_MAX_DIGITS; ++p)
          if (p < A.stride) po[p] = ORIENT ? (T)(dB[p] + dA[p] * (T)ent + dF[p] * df) : (T)(dB[p] + dA[p] * (T)ent);
        po += step; ec0 += m; ent += m;
        if (ec0 >= N0) {
          do { ec0 -= N0; if (++ec1 >= N1g) { ec1 = 0; ++ec2; } } while (ec0 >= N0);
          ent = ec0 + cs0 * (ec1 + cs1 * ec2);
        }
      }
    } else {
      T* po = GATHER ? nullptr : out + (el * (u64)nloc + (u64)li);
      size_t step = (size_t)m * nloc;
      double* pd = nullptr;
      if (GATHER) {
        const unsigned lf = packed & 31u;
        const u64 stepEl = G.posA[lf] * (u64)leaf.nloc;              // companion positions of one element further on
        pd = dst + (G.posB[lf] + G.posA[lf] * ((A.e0 + el) * (u64)leaf.nloc + (u64)(li - leaf.localOffset)));
        step = (size_t)(stepEl * (u64)m);
      }
      T v = (T)(P0 + S * (T)ent);
      const T dv = (T)(S * (T)m);
      for (int it = 0; it < nIt; ++it) {
        T w = v;
        if (ORIENT && perm) w = (T)(v + PA * ((T)permuted(ent, ec0, ec1, ec2) - (T)fr));
        if (GATHER) { *pd = src[w]; pd += step; }
        else { *po = w; po += step; }
        ec0 += m; v += dv; ent += m;
        if (ec0 >= N0) {
          do { ec0 -= N0; if (++ec1 >= N1g) { ec1 = 0; ++ec2; } } while (ec0 >= N0);
          ent = ec0 + cs0 * (ec1 + cs1 * ec2);
          v = (T)(P0 + S * (T)ent);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// The same table through a shared-memory tile (identity orientations, the case of every BASELINE config):
// a block owns `epb` consecutive elements, i.e. ONE contiguous run of the table; the threads keep their
// local index as above but write into the tile (32-bit shared-memory address arithmetic instead of 64-bit
// global pointers), and the run leaves as TMA bulk stores (cp.async.bulk.global.shared::cta, SASS UBLKCP) in
// whole 16-byte units — the direct version's 4-byte stores at odd multiples of 4 left partially written
// 32-byte sectors at both ends of every warp's run, which the L2 filled from DRAM (0.64 GB read for 0.69 GB
// written on Taylor-Hood 128^3, profiles/README.md).
struct IdxTileArgs {
  u64 e0, ne;
  FastDiv divN0, divN1, divM;
  int stride;          // digits per multi-index (NDIG != 0), else 1
  int lpb, m, epb;     // local indices per pass, elements in flight, elements per block
  int bulk;            // output run of every full block is 16-byte aligned
};

// NDIG: 0 flat positions; 2, 3 multi-indices with that many digits per entry (compile time); -1 run-time stride.
// Everything that depends on the local index only comes folded from the host (IdxLi); a thread walks its
// elements row segment by row segment, so the inner loop is store / advance pointer / advance value.
template <class T, int NDIG>
__global__ void __launch_bounds__(256)
k_indices_tile(const DevBasis* __restrict__ Bp, const __grid_constant__ IdxTileArgs A,
               const IdxLi* __restrict__ idxTab, T* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char tileRaw[];
  T* tile = reinterpret_cast<T*>(tileRaw);
  const DevGrid& g = Bp->grid;
  const int nloc = Bp->nlocTotal;
  const int sub = threadIdx.x / A.lpb, lane = threadIdx.x - sub * A.lpb;
  const u64 blkEl0 = (u64)blockIdx.x * (u64)A.epb;
  const u64 left = A.ne - blkEl0;
  const unsigned nEl = left < (u64)A.epb ? (unsigned)left : (unsigned)A.epb;       // elements of this block
  const unsigned N0 = (unsigned)g.N[0], N1g = (unsigned)g.N[1];
  const unsigned m = (unsigned)A.m;
  const int stride = NDIG == 0 ? 1 : (NDIG > 0 ? NDIG : A.stride);
  const unsigned rowW = (unsigned)nloc * (unsigned)stride;                          // table entries (T) per element
  if ((unsigned)sub < nEl) {
    unsigned s0, s1, s2;                          // coordinates of this thread's first element
    {
      unsigned q;
      A.divN0.divmod((unsigned)(A.e0 + blkEl0 + (u64)sub), q, s0);
      A.divN1.divmod(q, s2, s1);
    }
    const unsigned nIt = (nEl - (unsigned)sub + m - 1) / m;
    for (int li = lane; li < nloc; li += A.lpb) {
      const IdxLi* __restrict__ t = idxTab + li;
      const unsigned cs0 = __ldg(&t->cs0), cs1 = __ldg(&t->cs1);
      unsigned ec0 = s0, ec1 = s1, ec2 = s2, it = 0;
      if (NDIG == 0) {
        // all per-entry arithmetic runs in the output type (wrapping 32-bit multiply-adds give the same bits
        // as 64-bit ones when the true value fits, which the caller checked)
        const T P0 = (T)__ldg(&t->P0), S = (T)__ldg(&t->S);
        const T dv = (T)(S * (T)m);
        T* po = tile + ((unsigned)sub * (unsigned)nloc + (unsigned)li);
        const unsigned step = m * (unsigned)nloc;
        while (it < nIt) {
          unsigned run = A.divM.div(N0 - 1u - ec0) + 1u;         // steps until the walk leaves this element row
          run = min(run, nIt - it);
          T v = (T)(P0 + S * (T)(ec0 + cs0 * (ec1 + cs1 * ec2)));
#pragma unroll 4
          for (unsigned r = 0; r < run; ++r) { *po = v; po += step; v += dv; }
          it += run; ec0 += run * m;
          while (ec0 >= N0) { ec0 -= N0; if (++ec1 >= N1g) { ec1 = 0; ++ec2; } }
        }
      } else {
        constexpr int ND = NDIG > 0 ? NDIG : DFX_MAX_DIGITS;
        T dB[ND], dA[ND], dv[ND];
#pragma unroll
        for (int p = 0; p < ND; ++p) {
          dA[p] = (T)__ldg(&t->dA[p]); dB[p] = (T)__ldg(&t->dB[p]);
          dv[p] = (T)(dA[p] * (T)m);
        }
        T* po = tile + ((unsigned)sub * (unsigned)nloc + (unsigned)li) * (unsigned)stride;
        const unsigned step = m * rowW;
        while (it < nIt) {
          unsigned run = A.divM.div(N0 - 1u - ec0) + 1u;
          run = min(run, nIt - it);
          const T ent = (T)(ec0 + cs0 * (ec1 + cs1 * ec2));
          T v[ND];
#pragma unroll
          for (int p = 0; p < ND; ++p) v[p] = (T)(dB[p] + dA[p] * ent);
#pragma unroll 2
          for (unsigned r = 0; r < run; ++r) {
#pragma unroll
            for (int p = 0; p < ND; ++p)
              if (NDIG > 0 || p < stride) { po[p] = v[p]; v[p] += dv[p]; }
            po += step;
          }
          it += run; ec0 += run * m;
          while (ec0 >= N0) { ec0 -= N0; if (++ec1 >= N1g) { ec1 = 0; ++ec2; } }
        }
      }
    }
  }
  // ---- the block's run of the table leaves in one piece ----
  T* dst = out + blkEl0 * (u64)rowW;
  const unsigned bytes = nEl * rowW * (unsigned)sizeof(T);
  if (A.bulk && (bytes & 15u) == 0u) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                   "r"((unsigned)__cvta_generic_to_shared(tile)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  __syncthreads();
  for (unsigned i = threadIdx.x; i < nEl * rowW; i += blockDim.x) dst[i] = tile[i];
}

template <class T, int NDIG>
static int launchTileKernel(const dfx_basis* b, const IdxTileArgs& A, size_t smem, u64 blocks, T* out, cudaStream_t st) {
  auto kern = k_indices_tile<T, NDIG>;
  static std::atomic<unsigned long long> configured{0};
  if (smem > 48 * 1024) { int rc = optInSharedMemory(kern, b->device, configured); if (rc) return rc; }
  kern<<<(unsigned)blocks, A.lpb * A.m, smem, st>>>((const DevBasis*)b->dDev, A, (const IdxLi*)b->dIdxTab, out);
  return postLaunch("k_indices_tile");
}

template <class T, bool MULTI>
static int launchIndicesTile(const dfx_basis* b, u64 e0, u64 e1, T* out, int stride, cudaStream_t st) {
  IdxTileArgs A;
  A.e0 = e0; A.ne = e1 - e0; A.stride = MULTI ? stride : 1;
  if (A.ne == 0) return DFX_OK;
  const int nloc = b->dev.nlocTotal;
  A.lpb = std::min(nloc, 256);
  A.m = std::max(1, 256 / A.lpb);
  const size_t rowBytes = (size_t)nloc * A.stride * sizeof(T);
  // elements per block: a tile of about 24 KB (9 blocks per SM), a multiple of 4 elements and of m
  int epb = (int)std::max<size_t>(1, (24 * 1024) / rowBytes);
  epb = std::max(4, epb / 4 * 4);
  epb = std::max(A.m, epb / A.m * A.m);
  if ((size_t)epb * rowBytes > 96 * 1024) return -1;                // huge elements: the direct kernel
  A.epb = epb;
  A.divN0 = FastDiv((unsigned)b->dev.grid.N[0]); A.divN1 = FastDiv((unsigned)b->dev.grid.N[1]);
  A.divM = FastDiv((unsigned)A.m);
  // every full block's run starts at out + blockIdx * epb * rowBytes: aligned if the base and the block pitch are
  A.bulk = (((uintptr_t)out & 15) == 0 && (((size_t)epb * rowBytes) & 15) == 0) ? 1 : 0;
  const u64 blocks = (A.ne + epb - 1) / epb;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  const size_t smem = (size_t)epb * rowBytes;
  if (!MULTI) return launchTileKernel<T, 0>(b, A, smem, blocks, out, st);
  if (stride == 2) return launchTileKernel<T, 2>(b, A, smem, blocks, out, st);
  if (stride == 3) return launchTileKernel<T, 3>(b, A, smem, blocks, out, st);
  return launchTileKernel<T, -1>(b, A, smem, blocks, out, st);
}

template <class T, bool MULTI, bool ORIENT, bool GATHER>
static int launchIndicesFast(const dfx_basis* b, u64 e0, u64 e1, T* out, int stride, int liBegin, int liEnd, const double* src,
                             double* dst, const GatherTab& G, cudaStream_t st) {
  IdxArgs A;
  A.e0 = e0; A.ne = e1 - e0; A.stride = stride;
  A.liBegin = liBegin; A.liEnd = liEnd;
  if (A.ne == 0 || liEnd <= liBegin) return DFX_OK;
  const int nloc = liEnd - liBegin;
  A.lpb = std::min(nloc, 256);
  A.m = std::max(1, 256 / A.lpb);
  A.iters = 64;
  A.divN0 = FastDiv((unsigned)b->dev.grid.N[0]); A.divN1 = FastDiv((unsigned)b->dev.grid.N[1]);
  const u64 perBlock = (u64)A.m * A.iters;
  const u64 blocks = (A.ne + perBlock - 1) / perBlock;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  // permuted face DOFs: one orientation byte per entity instead of 2 - 4 vertex ids per DOF (the kernel derives the
  // code from the ids itself when the table cannot be built)
  if (ORIENT && b->oriented) (void)ensureOrientTable(const_cast<dfx_basis*>(b), st);
  k_indices_fast<T, MULTI, ORIENT, GATHER><<<(unsigned)blocks, A.lpb * A.m, 0, st>>>(
      (const DevBasis*)b->dDev, A, (const unsigned*)b->dLocalTab, out, src, dst, G);
  return postLaunch("k_indices_fast");
}

int g_indicesTile = 1;     // 0: the direct-store walking kernel (dfx_set_kernel_path(DFX_OP_INDICES, 2))

template <class T, bool MULTI>
int indicesFast(const dfx_basis* b, u64 e0, u64 e1, T* out, int stride, cudaStream_t st) {
  GatherTab G{};
  if (b->oriented) return launchIndicesFast<T, MULTI, true, false>(b, e0, e1, out, stride, 0, b->dev.nlocTotal, nullptr, nullptr, G, st);
  if (g_indicesTile) {
    const int rc = launchIndicesTile<T, MULTI>(b, e0, e1, out, stride, st);
    if (rc >= 0) return rc;
  }
  return launchIndicesFast<T, MULTI, false, false>(b, e0, e1, out, stride, 0, b->dev.nlocTotal, nullptr, nullptr, G, st);
}

template int indicesFast<uint32_t, true>(const dfx_basis*, u64, u64, uint32_t*, int, cudaStream_t);
template int indicesFast<u64, true>(const dfx_basis*, u64, u64, u64*, int, cudaStream_t);
template int indicesFast<uint32_t, false>(const dfx_basis*, u64, u64, uint32_t*, int, cudaStream_t);
template int indicesFast<u64, false>(const dfx_basis*, u64, u64, u64*, int, cudaStream_t);

// LocalFunction::bind for elements [e0, e1) of a basis with permuted face DOFs, written in the
// coefficient layout of its companion basis (same walk as the index table, 64-bit positions)
int gatherLocalFast(const dfx_basis* b, const dfx_basis* companion, int firstLeaf, int nLeaves, const double* src, double* dst,
                    u64 e0, u64 e1, cudaStream_t st) {
  GatherTab G;
  for (int l = 0; l < DFX_MAX_LEAVES; ++l) { G.posA[l] = companion->dev.leaves[l].posA; G.posB[l] = companion->dev.leaves[l].posB; }
  const int liBegin = b->dev.leaves[firstLeaf].localOffset;
  int liEnd = liBegin;
  for (int l = firstLeaf; l < firstLeaf + nLeaves; ++l) liEnd += b->dev.leaves[l].nloc;
  return launchIndicesFast<u64, false, true, true>(b, e0, e1, (u64*)nullptr, 1, liBegin, liEnd, src, dst, G, st);
}

}  // namespace dfx
