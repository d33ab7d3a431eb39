// dfx_interp_fast.cu — structured fast paths for interpolate() and the index table.
//
//   k_interp_rows      owner-writes interpolation of continuous Lagrange leaves.  The DOF
//                      lattice has (k N_x + 1) x (k N_y + 1) x (k N_z + 1) points; a block takes
//                      16 lattice rows (Y, Z fixed per row), 16 threads put the per-row index
//                      bases into shared memory, then all threads stream (row, x-entity)
//                      items: an item writes the row's vertex-type node and the k-1 edge-type
//                      nodes right of it, so consecutive threads write consecutive addresses
//                      in both index-set components the row touches.
//   k_interp_cell      the same for element-local DOFs (LagrangeDG, Lagrange<0>).
//   k_sep_tables       per-direction factor tables of a separable f (exp(-a|x|^2) =
//                      prod_j exp(-a x_j^2), prod_j sin(w x_j), x_j, const): O(N) instead of
//                      O(N^3) transcendental evaluations.
//   k_indices_fast     element -> global index table with 32-bit magic-number division
//                      and a per-local-index lookup table.
//
// Replaces the interpolate element loop (interpolate.hh:254-259 with :151-173) and
// DefaultLocalView::bind -> PreBasis::indices (defaultlocalview.hh:95-101).
#include "dfx_device.cuh"
#include "dfx_launch.h"

namespace dfx {

struct RowArgs {
  int layoutId, firstLeaf, nLeaves;
  int LX[3];               // table length per direction
  int tabOff[3];           // separable mode: offset of direction j's table
  FastDiv divK;            // lattice coordinate -> (entity, node inside)
  FastDiv divLY;           // row index -> (Z, Y)
  FastDiv divEnt;          // block item -> (row, x-entity)
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

// factor of a separable registry function in direction j
__device__ __forceinline__ double sepFactor(const dfx_function& f, int j, double x) {
  switch (f.id) {
    case DFX_FN_GAUSSIAN: return exp(__dmul_rn(-f.p[0], __dmul_rn(x, x)));
    case DFX_FN_SINPROD: return sin(__dmul_rn(f.p[0], x));
    case DFX_FN_COORD: return j == (int)f.p[0] ? x : 1.0;
    case DFX_FN_CONSTANT: return j == 0 ? f.p[0] : 1.0;
  }
  return 0.0;
}

__global__ void __launch_bounds__(256)
k_sep_tables(const __grid_constant__ DevGrid g, const __grid_constant__ dfx_function f, const __grid_constant__ RowArgs A,
             int k, int cellMode, double* __restrict__ tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = A.tabOff[2] + A.LX[2];
  if (t >= n) return;
  const int j = t >= A.tabOff[2] ? 2 : (t >= A.tabOff[1] ? 1 : 0);
  if (j >= g.dim) { tab[t] = 1.0; return; }
  const int X = t - A.tabOff[j];
  int ie, a;
  if (cellMode) { ie = g.off[j] + X / (k + 1); a = X % (k + 1); }
  else { const Dir d = latticeDir(g, A.divK, k, j, X); ie = d.ie; a = d.a; }
  tab[t] = sepFactor(f, j, nodeCoord(g, k, j, ie, a));
}

// per lattice row (Y, Z fixed): bases of the two index-set components it touches
struct RowInfo {
  u64 rowV, rowE;        // scalar index of entity 0's vertex-type / first edge-type node
  double u1, u2;         // MODE 0: node coordinates x_1, x_2;  MODE 1: table factors t_y, t_z
  int blkV, blkE;
};
constexpr int kRowsPerBlock = 16;

// MODE 0: f evaluated at every node; MODE 1: product of per-direction tables
template <class F, int MODE>
__global__ void __launch_bounds__(256)
k_interp_rows(const __grid_constant__ DevBasis B, const __grid_constant__ RowArgs A, const F f,
              const double* __restrict__ tab, double* __restrict__ coeff, const uint8_t* __restrict__ mask) {
  __shared__ RowInfo rows[kRowsPerBlock];
  const DevGrid& g = B.grid;
  const DevLayout& L = B.layouts[A.layoutId];
  const int k = L.k;                          // >= 1
  const unsigned nRows = (unsigned)A.LX[1] * (unsigned)A.LX[2];
  const unsigned row0 = blockIdx.x * kRowsPerBlock;
  const unsigned nr = min((unsigned)kRowsPerBlock, nRows - row0);
  if (threadIdx.x < nr) {
    unsigned Z, Y;
    A.divLY.divmod(row0 + threadIdx.x, Z, Y);
    const Dir dy = latticeDir(g, A.divK, k, 1, (int)Y);
    const Dir dz = latticeDir(g, A.divK, k, 2, (int)Z);
    unsigned syz = 0;
    u64 fyz = 0, fs = 1;
    if (g.dim > 1 && dy.inner) { syz |= 2u; fyz += (u64)(dy.inner - 1) * fs; fs *= (u64)(k - 1); }
    if (g.dim > 2 && dz.inner) { syz |= 4u; fyz += (u64)(dz.inner - 1) * fs; }
    const unsigned sV = syz, sE = syz | 1u;
    RowInfo r;
    r.blkV = L.blk[sV]; r.blkE = L.blk[sE];
    r.rowV = L.compOffset[sV] + (u64)L.blk[sV] * ((u64)L.csize[sV][0] * ((u64)dy.c + (u64)L.csize[sV][1] * (u64)dz.c)) + fyz;
    r.rowE = L.compOffset[sE] + (u64)L.blk[sE] * ((u64)L.csize[sE][0] * ((u64)dy.c + (u64)L.csize[sE][1] * (u64)dz.c)) + fyz * (u64)(k - 1);
    if (MODE == 0) {
      r.u1 = g.dim > 1 ? nodeCoord(g, k, 1, dy.ie, dy.a) : 0.0;
      r.u2 = g.dim > 2 ? nodeCoord(g, k, 2, dz.ie, dz.a) : 0.0;
    } else {
      r.u1 = tab[A.tabOff[1] + (int)Y];
      r.u2 = tab[A.tabOff[2] + (int)Z];
    }
    rows[threadIdx.x] = r;
  }
  __syncthreads();
  const bool scalar = MODE == 1 || f.ncomp() == 1;
  const unsigned nEnt = (unsigned)g.N[0] + 1u;          // x-entities per row
  const unsigned items = nr * nEnt;
  for (unsigned w = threadIdx.x; w < items; w += blockDim.x) {
    unsigned rI, c0;
    A.divEnt.divmod(w, rI, c0);
    const RowInfo r = rows[rI];
    // this item's nodes: X = k c0 (vertex-type along x) and k c0 + a, a = 1..k-1
    const int nNodes = c0 < (unsigned)g.N[0] ? k : 1;
    for (int a = 0; a < nNodes; ++a) {
      const u64 jdx = a == 0 ? r.rowV + (u64)r.blkV * (u64)c0 : r.rowE + (u64)r.blkE * (u64)c0 + (u64)(a - 1);
      double y[DFX_MAX_LEAVES];
      if (MODE == 0) {
        int ie = g.off[0] + (int)c0, aa = a;
        if (a == 0 && ie >= g.G[0]) { ie = g.G[0] - 1; aa = k; }      // last lattice plane: owner on the left
        double x[3];
        x[0] = nodeCoord(g, k, 0, ie, aa); x[1] = r.u1; x[2] = r.u2;
        f(x, y);
      } else {
        // (t_x * t_y) * t_z: the order in which the reference's callable multiplies its factors
        y[0] = __dmul_rn(__dmul_rn(tab[A.tabOff[0] + k * (int)c0 + a], r.u1), r.u2);
      }
      for (int l = 0; l < A.nLeaves; ++l) {
        const DevLeaf& leaf = B.leaves[A.firstLeaf + l];
        if (leaf.layout != A.layoutId) continue;
        const u64 p = leaf.posA * jdx + leaf.posB;
        if (mask == nullptr || mask[p]) coeff[p] = scalar ? y[0] : y[l];
      }
    }
  }
}

struct CellArgs {
  int layoutId, firstLeaf, nLeaves;
  FastDiv divNloc, divN0, divN1, divK1;
  u64 total;
  int tabOff[3];
};

template <class F, int MODE>
__global__ void __launch_bounds__(256)
k_interp_cell(const __grid_constant__ DevBasis B, const __grid_constant__ CellArgs A, const F f,
              const double* __restrict__ tab, double* __restrict__ coeff, const uint8_t* __restrict__ mask) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.total) return;
  const DevGrid& g = B.grid;
  const DevLayout& L = B.layouts[A.layoutId];
  const int k = L.k;
  unsigned e, i;                              // element; its DOFs are t itself: e*nloc + i
  A.divNloc.divmod((unsigned)t, e, i);
  int ec[3], a[3];
  {
    unsigned q, r, q2;
    A.divN0.divmod(e, q, r); ec[0] = (int)r;
    A.divN1.divmod(q, q2, r); ec[1] = (int)r; ec[2] = (int)q2;
    A.divK1.divmod(i, q, r); a[0] = (int)r;
    A.divK1.divmod(q, q2, r); a[1] = (int)r; a[2] = (int)q2;
  }
  double x[3] = {0, 0, 0};
  double prod = 1.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j >= g.dim) continue;
    if (MODE == 0) x[j] = nodeCoord(g, k, j, g.off[j] + ec[j], a[j]);
    else prod = __dmul_rn(prod, tab[A.tabOff[j] + ec[j] * (k + 1) + a[j]]);
  }
  double y[DFX_MAX_LEAVES];
  if (MODE == 0) f(x, y); else y[0] = prod;
  const bool scalar = MODE == 1 || f.ncomp() == 1;
  for (int l = 0; l < A.nLeaves; ++l) {
    const DevLeaf& leaf = B.leaves[A.firstLeaf + l];
    if (leaf.layout != A.layoutId) continue;
    const u64 p = leaf.posA * t + leaf.posB;
    if (mask == nullptr || mask[p]) coeff[p] = scalar ? y[0] : y[l];
  }
}

static bool separable(const dfx_function& f) {
  return f.n_comp == 1 && (f.id == DFX_FN_GAUSSIAN || f.id == DFX_FN_SINPROD || f.id == DFX_FN_COORD ||
                           f.id == DFX_FN_CONSTANT);
}

int g_interpMode = -1;   // -1 automatic, 0 per-node evaluation of f

int interpolateFast(dfx_basis* b, int firstLeaf, int nLeaves, const dfx_function& fn, double* coeff,
                    const uint8_t* mask, cudaStream_t st) {
  const DevBasis& D = b->dev;
  const DevGrid& g = D.grid;
  RegistryFn f; f.f = fn; f.dim = g.dim;
  const bool sep = separable(fn) && g_interpMode != 0;
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
    double* tab = nullptr;
    if (sep) {
      const size_t need = (size_t)tabN * sizeof(double);
      if (b->dSepBytes < need) {
        if (b->dSep) cudaFree(b->dSep);
        b->dSep = nullptr; b->dSepBytes = 0;
        DFX_CUDA(cudaMalloc(&b->dSep, need));
        b->dSepBytes = need;
      }
      tab = (double*)b->dSep;
      k_sep_tables<<<(tabN + 255) / 256, 256, 0, st>>>(g, fn, A, k, cell ? 1 : 0, tab);
      int rc = postLaunch("k_sep_tables");
      if (rc) return rc;
    }
    if (!cell) {
      const u64 nRows = (u64)A.LX[1] * A.LX[2];
      if (nRows > 0xffffffffull) return fail(DFX_ERR_RANGE, "lattice too large for one launch");
      A.divLY = FastDiv((unsigned)A.LX[1]); A.divEnt = FastDiv((unsigned)g.N[0] + 1u);
      const unsigned grid = (unsigned)((nRows + kRowsPerBlock - 1) / kRowsPerBlock);
      if (sep) k_interp_rows<RegistryFn, 1><<<grid, 256, 0, st>>>(D, A, f, tab, coeff, mask);
      else k_interp_rows<RegistryFn, 0><<<grid, 256, 0, st>>>(D, A, f, tab, coeff, mask);
      int rc = postLaunch("k_interp_rows");
      if (rc) return rc;
    } else {
      CellArgs C;
      C.layoutId = q; C.firstLeaf = firstLeaf; C.nLeaves = nLeaves;
      C.divNloc = FastDiv((unsigned)L.nloc); C.divN0 = FastDiv((unsigned)g.N[0]); C.divN1 = FastDiv((unsigned)g.N[1]);
      C.divK1 = FastDiv((unsigned)(k + 1));
      C.total = L.dimension;
      for (int j = 0; j < 3; ++j) C.tabOff[j] = A.tabOff[j];
      const u64 blocks = (C.total + 255) / 256;
      if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "basis too large for one launch");
      if (sep) k_interp_cell<RegistryFn, 1><<<(unsigned)blocks, 256, 0, st>>>(D, C, f, tab, coeff, mask);
      else k_interp_cell<RegistryFn, 0><<<(unsigned)blocks, 256, 0, st>>>(D, C, f, tab, coeff, mask);
      int rc = postLaunch("k_interp_cell");
      if (rc) return rc;
    }
  }
  return DFX_OK;
}

// ---------------------------------------------------------------------------
// Index table: one thread per (element, local index); the host guarantees < 2^32 outputs.
struct IdxArgs {
  u64 e0, total;
  FastDiv divNloc, divN0, divN1;
  int stride;
};

// per local index: leaf (bits 0-4), alpha_0/1/2 (bits 5-9, 10-14, 15-19)
template <class T, bool MULTI>
__global__ void __launch_bounds__(256)
k_indices_fast(const __grid_constant__ DevBasis B, const __grid_constant__ IdxArgs A,
               const unsigned* __restrict__ localTab, T* __restrict__ out) {
  const u64 t = (u64)blockIdx.x * 256 + threadIdx.x;
  if (t >= A.total) return;
  unsigned dq, li;
  A.divNloc.divmod((unsigned)t, dq, li);
  const unsigned e = (unsigned)A.e0 + dq;
  int ec[3];
  {
    unsigned q, r, q2;
    A.divN0.divmod(e, q, r); ec[0] = (int)r;
    A.divN1.divmod(q, q2, r); ec[1] = (int)r; ec[2] = (int)q2;
  }
  const unsigned packed = __ldg(localTab + li);
  const DevLeaf& leaf = B.leaves[packed & 31u];
  const DevLayout& L = B.layouts[leaf.layout];
  int a[3] = {(int)((packed >> 5) & 31u), (int)((packed >> 10) & 31u), (int)((packed >> 15) & 31u)};
  const u64 j = layoutIndex(L, B.grid, ec, a, (int)li - leaf.localOffset);
  if (MULTI) {
    for (int p = 0; p < A.stride; ++p)
      out[t * (u64)A.stride + p] = p < leaf.ndig ? (T)(leaf.digA[p] * j + leaf.digB[p]) : (T)0;
  } else {
    out[t] = (T)(leaf.posA * j + leaf.posB);
  }
}

template <class T, bool MULTI>
int indicesFast(const dfx_basis* b, u64 e0, u64 e1, T* out, int stride, cudaStream_t st) {
  IdxArgs A;
  A.e0 = e0; A.total = (e1 - e0) * (u64)b->dev.nlocTotal; A.stride = stride;
  if (A.total == 0) return DFX_OK;
  A.divNloc = FastDiv((unsigned)b->dev.nlocTotal);
  A.divN0 = FastDiv((unsigned)b->dev.grid.N[0]); A.divN1 = FastDiv((unsigned)b->dev.grid.N[1]);
  const u64 blocks = (A.total + 255) / 256;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  k_indices_fast<T, MULTI><<<(unsigned)blocks, 256, 0, st>>>(b->dev, A, (const unsigned*)b->dLocalTab, out);
  return postLaunch("k_indices_fast");
}

template int indicesFast<uint32_t, true>(const dfx_basis*, u64, u64, uint32_t*, int, cudaStream_t);
template int indicesFast<u64, true>(const dfx_basis*, u64, u64, u64*, int, cudaStream_t);
template int indicesFast<uint32_t, false>(const dfx_basis*, u64, u64, uint32_t*, int, cudaStream_t);
template int indicesFast<u64, false>(const dfx_basis*, u64, u64, u64*, int, cudaStream_t);

}  // namespace dfx
