// dfx_eval_dmma.cu — path 3: dense evaluation on the FP64 tensor cores.
//
// For an arbitrary (non tensor-product) point set the contraction of
// LocalFunction::operator() (discreteglobalbasisfunction.hh:358-363) and of the derivative
// (:607-612) is a GEMM per leaf:
//     C[element][n] = sum_i  coeff[element][i] * Bt[i][n],
// n = q (values) followed by q*dim + d (reference Jacobians), Bt = shape-function table of
// the points.  There is no FP64 kind of tcgen05.mma; the FP64 tensor path of sm_100a is
// mma.sync.aligned.m8n8k4 (SASS DMMA).  A CTA gathers the coefficients of 128 elements into
// shared memory (rows padded to a stride = 4 mod 16 doubles: conflict-free fragment loads),
// streams the table through shared memory in blocks of 64 columns and lets each of its
// 8 warps accumulate a 16 x 64 tile in registers.
//
// For tensor-product points the sum-factorised kernels do 4 (values) to 11 (with gradients)
// times fewer flops and win by a wide margin (profiles/); this path is
// selected only when they do not apply.
#include <algorithm>
#include <cstring>

#include "dfx_device.cuh"
#include "dfx_eval_fast.h"
#include "dfx_launch.h"

namespace dfx {

constexpr int kDmmaElems = 128;    // elements (GEMM rows) per CTA
constexpr int kDmmaCols = 64;      // table columns per pass
constexpr int kDmmaLdb = kDmmaCols + 4;

struct DmmaArgs {
  int firstLeaf, nGroup, compOffset, ncompTotal;
  int nloc, kp, lda;               // local DOFs, padded to a multiple of 4, smem row stride
  int nq, ncols, ncolsPad, dim;    // points, nq*(1+dim) (or nq), padded to a multiple of 64
  int wantVal, wantJac;
  u64 e0, ne;
  FastDiv divN0, divN1, divDim;
};

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct GatherLi { u64 P0, S; unsigned cs0, cs1; };     // position of local index i in element e: P0 + S * lex(e; cs0, cs1)

__global__ void __launch_bounds__(256)
k_eval_dmma(const DevBasis* __restrict__ Bp, const __grid_constant__ DmmaArgs R, const unsigned* __restrict__ localTab,
            const double* __restrict__ coeff, const double* __restrict__ Bt, double* __restrict__ values,
            double* __restrict__ jac) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                                   // [128][lda]
  double* Bs = smem + (size_t)kDmmaElems * R.lda;      // [kp][68]
  GatherLi* li = reinterpret_cast<GatherLi*>(Bs + (size_t)R.kp * kDmmaLdb);    // [kp]
  const DevGrid& g = Bp->grid;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tig = lane & 3;
  const u64 el0 = (u64)blockIdx.x * kDmmaElems;
  const unsigned nEl = (R.ne - el0) < (u64)kDmmaElems ? (unsigned)(R.ne - el0) : (unsigned)kDmmaElems;
  const DevLeaf& leaf0 = Bp->leaves[R.firstLeaf];
  const DevLayout& L = Bp->layouts[leaf0.layout];

  // the two rows of this thread's accumulator fragments: element, diag(J^-1) (AxisAlignedCubeGeometry), once per kernel
  // — the first version recomputed the three divisions in every pass of every leaf (31 % of the stall samples sat in
  // the epilogue)
  double jinvRow[2][3] = {{1, 1, 1}, {1, 1, 1}};
  if (R.wantJac) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const unsigned n = (unsigned)(warp * 16 + mt * 8 + grp);
      if (n >= nEl) continue;
      unsigned q, ec0, ec1, ec2;
      R.divN0.divmod((unsigned)(R.e0 + el0) + n, q, ec0);
      R.divN1.divmod(q, ec2, ec1);
      jinvRow[mt][0] = elementJacInv(g, 0, g.off[0] + (int)ec0);
      if (R.dim > 1) jinvRow[mt][1] = elementJacInv(g, 1, g.off[1] + (int)ec1);
      if (R.dim > 2) jinvRow[mt][2] = elementJacInv(g, 2, g.off[2] + (int)ec2);
    }
  }
  for (int gI = 0; gI < R.nGroup; ++gI) {
    const DevLeaf& leaf = Bp->leaves[R.firstLeaf + gI];
    __syncthreads();
    // ---- gather (LocalFunctionBase::bind): the per-local-index constants of the closed-form position go into shared
    // memory once, then the 128 x kp entries of the tile are spread over ALL threads as independent loads (the first
    // version let thread i walk its 128 elements one dependent iteration after the other: 125 threads, one load in
    // flight each — about 40 us of the CTA's 217 us) ----
    for (int i = tid; i < R.kp; i += blockDim.x) {
      GatherLi c; c.P0 = 0; c.S = 0; c.cs0 = 0; c.cs1 = 0;
      if (i < R.nloc) {
        const unsigned packed = __ldg(localTab + leaf0.localOffset + i);
        const int a[3] = {(int)((packed >> 5) & 31u), (int)((packed >> 10) & 31u), (int)((packed >> 15) & 31u)};
        const int zero[3] = {0, 0, 0};
        const u64 jbase = layoutIndex(L, g, zero, a, i);
        unsigned s = 0;
        if (L.kind == DFX_NODE_LAGRANGE_DG || L.k == 0) s = (1u << g.dim) - 1u;
        else
          for (int j = 0; j < 3; ++j) if (a[j] > 0 && a[j] < L.k) s |= 1u << j;
        c.cs0 = (unsigned)L.csize[s][0]; c.cs1 = (unsigned)L.csize[s][1];
        c.P0 = leaf.posA * jbase + leaf.posB; c.S = leaf.posA * (u64)L.blk[s];
      }
      li[i] = c;
    }
    __syncthreads();
    // eight independent loads in flight per thread (ncu: with one, 16 % of all stall samples sat on the store behind it)
    for (int idx0 = tid; idx0 < kDmmaElems * R.kp; idx0 += 8 * 256) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = idx0 + u * 256;
        const unsigned n = (unsigned)idx / (unsigned)R.kp, i = (unsigned)idx - n * (unsigned)R.kp;
        v[u] = 0.0;
        if (idx < kDmmaElems * R.kp && n < nEl && (int)i < R.nloc) {
          unsigned q, ec0, ec1, ec2;                     // element coordinates of the CTA's n-th element (x fastest)
          R.divN0.divmod((unsigned)(R.e0 + el0) + n, q, ec0);
          R.divN1.divmod(q, ec2, ec1);
          const GatherLi c = li[i];
          v[u] = __ldg(coeff + c.P0 + c.S * (u64)(ec0 + c.cs0 * (ec1 + c.cs1 * ec2)));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = idx0 + u * 256;
        if (idx < kDmmaElems * R.kp) {
          const unsigned n = (unsigned)idx / (unsigned)R.kp, i = (unsigned)idx - n * (unsigned)R.kp;
          As[n * R.lda + i] = v[u];
        }
      }
    }
    // ---- column blocks of the table ----
    for (int cb = 0; cb < R.ncolsPad; cb += kDmmaCols) {
      __syncthreads();
      // 16-byte loads, eight in flight per thread (rows of the table and of Bs are 16-byte aligned: ncolsPad, cb and
      // kDmmaLdb are even); one 8-byte load at a time left the table's L2 latency exposed in every pass (81 -> 63 ms).
      // Two 32-column buffers filled by cp.async under the previous pass were measured too: 65.6 ms, no better.
      {
        constexpr int kVec = kDmmaCols / 2;
        const int total = R.kp * kVec;
#pragma unroll 8
        for (int idx = tid; idx < total; idx += 256) {
          const int kk = idx / kVec, nn = idx - kk * kVec;
          const double2 v = __ldg(reinterpret_cast<const double2*>(Bt + (size_t)kk * R.ncolsPad + cb) + nn);
          *reinterpret_cast<double2*>(Bs + kk * kDmmaLdb + 2 * nn) = v;
        }
      }
      __syncthreads();
      constexpr int NT = kDmmaCols / 8;
      double acc[2][NT][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
      const int r0 = warp * 16;
      for (int ks = 0; ks < R.kp; ks += 4) {
        const double a0 = As[(r0 + grp) * R.lda + ks + tig];
        const double a1 = As[(r0 + 8 + grp) * R.lda + ks + tig];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double bv = Bs[(ks + tig) * kDmmaLdb + nt * 8 + grp];
          dmma(acc[0][nt][0], acc[0][nt][1], a0, bv);
          dmma(acc[1][nt][0], acc[1][nt][1], a1, bv);
        }
      }
      // ---- epilogue: C[row = r0 + mt*8 + grp][col = cb + nt*8 + 2 tig + {0,1}] ----
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const unsigned n = (unsigned)(r0 + mt * 8 + grp);
        if (n >= nEl) continue;
        const u64 el = el0 + n;
        const double jinv[3] = {jinvRow[mt][0], jinvRow[mt][1], jinvRow[mt][2]};
        // row bases once per row; with one range entry per point (the scalar case) a column is a plain offset
        const unsigned nT = (unsigned)R.ncompTotal, nvalCols = R.wantVal ? (unsigned)R.nq : 0u;
        double* rowV = values ? values + (el * (u64)R.nq * nT + (u64)(R.compOffset + gI)) : nullptr;
        double* rowJ = jac ? jac + (el * (u64)R.nq * nT + (u64)(R.compOffset + gI)) * (u64)R.dim : nullptr;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const unsigned col = (unsigned)(cb + nt * 8 + 2 * tig + h);
            if (col >= (unsigned)R.ncols) continue;
            const double v = acc[mt][nt][h];
            if (col < nvalCols) {
              rowV[col * nT] = v;
            } else {
              unsigned q, d;
              R.divDim.divmod(col - nvalCols, q, d);
              rowJ[nT == 1u ? col - nvalCols : (q * nT) * (unsigned)R.dim + d] = v * (d == 0 ? jinv[0] : (d == 1 ? jinv[1] : jinv[2]));
            }
          }
      }
    }
  }
}

bool dmmaPathHas(const dfx_basis* b, int firstLeaf, int nLeaves) {
  const DevBasis& D = b->dev;
  const int k = D.layouts[D.leaves[firstLeaf].layout].k;
  for (int l = 0; l < nLeaves; ++l)
    if (D.layouts[D.leaves[firstLeaf + l].layout].k != k) return false;     // one table for the subtree
  const int nloc = D.layouts[D.leaves[firstLeaf].layout].nloc;
  return nloc >= 16 && nloc <= 128;
}

// shape table of the points, column n = q (values) then q*dim + d (reference Jacobians)
int runDmmaEval(dfx_basis* b, int node, int firstLeaf, int nLeaves, const double* xhat, int nq, const double* coeff,
                u64 e0, u64 e1, double* values, double* jac, cudaStream_t st) {
  if (e1 == e0) return DFX_OK;
  const DevBasis& D = b->dev;
  const int dim = D.grid.dim;
  const DevLayout& L0 = D.layouts[D.leaves[firstLeaf].layout];
  const int k = L0.k, nloc = L0.nloc;
  DmmaArgs R;
  R.nloc = nloc; R.kp = (nloc + 3) & ~3; R.lda = R.kp + 4;
  R.nq = nq; R.dim = dim; R.wantVal = values ? 1 : 0; R.wantJac = jac ? 1 : 0;
  R.ncols = (values ? nq : 0) + (jac ? nq * dim : 0);
  R.ncolsPad = (R.ncols + kDmmaCols - 1) / kDmmaCols * kDmmaCols;
  R.e0 = e0; R.ne = e1 - e0;
  R.divN0 = FastDiv((unsigned)D.grid.N[0]); R.divN1 = FastDiv((unsigned)D.grid.N[1]); R.divDim = FastDiv((unsigned)dim);
  // table (cached per point set / request shape)
  const size_t tabN = (size_t)R.kp * R.ncolsPad;
  const bool cached = b->dDenseNode == node && b->dDenseCols == R.ncols && b->dDenseVal == R.wantVal &&
                      (int)b->denseXhat.size() == nq * dim &&
                      std::memcmp(b->denseXhat.data(), xhat, sizeof(double) * nq * dim) == 0 && b->dDense;
  if (!cached) {
    std::vector<double> h(tabN, 0.0);
    for (int q = 0; q < nq; ++q) {
      const double* x = xhat + (size_t)q * dim;
      for (int i = 0; i < nloc; ++i) {
        int a[3] = {0, 0, 0};
        { int ii = i; for (int j = 0; j < dim; ++j) { a[j] = ii % (k + 1); ii /= (k + 1); } }
        double v = 1.0, dv[3] = {1.0, 1.0, 1.0};
        if (k == 0) { v = 1; dv[0] = dv[1] = dv[2] = 0; }
        else if (k == 1) {
          for (int j = 0; j < dim; ++j) v *= (i & (1 << j)) ? x[j] : 1 - x[j];
          for (int d = 0; d < dim; ++d) {
            dv[d] = (i & (1 << d)) ? 1 : -1;
            for (int j = 0; j < dim; ++j) if (j != d) dv[d] *= (i & (1 << j)) ? x[j] : 1 - x[j];
          }
        } else {
          for (int j = 0; j < dim; ++j) v *= lagrangeP(k, a[j], x[j]);
          for (int d = 0; d < dim; ++d)
            for (int j = 0; j < dim; ++j) dv[d] *= (j == d) ? lagrangeDP(k, a[j], x[j]) : lagrangeP(k, a[j], x[j]);
        }
        int col = 0;
        if (values) { h[(size_t)i * R.ncolsPad + q] = v; col = nq; }
        if (jac) for (int d = 0; d < dim; ++d) h[(size_t)i * R.ncolsPad + col + q * dim + d] = dv[d];
      }
    }
    if (b->dDenseBytes < tabN * sizeof(double)) {
      if (b->dDense) cudaFree(b->dDense);
      b->dDense = nullptr; b->dDenseBytes = 0;
      DFX_CUDA(cudaMalloc(&b->dDense, tabN * sizeof(double)));
      b->dDenseBytes = tabN * sizeof(double);
    }
    DFX_CUDA(cudaMemcpyAsync(b->dDense, h.data(), tabN * sizeof(double), cudaMemcpyHostToDevice, st));
    DFX_CUDA(cudaStreamSynchronize(st));
    b->denseXhat.assign(xhat, xhat + (size_t)nq * dim);
    b->dDenseNode = node; b->dDenseCols = R.ncols; b->dDenseVal = R.wantVal;
  }
  const size_t smem = ((size_t)kDmmaElems * R.lda + (size_t)R.kp * kDmmaLdb) * sizeof(double) + (size_t)R.kp * sizeof(GatherLi);
  static std::atomic<unsigned long long> configured{0};
  if (smem > 48 * 1024) { int rc = optInSharedMemory(k_eval_dmma, b->device, configured); if (rc) return rc; }
  const u64 blocks = (R.ne + kDmmaElems - 1) / kDmmaElems;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  // runs of leaves sharing a layout (the gather constants are per layout)
  int l = 0;
  while (l < nLeaves) {
    int r = l + 1;
    while (r < nLeaves && D.leaves[firstLeaf + r].layout == D.leaves[firstLeaf + l].layout) ++r;
    R.firstLeaf = firstLeaf + l; R.nGroup = r - l; R.compOffset = l; R.ncompTotal = nLeaves;
    k_eval_dmma<<<(unsigned)blocks, 256, smem, st>>>((const DevBasis*)b->dDev, R, (const unsigned*)b->dLocalTab, coeff,
                                                     (const double*)b->dDense, values, jac);
    int rc = postLaunch("k_eval_dmma");
    if (rc) return rc;
    l = r;
  }
  return DFX_OK;
}

}  // namespace dfx
