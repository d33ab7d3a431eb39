// dfx_eval_fast.cu — specialised evaluation kernels for tensor-product point sets.
//
// Path 1 (k_eval_reg): one thread per element, x-blocked (consecutive threads = consecutive
// elements in x, so every coefficient load of a warp is a contiguous run inside one
// index-set component); the (k+1)^d coefficients live in registers and are contracted
// direction by direction (sum factorisation: d*(k+1)*m^d-ish FMAs instead of (k+1)^d*m^d);
// the m^d results of the block's elements are staged in shared memory in output order and
// leave the SM as ONE contiguous TMA bulk store (cp.async.bulk.global.shared::cta).
//
// Replaces LocalFunctionBase::bind + LocalFunction::operator() and the derivative's
// operator() (discreteglobalbasisfunction.hh:124-148, :336-371, :583-627).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "dfx_device.cuh"
#include "dfx_eval_fast.h"
#include "dfx_sumfactor.cuh"
#include "dfx_launch.h"

namespace dfx {

// ---------------------------------------------------------------------------
// planner
static bool tensorFactor(const double* xhat, int nq, int dim, int m, bool firstFastest, double x1[3][kMaxQ1]) {
  int stride[3];
  if (firstFastest) { stride[0] = 1; stride[1] = m; stride[2] = m * m; }
  else { int s = 1; for (int j = dim - 1; j >= 0; --j) { stride[j] = s; s *= m; } }
  for (int j = 0; j < dim; ++j)
    for (int t = 0; t < m; ++t) x1[j][t] = xhat[(size_t)(t * stride[j]) * dim + j];
  for (int q = 0; q < nq; ++q)
    for (int j = 0; j < dim; ++j) {
      const int t = (q / stride[j]) % m;
      if (std::memcmp(&xhat[(size_t)q * dim + j], &x1[j][t], sizeof(double)) != 0) return false;
    }
  return true;
}

static bool regPathHas(int dim, int k, int m) {
  return (dim == 2 || dim == 3) && (k == 1 || k == 2) && (m == 2 || m == 3);
}

int g_evalBlock = 32;      // elements per block of the values-only register kernel (DFX_OP_TUNE_EVAL_BLOCK): 32 = 20 single-warp blocks per SM
                           // at 96 registers (32 bytes of spill; no block barrier worth the name, one 6.9 KB tile per warp), 64 = 10
                           // blocks of 2 warps, 128 = 6 blocks of 4 warps at 80 registers (64 / 72 bytes of spill, whose reloads ncu's
                           // source view charged 11 % of the stall samples to, and 14 % to the one warp per block that waits for the
                           // bulk store before it exits).  Measured back to back: 128 -> 64: 512^3 6.53 -> 6.05 and 7.00 -> 6.58 ms;
                           // 64 -> 32: 256^3 0.740 -> 0.731 ms (three of three rounds), 512^3 6.12 -> 5.97, 6.64 -> 6.52, 6.93 -> 6.71 ms
                           // (seven of eight rounds, profiles/r02_evalblock32_ab.jsonl; the GPU's thermal drift inside one run is larger)
int g_evalStreamHint = 1;  // bulk stores of results with the L2 evict-first policy (DFX_OP_TUNE_STREAM_HINT)
int g_evalSplitByGroup = 1;   // DFX_OP_TUNE_SPLIT_MAP: fused Taylor-Hood evaluation with whole warps per leaf group (0: every warp runs both
                              // branches; measured 0.441 -> 0.370 ms on Taylor-Hood 128^3, profiles/r02_c3eval_split_map_ab.jsonl)
int g_evalSwizzle = 1;     // dfx_set_kernel_path(DFX_OP_TUNE_SWIZZLE, .): 0 blocks in grid-view order, 1 strip schedule when a
                           // z-layer outgrows the L2, 2 strip schedule whenever the shape allows (tests)

int planFastEval(const dfx_basis* b, int firstLeaf, int nLeaves, const double* xhat, int nq, bool wantJac, int forced,
                 FastPlan& plan) {
  if (forced == 0) return 0;
  const DevBasis& D = b->dev;
  const int dim = D.grid.dim;
  int m = 0;
  for (int t = 1; t <= kMaxQ1; ++t) {
    int p = 1; for (int j = 0; j < dim; ++j) p *= t;
    if (p == nq) m = t;
  }
  if (m == 0) return 0;
  double x1[3][kMaxQ1];
  bool ff = true;
  if (!tensorFactor(xhat, nq, dim, m, true, x1)) {
    ff = false;
    if (!tensorFactor(xhat, nq, dim, m, false, x1)) return 0;
  }
  // every leaf of the subtree must have the same order to share the 1-d tables
  const int k = D.layouts[D.leaves[firstLeaf].layout].k;
  for (int l = 0; l < nLeaves; ++l)
    if (D.layouts[D.leaves[firstLeaf + l].layout].k != k) return 0;
  if (k + 1 > kMaxN1) return 0;
  int path = 0;
  if (regPathHas(dim, k, m)) path = 1;
  else if (smemPathHas(dim, k, m)) path = 2;
  if (forced == 2 && smemPathHas(dim, k, m)) path = 2;     // both apply for k = 2: testing / comparison
  if (forced > 0 && forced != path) return 0;
  if (path == 0) return 0;
  plan.path = path; plan.dim = dim; plan.k = k; plan.m = m; plan.firstFastest = ff; plan.jac = wantJac; plan.nq = nq;
  for (int j = 0; j < dim; ++j)
    for (int t = 0; t < m; ++t)
      for (int a = 0; a <= k; ++a) {
        // LagrangeCubeLocalBasis: k = 1 uses (1-x, x); general k the product formula
        if (k == 1) { plan.A[j][t][a] = a ? x1[j][t] : 1 - x1[j][t]; plan.Dm[j][t][a] = a ? 1.0 : -1.0; }
        else { plan.A[j][t][a] = lagrangeP(k, a, x1[j][t]); plan.Dm[j][t][a] = lagrangeDP(k, a, x1[j][t]); }
      }
  // collocation derivative matrices (barycentric form)
  plan.colloc = true;
  for (int j = 0; j < dim; ++j) {
    double w[kMaxQ1];
    for (int q = 0; q < m; ++q) {
      w[q] = 1.0;
      for (int r = 0; r < m; ++r)
        if (r != q) {
          const double d = x1[j][q] - x1[j][r];
          if (std::abs(d) < 1e-9) plan.colloc = false;
          w[q] *= d;
        }
    }
    if (!plan.colloc) break;
    for (int q = 0; q < m; ++q) {
      double diag = 0.0;
      for (int r = 0; r < m; ++r) {
        if (r == q) continue;
        plan.Cm[j][q][r] = (w[q] / w[r]) / (x1[j][q] - x1[j][r]);
        diag += 1.0 / (x1[j][q] - x1[j][r]);
      }
      plan.Cm[j][q][q] = diag;
    }
  }
  return path;
}

// ---------------------------------------------------------------------------
template <int N1, int M>
struct Tables {
  double A[3][M][N1];
  double D[3][M][N1];
};

struct RegArgs {
  int firstLeaf, nGroup;       // leaves [firstLeaf, firstLeaf+nGroup) share one layout
  int compOffset, ncompTotal;  // where the group's range entries sit in the output
  int qs[3];                   // stride of q_j in the point index
  unsigned long long e0, ne;
  FastDiv divN0, divN1;
  int bulk;                    // 1: the block's output tile is contiguous and 16-byte aligned
  unsigned long long posA[DFX_MAX_LEAVES], posB[DFX_MAX_LEAVES];   // flat position map of the group's leaves
  // Block schedule (swz = 1): blocks do not walk the elements in grid-view order (x, then y, then z) but strip
  // by strip — a strip is `stripH` element rows in y, taken through ALL z-layers before the next strip starts.
  // The coefficients two z-neighbours share are then re-used after one strip-layer (a few MB) instead of after
  // a whole z-layer of elements: at 512^3 that layer is 73 MB of coefficients + results, more than the L2 keeps,
  // and ncu counted 11.3 GB of DRAM reads for 8.6 GB of coefficients (profiles/r02_ncu_configs.csv).  Every block
  // still writes its own contiguous run of the element-major output.
  int swz;
  unsigned n0, n1, stripH, fullStrips;
  FastDiv divBpr, divStripVol, divStripH, divLastH;
};

__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// contiguous shared -> global copy by the TMA engine (SASS: UBLKCP)
__device__ __forceinline__ void bulkStore(double* gdst, const double* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smemAddr(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the same with an L2 evict-first policy: results are written once and never read by this kernel, so they
// should not push the coefficients that neighbouring elements still need out of the L2
__device__ __forceinline__ void bulkStoreStreaming(double* gdst, const double* ssrc, unsigned bytes) {
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smemAddr(ssrc)),
               "r"(bytes), "l"(policy)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulkStoreWaitRead() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// LocalFunctionBase::bind for one leaf (discreteglobalbasisfunction.hh:140-147): the (KK+1)^DIM coefficients of
// element e, located by the closed form of MCMGMapper::subIndex on the structured grid
template <int DIM, int KK>
__device__ __forceinline__ void gatherLeaf(const DevLayout& L, const double* __restrict__ coeff, u64 posA, u64 posB, u64 e,
                                           const int (&ec)[3], double* __restrict__ c) {
  constexpr int N1 = KK + 1;
  constexpr int NLOC = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  if (L.kind == DFX_NODE_LAGRANGE_DG) {
    const u64 base = posA * (e * (u64)NLOC) + posB;
#pragma unroll
    for (int i = 0; i < NLOC; ++i) c[i] = __ldg(coeff + base + posA * (u64)i);
    return;
  }
#pragma unroll
  for (int a2 = 0; a2 < (DIM == 3 ? N1 : 1); ++a2)
#pragma unroll
    for (int a1 = 0; a1 < N1; ++a1)
#pragma unroll
      for (int a0 = 0; a0 < N1; ++a0) {
        const int s = ((a0 > 0 && a0 < KK) ? 1 : 0) | ((a1 > 0 && a1 < KK) ? 2 : 0) | ((DIM == 3 && a2 > 0 && a2 < KK) ? 4 : 0);
        const int c0 = ec[0] + (a0 == KK ? 1 : 0);
        const int c1 = ec[1] + (a1 == KK ? 1 : 0);
        const int c2 = ec[2] + ((DIM == 3 && a2 == KK) ? 1 : 0);
        // KK <= 2 here: at most one interior node per direction, so the in-entity index is 0
        // (blk[s] = (KK-1)^|s| = 1); entity counts of one component stay below 2^32
        const unsigned ent = (unsigned)c0 + (unsigned)L.csize[s][0] * ((unsigned)c1 + (unsigned)L.csize[s][1] * (unsigned)c2);
        const u64 j = L.compOffset[s] + (u64)ent;
        c[a0 + N1 * (a1 + N1 * a2)] = __ldg(coeff + posA * j + posB);
      }
}

// resident blocks per SM the register allocator is asked to allow (values-only kernels):
// 6 x 128 threads = 24 warps at <= 80 registers, no spills for any instantiated (DIM, K, M)
#ifndef DFX_EVAL_MINB
#define DFX_EVAL_MINB 6
#endif

template <int DIM, int K, int M, bool JAC, int EPB>
__global__ void __launch_bounds__(EPB, JAC ? 1 : (EPB == 64 ? 10 : (EPB == 32 ? 20 : DFX_EVAL_MINB)))
k_eval_reg(const __grid_constant__ DevGrid g, const __grid_constant__ DevLayout L,
           const __grid_constant__ Tables<K + 1, M> T, const __grid_constant__ RegArgs R,
           const double* __restrict__ coeff, double* __restrict__ values, double* __restrict__ jac) {
  constexpr int N1 = K + 1;
  constexpr int NLOC = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  constexpr int NQ = DIM == 3 ? M * M * M : M * M;
  extern __shared__ __align__(128) double stage[];       // values tile [EPB][NQ][nGroup], then jac tile
  // g and L (the group's DOF lattice) are direct kernel parameters: every access below has a
  // compile-time offset into the constant bank (no indexed-constant loads in the hot path)
  const int tid = threadIdx.x;
  u64 blk0 = (u64)blockIdx.x * EPB;
  if (R.swz) {
    unsigned r, xb, strip, rem, z, yy;
    R.divBpr.divmod(blockIdx.x, r, xb);                  // r: row slot in strip order, xb: block inside the row
    R.divStripVol.divmod(r, strip, rem);                 // strip volume = stripH rows x all layers of the range
    if (strip < R.fullStrips) R.divStripH.divmod(rem, z, yy);
    else R.divLastH.divmod(rem, z, yy);                  // the last, lower strip
    blk0 = ((u64)z * R.n1 + (u64)strip * R.stripH + yy) * (u64)R.n0 + (u64)xb * EPB;
  }
  const u64 el = blk0 + tid;
  const bool active = el < R.ne;
  const u64 e = R.e0 + (active ? el : 0);
  // element coordinates (x fastest)
  int ec[3];
  {
    const unsigned e32 = (unsigned)e;       // numElements < 2^32 checked on the host
    unsigned q, r;
    R.divN0.divmod(e32, q, r);
    ec[0] = (int)r;
    unsigned q2;
    R.divN1.divmod(q, q2, r);
    ec[1] = (int)r; ec[2] = (int)q2;
  }
  const int nG = R.nGroup;
  const int rowLen = NQ * nG;
  double* vTile = stage;
  double* jTile = stage + (size_t)EPB * rowLen;           // [EPB][NQ][nGroup][DIM]
  double jinv[3] = {0, 0, 0};
  if (JAC) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) jinv[d] = elementJacInv(g, d, g.off[d] + ec[d]);
  }

  for (int gI = 0; gI < nG; ++gI) {
    const u64 posA = R.posA[gI], posB = R.posB[gI];
    double c[NLOC];
    if (L.kind == DFX_NODE_LAGRANGE_DG) {
      const u64 base = posA * (e * (u64)NLOC) + posB;
#pragma unroll
      for (int i = 0; i < NLOC; ++i) c[i] = __ldg(coeff + base + posA * (u64)i);
    } else {
#pragma unroll
      for (int a2 = 0; a2 < (DIM == 3 ? N1 : 1); ++a2)
#pragma unroll
        for (int a1 = 0; a1 < N1; ++a1)
#pragma unroll
          for (int a0 = 0; a0 < N1; ++a0) {
            // closed form of MCMGMapper::subIndex on the structured grid
            constexpr int dummy = 0; (void)dummy;
            const int s = ((a0 > 0 && a0 < K) ? 1 : 0) | ((a1 > 0 && a1 < K) ? 2 : 0) | ((DIM == 3 && a2 > 0 && a2 < K) ? 4 : 0);
            const int c0 = ec[0] + (a0 == K ? 1 : 0);
            const int c1 = ec[1] + (a1 == K ? 1 : 0);
            const int c2 = ec[2] + ((DIM == 3 && a2 == K) ? 1 : 0);
            // K <= 2 here: at most one interior node per direction, so the in-entity index is 0
            // (blk[s] = (K-1)^|s| = 1); entity counts of one component stay below 2^32
            const unsigned ent = (unsigned)c0 + (unsigned)L.csize[s][0] * ((unsigned)c1 + (unsigned)L.csize[s][1] * (unsigned)c2);
            const u64 j = L.compOffset[s] + (u64)ent;
            c[a0 + N1 * (a1 + N1 * a2)] = __ldg(coeff + posA * j + posB);
          }
    }
    double u[NQ];
    sumFactor<DIM, N1, M>(T.A[0], T.A[1], T.A[2], c, u);
    if (values) {
#pragma unroll
      for (int q2 = 0; q2 < (DIM == 3 ? M : 1); ++q2)
#pragma unroll
        for (int q1 = 0; q1 < M; ++q1)
#pragma unroll
          for (int q0 = 0; q0 < M; ++q0) {
            const int q = q0 * R.qs[0] + q1 * R.qs[1] + q2 * R.qs[2];
            vTile[(size_t)tid * rowLen + q * nG + gI] = u[q0 + M * (q1 + M * q2)];
          }
    }
    if (JAC) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        sumFactor<DIM, N1, M>(d == 0 ? T.D[0] : T.A[0], d == 1 ? T.D[1] : T.A[1], d == 2 ? T.D[2] : T.A[2], c, u);
#pragma unroll
        for (int q2 = 0; q2 < (DIM == 3 ? M : 1); ++q2)
#pragma unroll
          for (int q1 = 0; q1 < M; ++q1)
#pragma unroll
            for (int q0 = 0; q0 < M; ++q0) {
              const int q = q0 * R.qs[0] + q1 * R.qs[1] + q2 * R.qs[2];
              jTile[((size_t)tid * rowLen + q * nG + gI) * DIM + d] = u[q0 + M * (q1 + M * q2)] * jinv[d];
            }
      }
    }
  }

  // ---- the block's results leave in output order ----
  const u64 nAct = R.ne - blk0 < (u64)EPB ? R.ne - blk0 : (u64)EPB;
  const bool full = nAct == (u64)EPB;
  if (R.bulk && full) {
    fenceProxyAsync();          // generic-proxy smem writes -> visible to the async proxy
    __syncthreads();
    if (tid == 0) {
      if (R.bulk == 2) {
        if (values) bulkStoreStreaming(values + blk0 * (u64)rowLen, vTile, (unsigned)(EPB * rowLen * sizeof(double)));
        if (JAC) bulkStoreStreaming(jac + blk0 * (u64)rowLen * DIM, jTile, (unsigned)(EPB * rowLen * DIM * sizeof(double)));
      } else {
        if (values) bulkStore(values + blk0 * (u64)rowLen, vTile, (unsigned)(EPB * rowLen * sizeof(double)));
        if (JAC) bulkStore(jac + blk0 * (u64)rowLen * DIM, jTile, (unsigned)(EPB * rowLen * DIM * sizeof(double)));
      }
      bulkStoreWaitRead();      // shared memory must outlive the copy
    }
    return;
  }
  __syncthreads();
  // partial tile, unaligned output or a sub-range of the range entries: cooperative stores
  const int nT = R.ncompTotal, cOff = R.compOffset;
  if (values && !JAC) {
    // values-only instantiations: tile sizes are far below 2^32, 32-bit index arithmetic (the Jacobian
    // instantiations keep the 64-bit loop below: they sit at the register limit and spill otherwise)
    const unsigned cnt = (unsigned)nAct * (unsigned)rowLen;
    double* dst = values + blk0 * (u64)NQ * (u64)nT + cOff;
    for (unsigned i = tid; i < cnt; i += EPB) {
      const unsigned row = i / (unsigned)nG, gI = i - row * (unsigned)nG;       // row = el_local*NQ + q
      dst[(size_t)row * nT + gI] = vTile[i];
    }
  } else if (values) {
    const u64 cnt = nAct * (u64)rowLen;
    for (u64 i = tid; i < cnt; i += EPB) {
      const u64 row = i / (u64)nG, gI = i - row * (u64)nG;       // row = el_local*NQ + q
      values[(blk0 * (u64)NQ + row) * (u64)nT + cOff + gI] = vTile[i];
    }
  }
  if (JAC) {
    const u64 cnt = nAct * (u64)rowLen * DIM;
    for (u64 i = tid; i < cnt; i += EPB) {
      const u64 rg = i / DIM, d = i - rg * DIM;
      const u64 row = rg / (u64)nG, gI = rg - row * (u64)nG;
      jac[((blk0 * (u64)NQ + row) * (u64)nT + cOff + gI) * DIM + d] = jTile[i];
    }
  }
}

// ---------------------------------------------------------------------------
// Several leaves per point (power / composite subtrees): ONE THREAD PER (element, leaf).  A block takes
// epb elements and G = nGroup + nGroup2 leaves, blockDim = epb * G with the leaf fastest, so
//  * the lanes of a warp gather neighbouring elements' coefficients of all leaves at once — for an
//    interleaved power node (position = C j + c) those are consecutive addresses;
//  * every thread contracts one leaf (sum factorisation in registers, as k_eval_reg) and writes its column of
//    the staged tile [epb][NQ][G], which then holds WHOLE points in output order and leaves as one TMA bulk
//    store per block — no 8-byte stores at a G * 8-byte stride;
//  * a block keeps G warps busy per 32 elements of staging tile instead of one.
// K2 >= 0: the last nGroup2 leaves have order K2 on their own lattice L2 (Taylor-Hood: three Q2 velocity
// components + the Q1 pressure, taylorhoodbasis.hh:229-251); the warp then runs both branches.
struct SplitArgs {
  int nGroup, nGroup2, epb;
  int compOffset, ncompTotal;
  int qs[3];
  unsigned long long e0, ne;
  FastDiv divN0, divN1, divG;
  int bulk;
  unsigned long long posA[DFX_MAX_LEAVES], posB[DFX_MAX_LEAVES];    // all G leaves, group 2 behind group 1
  // K2 >= 0, byGroup: the threads [0, epb * nGroup) take group 1 (leaf fastest), the rest group 2 — whole warps on one
  // branch instead of every warp running both (Taylor-Hood: three warps of Q2 velocity, one of Q1 pressure)
  int byGroup;
  FastDiv divGA, divGB;
};
constexpr int kSplitThreads = 128;

template <int DIM, int K, int M, bool JAC, int K2>
__global__ void __launch_bounds__(kSplitThreads, JAC ? 2 : 5)
k_eval_split(const __grid_constant__ DevGrid g, const __grid_constant__ DevLayout L,
             const __grid_constant__ Tables<K + 1, M> T, const __grid_constant__ SplitArgs R,
             const double* __restrict__ coeff, double* __restrict__ values, double* __restrict__ jac,
             const __grid_constant__ DevLayout L2, const __grid_constant__ Tables<(K2 >= 0 ? K2 : 0) + 1, M> T2) {
  constexpr int N1 = K + 1;
  constexpr int NLOC = DIM == 3 ? N1 * N1 * N1 : N1 * N1;
  constexpr int NQ = DIM == 3 ? M * M * M : M * M;
  constexpr int KB = K2 >= 0 ? K2 : 0;
  constexpr int N1B = KB + 1;
  constexpr int NLOCB = DIM == 3 ? N1B * N1B * N1B : N1B * N1B;
  extern __shared__ __align__(128) double stage[];       // values tile [epb][NQ][G], then jac tile [epb][NQ][G][DIM]
  const int G = R.nGroup + R.nGroup2;
  const int epb = R.epb;
  const int tid = threadIdx.x;
  unsigned elL, leaf;
  if (K2 >= 0 && R.byGroup) {
    const unsigned nA = (unsigned)(epb * R.nGroup);
    if ((unsigned)tid < nA) R.divGA.divmod((unsigned)tid, elL, leaf);
    else { R.divGB.divmod((unsigned)tid - nA, elL, leaf); leaf += (unsigned)R.nGroup; }
  } else {
    R.divG.divmod((unsigned)tid, elL, leaf);
  }
  const u64 blk0 = (u64)blockIdx.x * (u64)epb;
  const u64 el = blk0 + elL;
  const bool active = el < R.ne && (int)elL < epb;
  const u64 e = R.e0 + (active ? el : 0);
  int ec[3];
  {
    const unsigned e32 = (unsigned)e;       // numElements < 2^32 checked on the host
    unsigned q, r;
    R.divN0.divmod(e32, q, r);
    ec[0] = (int)r;
    unsigned q2;
    R.divN1.divmod(q, q2, r);
    ec[1] = (int)r; ec[2] = (int)q2;
  }
  const int rowLen = NQ * G;
  double* vTile = stage;
  double* jTile = stage + (size_t)epb * rowLen;
  double jinv[3] = {0, 0, 0};
  if (JAC) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) jinv[d] = elementJacInv(g, d, g.off[d] + ec[d]);
  }
  const u64 posA = R.posA[leaf], posB = R.posB[leaf];
  double* vOut = vTile + (size_t)elL * rowLen + leaf;
  double* jOut = jTile + ((size_t)elL * rowLen + leaf) * DIM;
  if ((int)elL < epb) {
    if (K2 < 0 || (int)leaf < R.nGroup) {
      double c[NLOC];
      gatherLeaf<DIM, K>(L, coeff, posA, posB, e, ec, c);
      double u[NQ];
      if (values) {
        sumFactor<DIM, N1, M>(T.A[0], T.A[1], T.A[2], c, u);
#pragma unroll
        for (int q2 = 0; q2 < (DIM == 3 ? M : 1); ++q2)
#pragma unroll
          for (int q1 = 0; q1 < M; ++q1)
#pragma unroll
            for (int q0 = 0; q0 < M; ++q0)
              vOut[(q0 * R.qs[0] + q1 * R.qs[1] + q2 * R.qs[2]) * G] = u[q0 + M * (q1 + M * q2)];
      }
      if (JAC) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          sumFactor<DIM, N1, M>(d == 0 ? T.D[0] : T.A[0], d == 1 ? T.D[1] : T.A[1], d == 2 ? T.D[2] : T.A[2], c, u);
#pragma unroll
          for (int q2 = 0; q2 < (DIM == 3 ? M : 1); ++q2)
#pragma unroll
            for (int q1 = 0; q1 < M; ++q1)
#pragma unroll
              for (int q0 = 0; q0 < M; ++q0)
                jOut[(q0 * R.qs[0] + q1 * R.qs[1] + q2 * R.qs[2]) * G * DIM + d] = u[q0 + M * (q1 + M * q2)] * jinv[d];
        }
      }
    } else {
      double c[NLOCB];
      gatherLeaf<DIM, KB>(L2, coeff, posA, posB, e, ec, c);
      double u[NQ];
      if (values) {
        sumFactor<DIM, N1B, M>(T2.A[0], T2.A[1], T2.A[2], c, u);
#pragma unroll
        for (int q2 = 0; q2 < (DIM == 3 ? M : 1); ++q2)
#pragma unroll
          for (int q1 = 0; q1 < M; ++q1)
#pragma unroll
            for (int q0 = 0; q0 < M; ++q0)
              vOut[(q0 * R.qs[0] + q1 * R.qs[1] + q2 * R.qs[2]) * G] = u[q0 + M * (q1 + M * q2)];
      }
      if (JAC) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          sumFactor<DIM, N1B, M>(d == 0 ? T2.D[0] : T2.A[0], d == 1 ? T2.D[1] : T2.A[1], d == 2 ? T2.D[2] : T2.A[2], c, u);
#pragma unroll
          for (int q2 = 0; q2 < (DIM == 3 ? M : 1); ++q2)
#pragma unroll
            for (int q1 = 0; q1 < M; ++q1)
#pragma unroll
              for (int q0 = 0; q0 < M; ++q0)
                jOut[(q0 * R.qs[0] + q1 * R.qs[1] + q2 * R.qs[2]) * G * DIM + d] = u[q0 + M * (q1 + M * q2)] * jinv[d];
        }
      }
    }
  }
  // ---- the block's results leave in output order ----
  const u64 nAct = R.ne - blk0 < (u64)epb ? R.ne - blk0 : (u64)epb;
  if (R.bulk && nAct == (u64)epb) {
    fenceProxyAsync();          // generic-proxy smem writes -> visible to the async proxy
    __syncthreads();
    if (tid == 0) {
      if (values) bulkStore(values + blk0 * (u64)rowLen, vTile, (unsigned)(epb * rowLen * sizeof(double)));
      if (JAC) bulkStore(jac + blk0 * (u64)rowLen * DIM, jTile, (unsigned)(epb * rowLen * DIM * sizeof(double)));
      bulkStoreWaitRead();      // shared memory must outlive the copy
    }
    return;
  }
  __syncthreads();
  // partial tile, unaligned output or a sub-range of the range entries: cooperative stores
  const int nT = R.ncompTotal, cOff = R.compOffset;
  if (values) {
    const unsigned cnt = (unsigned)nAct * (unsigned)rowLen;
    double* dst = values + blk0 * (u64)NQ * (u64)nT + cOff;
    for (unsigned i = tid; i < cnt; i += blockDim.x) {
      const unsigned row = i / (unsigned)G, gI = i - row * (unsigned)G;       // row = el_local*NQ + q
      dst[(size_t)row * nT + gI] = vTile[i];
    }
  }
  if (JAC) {
    const unsigned cnt = (unsigned)nAct * (unsigned)rowLen * DIM;
    double* dst = jac + (blk0 * (u64)NQ * (u64)nT + cOff) * DIM;
    for (unsigned i = tid; i < cnt; i += blockDim.x) {
      const unsigned rg = i / DIM, d = i - rg * DIM;
      const unsigned row = rg / (unsigned)G, gI = rg - row * (unsigned)G;
      dst[((size_t)row * nT + gI) * DIM + d] = jTile[i];
    }
  }
}

template <int DIM, int K, int M, bool JAC, int K2>
static int launchSplit(dfx_basis* b, const FastPlan& plan, const FastPlan* plan2, SplitArgs& R, int firstLeaf, int firstLeaf2,
                       const double* coeff, double* values, double* jac, cudaStream_t st) {
  constexpr int NQ = DIM == 3 ? M * M * M : M * M;
  constexpr int KB = K2 >= 0 ? K2 : 0;
  Tables<K + 1, M> T;
  for (int j = 0; j < 3; ++j)
    for (int q = 0; q < M; ++q)
      for (int a = 0; a <= K; ++a) { T.A[j][q][a] = plan.A[j][q][a]; T.D[j][q][a] = plan.Dm[j][q][a]; }
  Tables<KB + 1, M> T2{};
  if (K2 >= 0)
    for (int j = 0; j < 3; ++j)
      for (int q = 0; q < M; ++q)
        for (int a = 0; a <= KB; ++a) { T2.A[j][q][a] = plan2->A[j][q][a]; T2.D[j][q][a] = plan2->Dm[j][q][a]; }
  const int G = R.nGroup + R.nGroup2;
  // elements per block: as many as kSplitThreads threads cover, even (16-byte tiles), tile <= 56 KB
  int epb = kSplitThreads / G;
  const size_t perEl = (size_t)NQ * G * sizeof(double) * (JAC ? 1 + DIM : 1);
  while (epb > 2 && (size_t)epb * perEl > 56 * 1024) epb /= 2;
  epb &= ~1;
  if (epb < 2) return fail(DFX_ERR_NOT_IMPLEMENTED, "evaluation tile exceeds shared memory");
  R.epb = epb; R.divG = FastDiv((unsigned)G);
  R.byGroup = (K2 >= 0 && g_evalSplitByGroup) ? 1 : 0;
  R.divGA = FastDiv((unsigned)std::max(1, R.nGroup)); R.divGB = FastDiv((unsigned)std::max(1, R.nGroup2));
  const size_t smem = (size_t)epb * perEl;
  auto kern = k_eval_split<DIM, K, M, JAC, K2>;
  static std::atomic<unsigned long long> configured{0};
  if (smem > 48 * 1024) { int rc = optInSharedMemory(kern, b->device, configured); if (rc) return rc; }
  const u64 blocks = (R.ne + epb - 1) / epb;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  const DevLayout& L1 = b->dev.layouts[b->dev.leaves[firstLeaf].layout];
  const DevLayout& L2 = K2 >= 0 ? b->dev.layouts[b->dev.leaves[firstLeaf2].layout] : L1;
  const unsigned threads = (unsigned)(((epb * G) + 31) / 32 * 32);
  kern<<<(unsigned)blocks, threads, smem, st>>>(b->dev.grid, L1, T, R, coeff, values, jac, L2, T2);
  return postLaunch("k_eval_split");
}

// [firstA, firstA + nA): leaves of one order on one lattice; [firstB, firstB + nB) (nB may be 0): the run behind it,
// order 1 behind order 2 (the fused Taylor-Hood case).  -1 = not a configuration of this kernel.
int runSplitEval(dfx_basis* b, const FastPlan& planA, int firstA, int nA, const FastPlan* planB, int firstB, int nB,
                 const double* coeff, u64 e0, u64 e1, double* values, double* jac, cudaStream_t st, int compBase, int compTotal) {
  if (e1 == e0) return DFX_OK;
  const DevBasis& D = b->dev;
  if (planA.path != 1 || nA + nB > DFX_MAX_LEAVES || nA + nB < 2) return -1;
  if (nB > 0 && (planB == nullptr || planB->path != 1 || planA.k != 2 || planB->k != 1 || planA.m != planB->m ||
                 planA.firstFastest != planB->firstFastest || planA.jac != planB->jac)) return -1;
  for (int l = 1; l < nA; ++l) if (D.leaves[firstA + l].layout != D.leaves[firstA].layout) return -1;
  for (int l = 1; l < nB; ++l) if (D.leaves[firstB + l].layout != D.leaves[firstB].layout) return -1;
  if ((u64)D.grid.N[0] * D.grid.N[1] * D.grid.N[2] > 0xffffffffull) return -1;
  SplitArgs R;
  R.nGroup = nA; R.nGroup2 = nB;
  for (int q = 0; q < nA; ++q) { R.posA[q] = D.leaves[firstA + q].posA; R.posB[q] = D.leaves[firstA + q].posB; }
  for (int q = 0; q < nB; ++q) { R.posA[nA + q] = D.leaves[firstB + q].posA; R.posB[nA + q] = D.leaves[firstB + q].posB; }
  const int m = planA.m, dim = planA.dim;
  if (planA.firstFastest) { R.qs[0] = 1; R.qs[1] = m; R.qs[2] = m * m; }
  else { int s = 1; R.qs[0] = R.qs[1] = R.qs[2] = 0; for (int j = dim - 1; j >= 0; --j) { R.qs[j] = s; s *= m; } }
  R.e0 = e0; R.ne = e1 - e0;
  R.divN0 = FastDiv((unsigned)D.grid.N[0]); R.divN1 = FastDiv((unsigned)D.grid.N[1]);
  R.compOffset = compBase; R.ncompTotal = compTotal < 0 ? nA + nB : compTotal;
  const bool whole = nA + nB == R.ncompTotal;
  const bool aligned = (!values || ((uintptr_t)values & 15) == 0) && (!jac || ((uintptr_t)jac & 15) == 0);
  R.bulk = whole && aligned ? 1 : 0;
  int rc = -1;
#define DFX_SPLIT(DIM_, K_, M_)                                                                                              \
  if (nB == 0 && dim == DIM_ && planA.k == K_ && m == M_)                                                                    \
    rc = planA.jac ? launchSplit<DIM_, K_, M_, true, -1>(b, planA, nullptr, R, firstA, firstA, coeff, values, jac, st)       \
                   : launchSplit<DIM_, K_, M_, false, -1>(b, planA, nullptr, R, firstA, firstA, coeff, values, jac, st);
  DFX_SPLIT(2, 1, 2) DFX_SPLIT(2, 1, 3) DFX_SPLIT(2, 2, 2) DFX_SPLIT(2, 2, 3)
  DFX_SPLIT(3, 1, 2) DFX_SPLIT(3, 1, 3) DFX_SPLIT(3, 2, 2) DFX_SPLIT(3, 2, 3)
#undef DFX_SPLIT
#define DFX_FUSED(DIM_, M_)                                                                                                  \
  if (nB > 0 && dim == DIM_ && m == M_)                                                                                      \
    rc = planA.jac ? launchSplit<DIM_, 2, M_, true, 1>(b, planA, planB, R, firstA, firstB, coeff, values, jac, st)           \
                   : launchSplit<DIM_, 2, M_, false, 1>(b, planA, planB, R, firstA, firstB, coeff, values, jac, st);
  DFX_FUSED(2, 2) DFX_FUSED(2, 3) DFX_FUSED(3, 2) DFX_FUSED(3, 3)
#undef DFX_FUSED
  return rc;
}

template <int DIM, int K, int M, bool JAC, int EPB>
static int launchRegE(dfx_basis* b, const FastPlan& plan, const RegArgs& R, const double* coeff, double* values,
                      double* jac, cudaStream_t st) {
  constexpr int NQ = DIM == 3 ? M * M * M : M * M;
  Tables<K + 1, M> T;
  for (int j = 0; j < 3; ++j)
    for (int q = 0; q < M; ++q)
      for (int a = 0; a <= K; ++a) { T.A[j][q][a] = plan.A[j][q][a]; T.D[j][q][a] = plan.Dm[j][q][a]; }
  size_t smem = (size_t)EPB * NQ * R.nGroup * sizeof(double) * (JAC ? 1 + DIM : 1);
  auto kern = k_eval_reg<DIM, K, M, JAC, EPB>;
  if (smem > 200 * 1024) return fail(DFX_ERR_NOT_IMPLEMENTED, "evaluation tile exceeds shared memory");
  static std::atomic<unsigned long long> configured{0};
  if (smem > 48 * 1024) { int rc = optInSharedMemory(kern, b->device, configured); if (rc) return rc; }
  const u64 blocks = (R.ne + EPB - 1) / EPB;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  kern<<<(unsigned)blocks, EPB, smem, st>>>(b->dev.grid, b->dev.layouts[b->dev.leaves[R.firstLeaf].layout], T, R, coeff,
                                            values, jac);
  return postLaunch("k_eval_reg");
}

// elements per block: 128 (values) / 64 (with Jacobians); vector-valued groups of a values-only
// evaluation take 32 so that the staging tile (EPB x points x leaves) leaves room for >= 10 blocks per SM
template <int DIM, int K, int M, bool JAC>
static int launchReg(dfx_basis* b, const FastPlan& plan, const RegArgs& R, const double* coeff, double* values,
                     double* jac, cudaStream_t st) {
  if (JAC) return launchRegE<DIM, K, M, JAC, 64>(b, plan, R, coeff, values, jac, st);
  if (R.nGroup >= 2) return launchRegE<DIM, K, M, JAC, 32>(b, plan, R, coeff, values, jac, st);
  if (g_evalBlock == 32) return launchRegE<DIM, K, M, JAC, 32>(b, plan, R, coeff, values, jac, st);
  if (g_evalBlock != 128) return launchRegE<DIM, K, M, JAC, 64>(b, plan, R, coeff, values, jac, st);
  return launchRegE<DIM, K, M, JAC, 128>(b, plan, R, coeff, values, jac, st);
}

int runFastEval(dfx_basis* b, const FastPlan& plan, int firstLeaf, int nLeaves, const double* coeff, u64 e0, u64 e1,
                double* values, double* jac, cudaStream_t st, int compBase, int compTotal) {
  if (e1 == e0) return DFX_OK;
  if (compTotal < 0) compTotal = nLeaves;
  if (plan.path == 2) return runSmemEval(b, plan, firstLeaf, nLeaves, coeff, e0, e1, values, jac, st, compBase, compTotal);
  const DevBasis& D = b->dev;
  if ((u64)D.grid.N[0] * D.grid.N[1] * D.grid.N[2] > 0xffffffffull)
    return fail(DFX_ERR_RANGE, "more than 2^32 elements in one box");
  // runs of consecutive leaves that share a layout
  int l = 0;
  while (l < nLeaves) {
    int r = l + 1;
    while (r < nLeaves && D.leaves[firstLeaf + r].layout == D.leaves[firstLeaf + l].layout) ++r;
    {
      // bound the staging tile (<= 96 KB per block): long runs are cut into chunks of leaves
      int nq = 1; for (int j = 0; j < plan.dim; ++j) nq *= plan.m;
      const size_t perLeaf = (size_t)(plan.jac ? 64 : 32) * nq * sizeof(double) * (plan.jac ? 1 + plan.dim : 1);
      const int maxG = (int)std::max<size_t>(1, (96 * 1024) / perLeaf);
      if (r - l > maxG) r = l + maxG;
    }
    if (r - l >= 2) {
      // several leaves on one lattice: one thread per (element, leaf), whole range entries staged together
      const int rcS = runSplitEval(b, plan, firstLeaf + l, r - l, nullptr, 0, 0, coeff, e0, e1, values, jac, st, compBase + l, compTotal);
      if (rcS >= 0) { if (rcS) return rcS; l = r; continue; }
    }
    RegArgs R;
    R.firstLeaf = firstLeaf + l; R.nGroup = r - l; R.compOffset = compBase + l; R.ncompTotal = compTotal;
    for (int q = 0; q < R.nGroup; ++q) { R.posA[q] = D.leaves[firstLeaf + l + q].posA; R.posB[q] = D.leaves[firstLeaf + l + q].posB; }
    const int m = plan.m, dim = plan.dim;
    if (plan.firstFastest) { R.qs[0] = 1; R.qs[1] = m; R.qs[2] = m * m; }
    else { int s = 1; R.qs[0] = R.qs[1] = R.qs[2] = 0; for (int j = dim - 1; j >= 0; --j) { R.qs[j] = s; s *= m; } }
    R.e0 = e0; R.ne = e1 - e0;
    R.divN0 = FastDiv((unsigned)D.grid.N[0]); R.divN1 = FastDiv((unsigned)D.grid.N[1]);
    const bool whole = R.nGroup == compTotal;
    const bool aligned = (!values || ((uintptr_t)values & 15) == 0) && (!jac || ((uintptr_t)jac & 15) == 0);
    R.bulk = whole && aligned ? (g_evalStreamHint ? 2 : 1) : 0;
    {
      // strip schedule: 3-d, whole z-layers, rows that are whole blocks, and a z-layer too big for the L2 to keep
      const u64 n0 = (u64)D.grid.N[0], n1 = (u64)D.grid.N[1], layer = n0 * n1;
      const unsigned epb = plan.jac ? 64u : (R.nGroup >= 2 ? 32u : (g_evalBlock == 32 ? 32u : (g_evalBlock != 128 ? 64u : 128u)));
      int nqp = 1; for (int j = 0; j < plan.dim; ++j) nqp *= plan.m;
      const u64 layerBytes = layer * (u64)nqp * R.nGroup * sizeof(double) * (plan.jac ? 1 + plan.dim : 1);
      constexpr unsigned kStripH = 32;
      R.swz = 0;
      if (g_evalSwizzle && plan.dim == 3 && n0 % epb == 0 && e0 % layer == 0 && (e1 - e0) % layer == 0 && (e1 - e0) >= 2 * layer &&
          n1 >= 2 * kStripH && (layerBytes > (24ull << 20) || g_evalSwizzle == 2)) {
        const unsigned nz = (unsigned)((e1 - e0) / layer);
        R.swz = 1;
        R.n0 = (unsigned)n0; R.n1 = (unsigned)n1; R.stripH = kStripH;
        R.fullStrips = (unsigned)(n1 / kStripH);
        R.divBpr = FastDiv((unsigned)(n0 / epb));
        R.divStripVol = FastDiv(kStripH * nz);
        R.divStripH = FastDiv(kStripH);
        R.divLastH = FastDiv(std::max(1u, (unsigned)(n1 % kStripH)));
      }
    }
    int rc = DFX_OK;
#define DFX_REG(DIM_, K_, M_)                                                                      \
  if (dim == DIM_ && plan.k == K_ && m == M_)                                                      \
    rc = plan.jac ? launchReg<DIM_, K_, M_, true>(b, plan, R, coeff, values, jac, st)             \
                  : launchReg<DIM_, K_, M_, false>(b, plan, R, coeff, values, jac, st);
    DFX_REG(2, 1, 2) DFX_REG(2, 1, 3) DFX_REG(2, 2, 2) DFX_REG(2, 2, 3)
    DFX_REG(3, 1, 2) DFX_REG(3, 1, 3) DFX_REG(3, 2, 2) DFX_REG(3, 2, 3)
#undef DFX_REG
    if (rc) return rc;
    l = r;
  }
  return DFX_OK;
}

}  // namespace dfx
