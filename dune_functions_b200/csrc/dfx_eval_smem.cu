// dfx_eval_smem.cu — path 2: sum-factorised evaluation for order >= 3 in three dimensions
// (LagrangeDG<4> at 5^3 points is BASELINE config 4; continuous Q3..Q5 use the same kernel).
//
// One element per warp at a time, persistent warps.  The (k+1)^3 coefficients do not fit a
// thread's registers, so the three 1-d contractions run through warp-private shared memory,
// one "pencil" per lane:
//   stage x: c[a0][a1][a2]   -> t1[q0][a1][a2] (+ d1 with the derivative table)
//   stage y: t1[q0][a1][a2]  -> t2[q1][q0][a2] (+ t2y, and t2x from d1)
//   stage z: t2[q1][q0][a2]  -> u[q0][q1][q2], du/dx, du/dy, du/dz
// 9 x (k+1) m^3-ish FMAs per element with gradients instead of 4 (k+1)^3 m^3 for the dense
// contraction (DG4 at 5^3: 5 625 vs 62 500).  Buffers are laid out so that lane p reads
// N1*p + a (odd stride: conflict free) and writes q*P + p (unit stride).  Only __syncwarp
// separates the stages.  The coefficients of the NEXT (element, leaf) item are fetched into
// registers while the current one is contracted.  A warp handles PAIRS of consecutive
// elements so that its result tile is a multiple of 16 bytes and leaves as one TMA bulk store
// per output array; the warp only waits for that store to have read the tile when it is
// about to overwrite it one pair later.  1-d tables sit in the constant bank (kernel
// parameter) and feed DFMA directly.
//
// Replaces LocalFunction::operator() / derivative (discreteglobalbasisfunction.hh:336-371,
// :583-627) for these orders.
#include <algorithm>

#include "dfx_device.cuh"
#include "dfx_eval_fast.h"
#include "dfx_launch.h"

namespace dfx {

template <int N1, int M>
struct Tables3 {
  double A[3][M][N1];
  double D[3][M][N1];
};

struct SmemArgs {
  int firstLeaf, nGroup, compOffset, ncompTotal;
  int qs[3];
  u64 e0, ne;
  FastDiv divN0, divN1, divItems;     // divItems: item -> (pair slot, element in pair, leaf)
  int bulk;
  int perm;                           // 1: contract direction 2 first (points ordered last-coordinate-fastest)
};

__device__ __forceinline__ unsigned smemAddr2(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

constexpr int kSmemWarps = 4;

template <int N1, int M, bool JAC>
struct SmemLayout {
  static constexpr int NLOC = N1 * N1 * N1;
  static constexpr int NQ = M * M * M;
  static constexpr int P1 = N1 * N1, P2 = N1 * M, P3 = M * M;     // pencils per stage
  static constexpr int W1 = (JAC ? 2 : 1) * M * P1;               // t1 (+ d1)
  static constexpr int W2raw = (JAC ? 3 : 1) * M * P2;            // t2 (+ t2x, t2y)
  static constexpr int W2 = W2raw > NLOC ? W2raw : NLOC;          // c aliases the t2 region
  static constexpr int work = (W1 + W2 + 1) & ~1;                 // doubles per warp (even: 16-byte tiles)
  static int staging(int nG) { return 2 * NQ * nG * (JAC ? 4 : 1); }
};

template <int N1, int M, bool JAC>
__global__ void __launch_bounds__(kSmemWarps * 32, 4)
k_eval_smem(const DevBasis* __restrict__ Bp, const __grid_constant__ Tables3<N1, M> T,
            const __grid_constant__ SmemArgs R, const unsigned* __restrict__ localTab,
            const double* __restrict__ coeff, double* __restrict__ values, double* __restrict__ jac) {
  using Lay = SmemLayout<N1, M, JAC>;
  constexpr int NLOC = Lay::NLOC, NQ = Lay::NQ, P1 = Lay::P1, P2 = Lay::P2, P3 = Lay::P3;
  constexpr int SETS = (NLOC + 31) / 32;
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nG = R.nGroup;
  const int stageLen = 2 * NQ * nG;                       // values of the pair
  const int perWarp = Lay::work + stageLen * (JAC ? 4 : 1);
  double* t1 = smem + (size_t)warp * perWarp;
  double* d1 = t1 + M * P1;
  double* t2 = t1 + Lay::W1;
  double* t2x = t2 + M * P2;
  double* t2y = t2x + M * P2;
  double* cbuf = t2;                                       // dead before stage y writes t2
  double* vst = t1 + Lay::work;                            // [2][NQ][nG]
  double* jst = vst + stageLen;                            // [2][NQ][nG][3]
  const DevGrid& g = Bp->grid;

  const u64 nPairs = (R.ne + 1) / 2;
  const u64 warpId = (u64)blockIdx.x * kSmemWarps + warp, nWarps = (u64)gridDim.x * kSmemWarps;
  if (warpId >= nPairs) return;
  // this warp's pairs: warpId, warpId + nWarps, ...; its items: (pair slot, element, leaf)
  const u64 myPairs = (nPairs - warpId + nWarps - 1) / nWarps;
  const unsigned itemsPerPair = 2u * (unsigned)nG;
  const u64 nItems = myPairs * itemsPerPair;

  // per-lane gather constants of the local indices lane, lane+32, ... (same for every leaf of
  // the group: they share the layout): scalar index j = jbase + blk * lex(ec; cs0, cs1)
  const DevLeaf& leaf0 = Bp->leaves[R.firstLeaf];
  const DevLayout& L = Bp->layouts[leaf0.layout];
  u64 jbase[SETS];
  unsigned blk[SETS], cs0[SETS], cs1[SETS], cpos[SETS];
#pragma unroll
  for (int t = 0; t < SETS; ++t) {
    const int i = lane + 32 * t;
    jbase[t] = 0; blk[t] = 0; cs0[t] = 0; cs1[t] = 0; cpos[t] = 0;
    if (i < NLOC) {
      // perm: the coefficient of lattice node (a0, a1, a2) goes to cube position (a2, a1, a0), i.e. the
      // kernel's first contraction runs over z.  The gather itself stays in local-index order (coalesced).
      const int a0 = i % N1, a1 = (i / N1) % N1, a2 = i / (N1 * N1);
      cpos[t] = R.perm ? (unsigned)(a2 + N1 * a1 + N1 * N1 * a0) : (unsigned)i;
      const unsigned packed = __ldg(localTab + leaf0.localOffset + i);
      const int a[3] = {(int)((packed >> 5) & 31u), (int)((packed >> 10) & 31u), (int)((packed >> 15) & 31u)};
      const int zero[3] = {0, 0, 0};
      jbase[t] = layoutIndex(L, g, zero, a, i);
      unsigned s = 0;
      if (L.kind == DFX_NODE_LAGRANGE_DG) s = 7u;
      else
        for (int j = 0; j < 3; ++j) if (a[j] > 0 && a[j] < L.k) s |= 1u << j;
      blk[t] = (unsigned)L.blk[s]; cs0[t] = (unsigned)L.csize[s][0]; cs1[t] = (unsigned)L.csize[s][1];
    }
  }

  // item n of this warp -> element (relative to e0), element-in-pair, leaf; false past the end
  auto locate = [&](u64 n, u64& el, int& eI, int& gI) -> bool {
    if (n >= nItems) return false;
    const u64 slot = n / itemsPerPair;
    const unsigned rem = (unsigned)(n - slot * itemsPerPair);
    eI = (int)(rem / (unsigned)nG); gI = (int)(rem - (unsigned)eI * (unsigned)nG);
    el = (warpId + slot * nWarps) * 2 + (u64)eI;
    return el < R.ne;
  };
  // gather of one item into registers (LocalFunctionBase::bind, :140-147)
  auto fetch = [&](u64 el, int gI, double (&c)[SETS]) {
    const unsigned e = (unsigned)(R.e0 + el);
    unsigned q, ec0, ec1, ec2;
    R.divN0.divmod(e, q, ec0);
    R.divN1.divmod(q, ec2, ec1);
    const DevLeaf& leaf = Bp->leaves[R.firstLeaf + gI];
    const u64 posA = leaf.posA, posB = leaf.posB;
#pragma unroll
    for (int t = 0; t < SETS; ++t) {
      const int i = lane + 32 * t;
      if (i < NLOC) {
        const unsigned ent = ec0 + cs0[t] * (ec1 + cs1[t] * ec2);
        c[t] = __ldg(coeff + posA * (jbase[t] + (u64)blk[t] * (u64)ent) + posB);
      }
    }
  };

  double cur[SETS], nxt[SETS];
  u64 el = 0, elN = 0; int eI = 0, gI = 0, eIN = 0, gIN = 0;
  bool have = locate(0, el, eI, gI);
  if (have) fetch(el, gI, cur);
  bool pendingStore = false;               // lane 0: a bulk store of the staging tile is in flight
  const int nT = R.ncompTotal, cOff = R.compOffset;

  for (u64 n = 0; n < nItems; ++n) {
    const bool haveN = locate(n + 1, elN, eIN, gIN);
    if (haveN) fetch(elN, gIN, nxt);       // in flight while this item is contracted
    if (have) {
      double jinv[3] = {0, 0, 0};
      if (JAC) {
        const unsigned e = (unsigned)(R.e0 + el);
        unsigned q, ec0, ec1, ec2;
        R.divN0.divmod(e, q, ec0);
        R.divN1.divmod(q, ec2, ec1);
        jinv[0] = elementJacInv(g, 0, g.off[0] + (int)ec0);
        jinv[1] = elementJacInv(g, 1, g.off[1] + (int)ec1);
        jinv[2] = elementJacInv(g, 2, g.off[2] + (int)ec2);
      }
      // physical direction of the kernel's first / last contraction and its inverse-Jacobian entry
      const int dFirst = R.perm ? 2 : 0, dLast = R.perm ? 0 : 2;
      const double jFirst = R.perm ? jinv[2] : jinv[0], jLast = R.perm ? jinv[0] : jinv[2];
      __syncwarp();
#pragma unroll
      for (int t = 0; t < SETS; ++t) { const int i = lane + 32 * t; if (i < NLOC) cbuf[cpos[t]] = cur[t]; }
      __syncwarp();
      // ---- stage x: pencil p = a1 + N1 a2 ----
      for (int p = lane; p < P1; p += 32) {
        double in[N1];
#pragma unroll
        for (int a = 0; a < N1; ++a) in[a] = cbuf[N1 * p + a];
#pragma unroll
        for (int q0 = 0; q0 < M; ++q0) {
          double s = T.A[0][q0][0] * in[0];
#pragma unroll
          for (int a = 1; a < N1; ++a) s = fma(T.A[0][q0][a], in[a], s);
          t1[q0 * P1 + p] = s;
          if (JAC) {
            double sd = T.D[0][q0][0] * in[0];
#pragma unroll
            for (int a = 1; a < N1; ++a) sd = fma(T.D[0][q0][a], in[a], sd);
            d1[q0 * P1 + p] = sd;
          }
        }
      }
      __syncwarp();
      // ---- stage y: pencil p = a2 + N1 q0; t1[q0][a1 + N1 a2] = t1[N1 p + a1] ----
      for (int p = lane; p < P2; p += 32) {
        const int base = N1 * p;
        double in[N1], din[N1];
#pragma unroll
        for (int a = 0; a < N1; ++a) { in[a] = t1[base + a]; if (JAC) din[a] = d1[base + a]; }
#pragma unroll
        for (int q1 = 0; q1 < M; ++q1) {
          double s = T.A[1][q1][0] * in[0];
#pragma unroll
          for (int a = 1; a < N1; ++a) s = fma(T.A[1][q1][a], in[a], s);
          t2[q1 * P2 + p] = s;
          if (JAC) {
            double sy = T.D[1][q1][0] * in[0], sx = T.A[1][q1][0] * din[0];
#pragma unroll
            for (int a = 1; a < N1; ++a) { sy = fma(T.D[1][q1][a], in[a], sy); sx = fma(T.A[1][q1][a], din[a], sx); }
            t2y[q1 * P2 + p] = sy;
            t2x[q1 * P2 + p] = sx;
          }
        }
      }
      // first item of a pair: the previous pair's bulk store must have read the staging tile
      if (eI == 0 && gI == 0) {
        if (lane == 0 && pendingStore) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); pendingStore = false; }
      }
      __syncwarp();
      // ---- stage z: pencil p = q0 + M q1; t2[q1][a2 + N1 q0] = t2[N1 p + a2] ----
      for (int p = lane; p < P3; p += 32) {
        const int q1 = p / M, q0 = p - q1 * M;
        const int base = N1 * p;
        double in[N1], inx[N1], iny[N1];
#pragma unroll
        for (int a = 0; a < N1; ++a) { in[a] = t2[base + a]; if (JAC) { inx[a] = t2x[base + a]; iny[a] = t2y[base + a]; } }
#pragma unroll
        for (int q2 = 0; q2 < M; ++q2) {
          double s = T.A[2][q2][0] * in[0];
#pragma unroll
          for (int a = 1; a < N1; ++a) s = fma(T.A[2][q2][a], in[a], s);
          const int qi = q0 * R.qs[0] + q1 * R.qs[1] + q2 * R.qs[2];
          const int o = (eI * NQ + qi) * nG + gI;
          vst[o] = s;
          if (JAC) {
            double gz = T.D[2][q2][0] * in[0], gx = T.A[2][q2][0] * inx[0], gy = T.A[2][q2][0] * iny[0];
#pragma unroll
            for (int a = 1; a < N1; ++a) {
              gz = fma(T.D[2][q2][a], in[a], gz); gx = fma(T.A[2][q2][a], inx[a], gx); gy = fma(T.A[2][q2][a], iny[a], gy);
            }
            // with permuted axes the first contracted direction is z: swap the x and z components back
            jst[o * 3 + dFirst] = gx * jFirst;
            jst[o * 3 + 1] = gy * jinv[1];
            jst[o * 3 + dLast] = gz * jLast;
          }
        }
      }
      // ---- last item of the pair: the tile leaves in output order ----
      const bool lastOfPair = gI == nG - 1 && (eI == 1 || el + 1 >= R.ne);
      if (lastOfPair) {
        const u64 el0 = el - (u64)eI;
        const int nEl = eI + 1;
        if (R.bulk && nEl == 2) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (values)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(values + el0 * (u64)(NQ * nG)),
                           "r"(smemAddr2(vst)), "r"((unsigned)(stageLen * sizeof(double))) : "memory");
            if (JAC)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(jac + el0 * (u64)(NQ * nG) * 3),
                           "r"(smemAddr2(jst)), "r"((unsigned)(stageLen * 3 * sizeof(double))) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            pendingStore = true;
          }
        } else {
          __syncwarp();
          if (values)
            for (int i = lane; i < nEl * NQ * nG; i += 32) {
              const int row = i / nG, c = i - row * nG;                 // row = eI*NQ + q
              values[(el0 * (u64)NQ + (u64)row) * (u64)nT + cOff + c] = vst[i];
            }
          if (JAC)
            for (int i = lane; i < nEl * NQ * nG * 3; i += 32) {
              const int rg = i / 3, d = i - rg * 3;
              const int row = rg / nG, c = rg - row * nG;
              jac[((el0 * (u64)NQ + (u64)row) * (u64)nT + cOff + c) * 3 + d] = jst[i];
            }
          __syncwarp();
        }
      }
    }
    // rotate the prefetched item in
    have = haveN; el = elN; eI = eIN; gI = gIN;
#pragma unroll
    for (int t = 0; t < SETS; ++t) cur[t] = nxt[t];
  }
  // shared memory must outlive the last bulk store's read
  if (lane == 0 && pendingStore) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

bool smemPathHas(int dim, int k, int m) {
  const int n1 = k + 1;
  return dim == 3 && n1 >= 3 && n1 <= 6 && m >= n1 - 1 && m <= n1 + 1 && m >= 2;
}

template <int N1, int M, bool JAC>
static int launchSmem(dfx_basis* b, const FastPlan& plan, const SmemArgs& R, const double* coeff, double* values,
                      double* jac, cudaStream_t st) {
  using Lay = SmemLayout<N1, M, JAC>;
  Tables3<N1, M> T;
  for (int j = 0; j < 3; ++j) {
    const int js = R.perm ? 2 - j : j;           // tables follow the contraction order
    for (int q = 0; q < M; ++q)
      for (int a = 0; a < N1; ++a) { T.A[j][q][a] = plan.A[js][q][a]; T.D[j][q][a] = plan.Dm[js][q][a]; }
  }
  const size_t smem = (size_t)kSmemWarps * (Lay::work + Lay::staging(R.nGroup)) * sizeof(double);
  if (smem > 220 * 1024) return fail(DFX_ERR_NOT_IMPLEMENTED, "evaluation tile exceeds shared memory");
  auto kern = k_eval_smem<N1, M, JAC>;
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  // persistent warps: as many blocks as fit on the device
  int perSM = 0, sms = 0;
  DFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kSmemWarps * 32, smem));
  DFX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
  const u64 pairs = (R.ne + 1) / 2;
  const u64 want = (pairs + kSmemWarps - 1) / kSmemWarps;
  const u64 blocks = std::min<u64>(want, (u64)std::max(1, perSM) * (u64)std::max(1, sms));
  kern<<<(unsigned)blocks, kSmemWarps * 32, smem, st>>>((const DevBasis*)b->dDev, T, R, (const unsigned*)b->dLocalTab,
                                                         coeff, values, jac);
  return postLaunch("k_eval_smem");
}

int runSmemEval(dfx_basis* b, const FastPlan& plan, int firstLeaf, int nLeaves, const double* coeff, u64 e0, u64 e1,
                double* values, double* jac, cudaStream_t st) {
  if (e1 == e0) return DFX_OK;
  const DevBasis& D = b->dev;
  if ((u64)D.grid.N[0] * D.grid.N[1] * D.grid.N[2] > 0xffffffffull)
    return fail(DFX_ERR_RANGE, "more than 2^32 elements in one box");
  const int n1 = plan.k + 1, m = plan.m;
  int l = 0;
  while (l < nLeaves) {
    int r = l + 1;
    while (r < nLeaves && D.leaves[firstLeaf + r].layout == D.leaves[firstLeaf + l].layout) ++r;
    {
      // staging tile <= 24 KB per warp
      const size_t perLeaf = (size_t)2 * m * m * m * sizeof(double) * (plan.jac ? 4 : 1);
      const int maxG = (int)std::max<size_t>(1, (24 * 1024) / perLeaf);
      if (r - l > maxG) r = l + maxG;
    }
    SmemArgs R;
    R.firstLeaf = firstLeaf + l; R.nGroup = r - l; R.compOffset = l; R.ncompTotal = nLeaves;
    // the kernel's first contracted direction is the fastest output index: for points ordered
    // last-coordinate-fastest the axes are permuted (z, y, x) instead of the output strides,
    // so that the staging tile is always written with unit stride across the lanes
    R.qs[0] = 1; R.qs[1] = m; R.qs[2] = m * m;
    R.perm = plan.firstFastest ? 0 : 1;
    R.e0 = e0; R.ne = e1 - e0;
    R.divN0 = FastDiv((unsigned)D.grid.N[0]); R.divN1 = FastDiv((unsigned)D.grid.N[1]);
    R.divItems = FastDiv(2u * (unsigned)R.nGroup);
    const bool whole = R.nGroup == nLeaves;
    const bool aligned = (!values || ((uintptr_t)values & 15) == 0) && (!jac || ((uintptr_t)jac & 15) == 0);
    R.bulk = whole && aligned ? 1 : 0;
    int rc = DFX_ERR_NOT_IMPLEMENTED;
#define DFX_SMEM(N1_, M_)                                                                    \
  if (n1 == N1_ && m == M_)                                                                  \
    rc = plan.jac ? launchSmem<N1_, M_, true>(b, plan, R, coeff, values, jac, st)           \
                  : launchSmem<N1_, M_, false>(b, plan, R, coeff, values, jac, st);
    DFX_SMEM(3, 2) DFX_SMEM(3, 3) DFX_SMEM(3, 4)
    DFX_SMEM(4, 3) DFX_SMEM(4, 4) DFX_SMEM(4, 5)
    DFX_SMEM(5, 4) DFX_SMEM(5, 5) DFX_SMEM(5, 6)
    DFX_SMEM(6, 5) DFX_SMEM(6, 6) DFX_SMEM(6, 7)
#undef DFX_SMEM
    if (rc) return rc;
    l = r;
  }
  return DFX_OK;
}

}  // namespace dfx
