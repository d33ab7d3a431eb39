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
// contraction (DG4 at 5^3: 5 625 vs 62 500).  When there are at least k+1 distinct points per
// direction (GRAD == 2) the gradient is taken from the point values instead: u restricted to
// a grid line is a polynomial of degree k, so its derivative at the m >= k+1 points is the
// collocation derivative C[q][r] = l_r'(x_q) (l_r: Lagrange polynomials through the points)
// applied to the values on that line — 6 contractions (3 750 FMAs) and no derivative
// intermediates in shared memory, which is what bounds this kernel.  Buffers are laid out so that lane p reads
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
  double D[3][M][N1];     // GRAD == 1: l_a'(x_q);  GRAD == 2: unused
  double C[3][M][M];      // GRAD == 2: collocation derivative on the points
};

struct SmemArgs {
  int firstLeaf, nGroup, compOffset, ncompTotal;
  u64 e0, ne;
  FastDiv divN0, divN1, divItems;     // divItems: item -> (pair slot, element in pair, leaf)
  int bulk;
  int isDG;                           // element-local DOFs: scalar index = e * NLOC + i
  int perm;                           // 1: contract direction 2 first (points ordered last-coordinate-fastest)
  int tma;                            // 1 (TMA instantiation): coefficients of element pairs arrive by bulk copies
  int n0, n1;                         // elements per direction (offsets into the diag(J^-1) table)
};

__device__ __forceinline__ unsigned smemAddr2(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

constexpr int kSmemWarps = 4;
int g_smemGrad = -1;        // 1 forces the differentiated-table gradients (testing / comparison)
int g_smemTma = 1;          // DFX_OP_TUNE_SMEM_TMA: 0 coefficients through registers, 1 TMA pair loads for values-only
                            // evaluations of element-local DOFs, 2 also with gradients

// GRAD: 0 values only, 1 gradients by differentiated tables, 2 gradients by collocation
template <int N1, int M, int GRAD>
struct SmemLayout {
  static constexpr bool JAC = GRAD == 1;                          // derivative intermediates t1/t2
  static constexpr int NLOC = N1 * N1 * N1;
  static constexpr int NQ = M * M * M;
  static constexpr int P1 = N1 * N1, P2 = N1 * M, P3 = M * M;     // pencils per stage
  static constexpr int W1 = (JAC ? 2 : 1) * M * P1;               // t1 (+ d1)
  static constexpr int W2raw = (JAC ? 3 : 1) * M * P2;            // t2 (+ t2x, t2y)
  static constexpr int W2 = W2raw > NLOC ? W2raw : NLOC;          // c aliases the t2 region
  static constexpr int work = (W1 + W2 + 1) & ~1;                 // doubles per warp (even: 16-byte tiles)
  static int staging(int nG) { return 2 * NQ * nG * (GRAD ? 4 : 1); }
};

// TMA (element-local DOFs with unit position stride, one leaf, even element range): the 2 (k+1)^3 coefficients of
// an element PAIR are one contiguous, 16-byte aligned run of the vector, so they arrive in shared memory by one
// cp.async.bulk.shared.global (TMA bulk load, mbarrier completion) per pair, two pair buffers per warp: the next
// pair is in flight while the current one is contracted, no LDG -> register -> STS round trip, no register
// rotation, and the permuted point order costs nothing (stage x reads the natural layout with a stride).
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr2(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulkLoad(double* sdst, const double* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr2(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr2(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smemAddr2(bar))
               : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smemAddr2(bar)),
      "r"(parity)
      : "memory");
}

template <int N1, int M, int GRAD, bool TMA>
__global__ void __launch_bounds__(kSmemWarps * 32, GRAD == 1 ? 4 : (GRAD == 2 ? 5 : 6))
k_eval_smem(const DevBasis* __restrict__ Bp, const __grid_constant__ Tables3<N1, M> T,
            const __grid_constant__ SmemArgs R, const unsigned* __restrict__ localTab,
            const double* __restrict__ jinvTab, const double* __restrict__ coeff, double* __restrict__ values, double* __restrict__ jac) {
  using Lay = SmemLayout<N1, M, GRAD>;
  constexpr bool JAC = GRAD == 1;            // differentiated tables through every stage
  constexpr bool OUTJ = GRAD != 0;           // Jacobians are produced
  constexpr int NLOC = Lay::NLOC, NQ = Lay::NQ, P1 = Lay::P1, P2 = Lay::P2, P3 = Lay::P3;
  constexpr int SETS = (NLOC + 31) / 32;
  constexpr int PAIRBUF = 2 * NLOC;          // TMA: doubles per pair buffer (2 (k+1)^3 * 8 bytes is a multiple of 16)
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nG = R.nGroup;
  const int stageLen = 2 * NQ * nG;                       // values of the pair
  const int perWarp = Lay::work + stageLen * (OUTJ ? 4 : 1) + (TMA ? 2 * PAIRBUF + 2 : 0);
  double* t1 = smem + (size_t)warp * perWarp;
  double* d1 = t1 + M * P1;
  double* t2 = t1 + Lay::W1;
  double* t2x = t2 + M * P2;
  double* t2y = t2x + M * P2;
  double* vst = t1 + Lay::work;                            // [2][NQ][nG]
  double* jst = vst + stageLen;                            // [2][NQ][nG][3]
  double* pairBuf = jst + (OUTJ ? 3 * stageLen : 0);       // TMA: [2][PAIRBUF], then the two mbarriers
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(pairBuf + 2 * PAIRBUF);
  const DevGrid& g = Bp->grid;

  const u64 nPairs = (R.ne + 1) / 2;
  const u64 warpId = (u64)blockIdx.x * kSmemWarps + warp, nWarps = (u64)gridDim.x * kSmemWarps;
  if (warpId >= nPairs) return;
  const int nT = R.ncompTotal, cOff = R.compOffset;
  bool pendingStore = false;               // lane 0: a bulk store of the staging tile is in flight

  // ---- the three contractions of one item whose coefficients sit in shared memory at cin[cidx(p) + cstr * a]
  // (p: pencil of stage x, a: index along the contracted direction), and the flush of a finished pair ----
  // diag(J^-1) of an element (AxisAlignedCubeGeometry::jacobianInverse per direction, tabulated once per handle).
  // Loaded ONE ITEM AHEAD like the coefficients: the first use is a select / multiply at the top of the item, and
  // ncu's source view charged 48 % of all stall samples (long scoreboard) to exactly those instructions when the
  // loads were issued there.
  auto loadJinv = [&](const unsigned (&ec)[3], double (&jinv)[3]) {
    if (OUTJ) {
      jinv[0] = __ldg(jinvTab + ec[0]);
      jinv[1] = __ldg(jinvTab + R.n0 + ec[1]);
      jinv[2] = __ldg(jinvTab + R.n0 + R.n1 + ec[2]);
    }
  };
  auto contractItem = [&](const double* cin, bool natural, u64 el, int eI, int gI, const double (&jinv)[3]) {
    // physical direction of the kernel's first / last contraction and its inverse-Jacobian entry
    const int dFirst = R.perm ? 2 : 0, dLast = R.perm ? 0 : 2;
    const double jFirst = R.perm ? jinv[2] : jinv[0], jLast = R.perm ? jinv[0] : jinv[2];
    // ---- stage x: pencil p = a1 + N1 a2 (of the cube in contraction order) ----
    for (int p = lane; p < P1; p += 32) {
      // natural layout + permuted axes: cube position (a, p % N1, p / N1) is lattice node (p / N1, p % N1, a)
      const int cb = (natural && R.perm) ? (p / N1 + N1 * (p % N1)) : N1 * p;
      const int cs = (natural && R.perm) ? N1 * N1 : 1;
      double in[N1];
#pragma unroll
      for (int a = 0; a < N1; ++a) in[a] = cin[cb + cs * a];
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
      const int base = N1 * p;
      double in[N1], inx[N1], iny[N1];
#pragma unroll
      for (int a = 0; a < N1; ++a) { in[a] = t2[base + a]; if (JAC) { inx[a] = t2x[base + a]; iny[a] = t2y[base + a]; } }
      double u[M];
#pragma unroll
      for (int q2 = 0; q2 < M; ++q2) {
        double s = T.A[2][q2][0] * in[0];
#pragma unroll
        for (int a = 1; a < N1; ++a) s = fma(T.A[2][q2][a], in[a], s);
        const int o = (eI * NQ + p + q2 * P3) * nG + gI;      // point q0 + M (q1 + M q2)
        vst[o] = s;
        u[q2] = s;
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
      if (GRAD == 2) {
        // derivative along the last contracted direction: this lane holds the whole line
#pragma unroll
        for (int q2 = 0; q2 < M; ++q2) {
          double gz = T.C[2][q2][0] * u[0];
#pragma unroll
          for (int r = 1; r < M; ++r) gz = fma(T.C[2][q2][r], u[r], gz);
          jst[((eI * NQ + p + q2 * P3) * nG + gI) * 3 + dLast] = gz * jLast;
        }
      }
    }
    if (GRAD == 2) {
      __syncwarp();
      // ---- lines along the first direction: pencil p = q1 + M q2, points r + M p ----
      for (int p = lane; p < P3; p += 32) {
        const int o0 = (eI * NQ + M * p) * nG + gI;
        double u[M];
#pragma unroll
        for (int r = 0; r < M; ++r) u[r] = vst[o0 + r * nG];
#pragma unroll
        for (int q0 = 0; q0 < M; ++q0) {
          double gx = T.C[0][q0][0] * u[0];
#pragma unroll
          for (int r = 1; r < M; ++r) gx = fma(T.C[0][q0][r], u[r], gx);
          jst[(o0 + q0 * nG) * 3 + dFirst] = gx * jFirst;
        }
      }
      // ---- lines along the middle direction: pencil p = q0 + M q2, points q0 + M r + M^2 q2 ----
      for (int p = lane; p < P3; p += 32) {
        const int q2 = p / M, q0 = p - q2 * M;
        const int o0 = (eI * NQ + q0 + P3 * q2) * nG + gI;
        double u[M];
#pragma unroll
        for (int r = 0; r < M; ++r) u[r] = vst[o0 + r * M * nG];
#pragma unroll
        for (int q1 = 0; q1 < M; ++q1) {
          double gy = T.C[1][q1][0] * u[0];
#pragma unroll
          for (int r = 1; r < M; ++r) gy = fma(T.C[1][q1][r], u[r], gy);
          jst[(o0 + q1 * M * nG) * 3 + 1] = gy * jinv[1];
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
          if (OUTJ)
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
        if (OUTJ)
          for (int i = lane; i < nEl * NQ * nG * 3; i += 32) {
            const int rg = i / 3, d = i - rg * 3;
            const int row = rg / nG, c = rg - row * nG;
            jac[((el0 * (u64)NQ + (u64)row) * (u64)nT + cOff + c) * 3 + d] = jst[i];
          }
        __syncwarp();
      }
    }
  };
  auto elementCoords = [&](u64 el, unsigned (&ec)[3]) {
    const unsigned e = (unsigned)(R.e0 + el);
    unsigned q;
    R.divN0.divmod(e, q, ec[0]);
    R.divN1.divmod(q, ec[2], ec[1]);
  };

  if (TMA) {
    // ---- element-local DOFs, one leaf, unit stride: pairs by TMA bulk loads, two buffers per warp ----
    const DevLeaf& leaf = Bp->leaves[R.firstLeaf];
    const double* src = coeff + leaf.posB + R.e0 * (u64)NLOC;
    constexpr unsigned pairBytes = (unsigned)(PAIRBUF * sizeof(double));
    if (lane == 0) {
      mbarInit(&mbar[0], 1);
      mbarInit(&mbar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      bulkLoad(pairBuf, src + warpId * (u64)PAIRBUF, pairBytes, &mbar[0]);
    }
    __syncwarp();
    double jcur[3] = {0, 0, 0}, jnxt[3] = {0, 0, 0};
    {
      unsigned ec0[3];
      elementCoords(warpId * 2, ec0);
      loadJinv(ec0, jcur);
    }
    unsigned it = 0;
    for (u64 pair = warpId; pair < nPairs; pair += nWarps, ++it) {
      const unsigned b = it & 1u;
      const u64 pairN = pair + nWarps;
      // the other buffer was read during the previous iteration: every lane is past those reads
      __syncwarp();
      if (lane == 0 && pairN < nPairs) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulkLoad(pairBuf + (b ^ 1u) * PAIRBUF, src + pairN * (u64)PAIRBUF, pairBytes, &mbar[b ^ 1u]);
      }
      mbarWait(&mbar[b], (it >> 1) & 1u);
      const double* cpair = pairBuf + b * PAIRBUF;
#pragma unroll 1
      for (int eI = 0; eI < 2; ++eI) {
        const u64 el = pair * 2 + (u64)eI;
        // diag(J^-1) of the element after this one (the pair's second, or the next pair's first)
        const u64 elN = eI == 0 ? el + 1 : pairN * 2;
        if (OUTJ && elN < R.ne) {
          unsigned ecN[3];
          elementCoords(elN, ecN);
          loadJinv(ecN, jnxt);
        }
        __syncwarp();
        contractItem(cpair + eI * NLOC, true, el, eI, 0, jcur);
#pragma unroll
        for (int j = 0; j < 3; ++j) jcur[j] = jnxt[j];
      }
    }
    if (lane == 0 && pendingStore) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    return;
  }

  // ---- general path: per-lane gather into registers, prefetched one item ahead ----
  double* cbuf = t2;                                       // dead before stage y writes t2
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

  // next item after (pr, e_, g_): leaves of the group, then the second element, then the next pair
  auto advance = [&](u64& pr, int& e_, int& g_) {
    if (++g_ < nG) return;
    g_ = 0;
    if (e_ == 0 && pr * 2 + 1 < R.ne) { e_ = 1; return; }
    e_ = 0; pr += nWarps;
  };
  // gather of one item into registers (LocalFunctionBase::bind, :140-147)
  auto fetch = [&](u64 el, int gI, double (&c)[SETS], unsigned (&ec)[3], double (&jinv)[3]) {
    elementCoords(el, ec);
    loadJinv(ec, jinv);
    const unsigned e = (unsigned)(R.e0 + el);
    const DevLeaf& leaf = Bp->leaves[R.firstLeaf + gI];
    const u64 posA = leaf.posA, posB = leaf.posB;
    if (R.isDG) {
      const double* srcp = coeff + posA * ((u64)e * (u64)NLOC) + posB;
#pragma unroll
      for (int t = 0; t < SETS; ++t) {
        const int i = lane + 32 * t;
        if (i < NLOC) c[t] = __ldg(srcp + posA * (u64)i);
      }
      return;
    }
#pragma unroll
    for (int t = 0; t < SETS; ++t) {
      const int i = lane + 32 * t;
      if (i < NLOC) {
        const unsigned ent = ec[0] + cs0[t] * (ec[1] + cs1[t] * ec[2]);
        c[t] = __ldg(coeff + posA * (jbase[t] + (u64)blk[t] * (u64)ent) + posB);
      }
    }
  };

  double cur[SETS], nxt[SETS];
  double jcur[3] = {0, 0, 0}, jnxt[3] = {0, 0, 0};
  unsigned ec[3], ecN[3];
  u64 pair = warpId, pairN;
  int eI = 0, gI = 0, eIN, gIN;
  fetch(pair * 2, 0, cur, ec, jcur);

  while (true) {
    const u64 el = pair * 2 + (u64)eI;
    pairN = pair; eIN = eI; gIN = gI;
    advance(pairN, eIN, gIN);
    const bool haveN = pairN < nPairs;
    if (haveN) fetch(pairN * 2 + (u64)eIN, gIN, nxt, ecN, jnxt);       // in flight while this item is contracted
    __syncwarp();
#pragma unroll
    for (int t = 0; t < SETS; ++t) { const int i = lane + 32 * t; if (i < NLOC) cbuf[cpos[t]] = cur[t]; }
    __syncwarp();
    contractItem(cbuf, false, el, eI, gI, jcur);
    if (!haveN) break;
    // rotate the prefetched item in
    pair = pairN; eI = eIN; gI = gIN;
#pragma unroll
    for (int t = 0; t < SETS; ++t) cur[t] = nxt[t];
#pragma unroll
    for (int j = 0; j < 3; ++j) { ec[j] = ecN[j]; jcur[j] = jnxt[j]; }
  }
  // shared memory must outlive the last bulk store's read
  if (lane == 0 && pendingStore) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// diag(J^-1) per direction and element coordinate, [N0 | N1 | N2] (elementJacInv of the local box)
__global__ void k_jinv_table(const __grid_constant__ DevGrid g, double* __restrict__ tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n01 = g.N[0] + g.N[1];
  if (t >= n01 + g.N[2]) return;
  const int j = t >= n01 ? 2 : (t >= g.N[0] ? 1 : 0);
  const int i = t - (j == 2 ? n01 : (j == 1 ? g.N[0] : 0));
  tab[t] = j < g.dim ? elementJacInv(g, j, g.off[j] + i) : 1.0;
}

int ensureJinvTable(dfx_basis* b, cudaStream_t st) {
  if (b->dJinv) return DFX_OK;
  const DevGrid& g = b->dev.grid;
  const int n = g.N[0] + g.N[1] + g.N[2];
  DFX_CUDA(cudaMalloc(&b->dJinv, (size_t)n * sizeof(double)));
  k_jinv_table<<<(n + 255) / 256, 256, 0, st>>>(g, (double*)b->dJinv);
  return postLaunch("k_jinv_table");
}

bool smemPathHas(int dim, int k, int m) {
  const int n1 = k + 1;
  return dim == 3 && n1 >= 3 && n1 <= 6 && m >= n1 - 1 && m <= n1 + 1 && m >= 2;
}

template <int N1, int M, int GRAD, bool TMA>
static int launchSmemT(dfx_basis* b, const FastPlan& plan, const SmemArgs& R, const double* coeff, double* values,
                       double* jac, cudaStream_t st) {
  using Lay = SmemLayout<N1, M, GRAD>;
  Tables3<N1, M> T;
  for (int j = 0; j < 3; ++j) {
    const int js = R.perm ? 2 - j : j;           // tables follow the contraction order
    for (int q = 0; q < M; ++q) {
      for (int a = 0; a < N1; ++a) { T.A[j][q][a] = plan.A[js][q][a]; T.D[j][q][a] = plan.Dm[js][q][a]; }
      for (int r = 0; r < M; ++r) T.C[j][q][r] = plan.Cm[js][q][r];
    }
  }
  const size_t smem = (size_t)kSmemWarps * (Lay::work + Lay::staging(R.nGroup) + (TMA ? 4 * Lay::NLOC + 2 : 0)) * sizeof(double);
  if (smem > 220 * 1024) return fail(DFX_ERR_NOT_IMPLEMENTED, "evaluation tile exceeds shared memory");
  auto kern = k_eval_smem<N1, M, GRAD, TMA>;
  static std::atomic<unsigned long long> configured{0};
  if (smem > 48 * 1024) { int rc = optInSharedMemory(kern, b->device, configured); if (rc) return rc; }
  // persistent warps: as many blocks as fit on the device
  int perSM = 0, sms = 0;
  DFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kSmemWarps * 32, smem));
  DFX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
  const u64 pairs = (R.ne + 1) / 2;
  const u64 want = (pairs + kSmemWarps - 1) / kSmemWarps;
  u64 blocks = std::min<u64>(want, (u64)std::max(1, perSM) * (u64)std::max(1, sms));
  if (g_smemMaxBlocks > 0) blocks = std::min<u64>(blocks, (u64)g_smemMaxBlocks);
  kern<<<(unsigned)blocks, kSmemWarps * 32, smem, st>>>((const DevBasis*)b->dDev, T, R, (const unsigned*)b->dLocalTab,
                                                         (const double*)b->dJinv, coeff, values, jac);
  return postLaunch("k_eval_smem");
}

template <int N1, int M, int GRAD>
static int launchSmem(dfx_basis* b, const FastPlan& plan, const SmemArgs& R, const double* coeff, double* values,
                      double* jac, cudaStream_t st) {
  // pairs of elements are 16-byte aligned runs of the vector: (k+1)^3 doubles per element, two of them per pair
  if (R.tma && (2 * N1 * N1 * N1 * sizeof(double)) % 16 == 0) return launchSmemT<N1, M, GRAD, true>(b, plan, R, coeff, values, jac, st);
  return launchSmemT<N1, M, GRAD, false>(b, plan, R, coeff, values, jac, st);
}

int runSmemEval(dfx_basis* b, const FastPlan& plan, int firstLeaf, int nLeaves, const double* coeff, u64 e0, u64 e1,
                double* values, double* jac, cudaStream_t st, int compBase, int compTotal) {
  if (e1 == e0) return DFX_OK;
  if (compTotal < 0) compTotal = nLeaves;
  const DevBasis& D = b->dev;
  if ((u64)D.grid.N[0] * D.grid.N[1] * D.grid.N[2] > 0xffffffffull)
    return fail(DFX_ERR_RANGE, "more than 2^32 elements in one box");
  const int n1 = plan.k + 1, m = plan.m;
  if (nLeaves == 1 && compTotal == 1 && compBase == 0) {
    // element-local DOFs, one leaf, whole pairs: both elements of a pair at once (dfx_eval_pair.cu)
    int rc = runPairEval(b, plan, firstLeaf, coeff, e0, e1, values, jac, st);
    if (rc >= 0) return rc;
    // continuous order 3 at 4^3 points, values only: thread per element in registers (dfx_eval_q3.cu)
    if (!jac) { rc = runQ3Eval(b, plan, firstLeaf, coeff, e0, e1, values, st); if (rc >= 0) return rc; }
  }
  if (plan.jac) { int rc = ensureJinvTable(b, st); if (rc) return rc; }
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
    R.firstLeaf = firstLeaf + l; R.nGroup = r - l; R.compOffset = compBase + l; R.ncompTotal = compTotal;
    // the kernel's first contracted direction is the fastest output index: for points ordered
    // last-coordinate-fastest the axes are permuted (z, y, x) instead of the output strides,
    // so that the staging tile is always written with unit stride across the lanes
    R.perm = plan.firstFastest ? 0 : 1;
    R.isDG = D.layouts[D.leaves[firstLeaf + l].layout].kind == DFX_NODE_LAGRANGE_DG ? 1 : 0;
    R.e0 = e0; R.ne = e1 - e0;
    R.n0 = D.grid.N[0]; R.n1 = D.grid.N[1];
    {
      // TMA pair loads: element-local DOFs, one leaf with unit position stride, whole pairs, every pair a 16-byte
      // aligned run of the vector
      const DevLeaf& lf = D.leaves[firstLeaf + l];
      const size_t pairBytes = (size_t)2 * n1 * n1 * n1 * sizeof(double);
      const uintptr_t first = (uintptr_t)(coeff + lf.posB + e0 * (u64)(n1 * n1 * n1));
      // values only: with gradients the two pair buffers cost a block per SM (14 KB of shared memory per warp) and the
      // kernel gets slower (DG4 192^3: 6.65 -> 6.97 ms), while values alone gain (2.92 -> 2.60 ms); g_smemTma == 2 forces it
      const bool wanted = g_smemTma == 2 || (g_smemTma == 1 && !plan.jac);
      R.tma = (wanted && R.isDG && R.nGroup == 1 && lf.posA == 1 && (e1 - e0) % 2 == 0 && pairBytes % 16 == 0 && first % 16 == 0) ? 1 : 0;
    }
    R.divN0 = FastDiv((unsigned)D.grid.N[0]); R.divN1 = FastDiv((unsigned)D.grid.N[1]);
    R.divItems = FastDiv(2u * (unsigned)R.nGroup);
    const bool whole = R.nGroup == compTotal;
    const bool aligned = (!values || ((uintptr_t)values & 15) == 0) && (!jac || ((uintptr_t)jac & 15) == 0);
    R.bulk = whole && aligned ? 1 : 0;
    int rc = DFX_ERR_NOT_IMPLEMENTED;
#define DFX_SMEM(N1_, M_)                                                                    \
  if (n1 == N1_ && m == M_)                                                                  \
    rc = !plan.jac ? launchSmem<N1_, M_, 0>(b, plan, R, coeff, values, jac, st)             \
         : (M_ >= N1_ && plan.colloc && g_smemGrad != 1)                                     \
             ? launchSmem<N1_, M_, (M_ >= N1_ ? 2 : 1)>(b, plan, R, coeff, values, jac, st)  \
             : launchSmem<N1_, M_, 1>(b, plan, R, coeff, values, jac, st);
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
