// dfx_eval_pair.cu — sum-factorised evaluation of element-local DOFs (LagrangeDG, order >= 2 in 3-d; LagrangeDG<4> at
// 5^3 points with values and gradients is BASELINE config 4): two elements per warp at once, results leaving in
// BLOCK tiles.
//
// Why this kernel exists (round 2, `profiles/r02_dgstream.jsonl`): k_eval_smem (one element per warp at a time,
// 599 warp instructions per DG4 element with gradients) and a first version of this kernel (425 instructions, twice
// the independent work per stage) both ran DG4 192^3 values + gradients in 6.7-6.9 ms — and so did a probe that
// moves the same bytes with NO arithmetic (tools/probes/dgstream.cu: 6.6-6.7 ms).  The limit was the store pattern:
// one 2 000-byte and one 6 000-byte bulk store per warp, at addresses that are not multiples of 128, reach
// 5.3 TB/s; the same bytes leaving as one 16 KB and one 48 KB store per block of 8 warps, every chunk a whole
// number of 128-byte lines, reach 6.2 TB/s (5.68 ms).  So the warps of a block put their pairs' results side by
// side into one block tile and ONE thread stores it.
//
// A lane owns the SAME pencil of both elements of its warp's pair, so every table entry, address and loop test serves
// two pencils and each stage carries two independent sets of dependency chains.  Everything that is a run-time
// quantity in the general kernel is a template parameter (one leaf, unit position stride, point order, gradients by
// collocation).
//
// Data flow per block iteration (WARPS consecutive pairs):
//   per block: ONE TMA bulk load (mbarrier) of the 8 pairs' 16 (k+1)^3 coefficients — one 16-byte aligned run of
//     the vector, whole 128-byte lines for every k — into the landing area; the NEXT group's load is issued right
//     behind the first barrier, when every warp is past stage x (one load per warp, issued as soon as that warp had read its buffer, arrived late under the write
//     stream and left the warps of a block out of step: 6.1-6.3 ms against 5.8-5.9 ms)
//   stage x: pb -> t1[e][q0][a2][a1]      stage y: t1 -> t2[e][q1][q0][a2]            (warp-private, __syncwarp)
//   barrier: the previous iteration's VALUES store (the older, smaller bulk group) has read its tile
//   stage z: t2 -> u = vst[warp][e][point]; the derivative of u along the last direction stays in registers
//   barrier: the previous iteration's JACOBIAN store has read its tile (thread 0 waits, __syncthreads)
//   gradients = collocation derivative of u along the three grid lines (jst[warp][e][point][3], scaled by diag(J^-1))
//   barrier, then thread 0 issues two bulk stores in two groups: WARPS * 2 m^3 values, then three times as many
//     Jacobian entries.
// Two blocks per SM run out of phase: one contracts stages x / y while the other's tile drains.
//
// Replaces LocalFunction::operator() / derivative (discreteglobalbasisfunction.hh:336-371, :583-627) for
// LagrangeDGPreBasis leaves (lagrangedgbasis.hh:32-170: element-local numbering e * (k+1)^3 + i).
#include <algorithm>

#include "dfx_device.cuh"
#include "dfx_eval_fast.h"
#include "dfx_launch.h"

namespace dfx {

int g_smemPair = -1;  // DFX_OP_TUNE_SMEM_PAIR: -1 default (with gradients only), 0 off, 1 always (8 warps per block when they fit), 2 the
                      // same with one bulk load per warp instead of one per block, 4 always with four warps
int g_pairL2Ahead = 0;     // DFX_OP_TUNE_PAIR_PREFETCH: iterations ahead of an L2 prefetch of coefficients (0 / 1: none).  Helped the
                           // per-warp loads (2 ahead: 1-3 %), costs 3 % with one load per block (profiles/r02_c4eval_l2_prefetch_ab.jsonl)
int g_smemMaxBlocks = 0;   // DFX_OP_TUNE_SMEM_BLOCKS: > 0 caps the grid of the persistent-warp kernels (tests: many items per warp)

template <int N1, int M>
struct PairTables {
  double A[3][M][N1];     // l_a(x_q) per contraction stage
  double C[3][M][M];      // collocation derivative on the points
};

struct PairArgs {
  int l2ahead;            // > 0: the coefficients of the pair that many iterations ahead are prefetched into the L2
  u64 nPairs;
  unsigned e0;            // first element of the range (box-linear)
  unsigned n0, n1;        // elements per direction (offsets into the diag(J^-1) table)
  FastDiv divN0, divN1;
};

template <int N1, int M, bool GRAD>
struct PairLayout {
  static constexpr int S = N1 | 1;                       // pencil stride of t1 / t2 (odd: conflict-free reads)
  static constexpr int NLOC = N1 * N1 * N1, NQ = M * M * M;
  static constexpr int P1 = N1 * N1, P2 = N1 * M, P3 = M * M;
  static constexpr int T1E = M * N1 * S, T2E = M * M * S; // per element
  // warp-private work area (doubles; every offset even: 16-byte aligned TMA destinations)
  static constexpr int oPB = 0;
  static constexpr int oT1 = oPB + 2 * NLOC;
  static constexpr int oT2 = oT1 + 2 * T1E;
  static constexpr int oBAR = oT2 + 2 * T2E;
  static constexpr int perWarp = oBAR + 2;
  // block tiles behind the work areas: values [WARPS][2][NQ], Jacobians [WARPS][2][NQ][3]
  static constexpr int tileV = 2 * NQ, tileJ = GRAD ? 6 * NQ : 0;
  static constexpr size_t bytes(int warps) { return (size_t)warps * (perWarp + tileV + tileJ) * sizeof(double); }
};

namespace {
__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pairBarInit(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(bar)) : "memory");
}
__device__ __forceinline__ void pairLoad(double* sdst, const double* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(sdst)),
               "l"(gsrc), "r"(bytes), "r"(sa(bar))
               : "memory");
}
__device__ __forceinline__ void pairWait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PD_%=;\n"
      "bra PW_%=;\n"
      "PD_%=:\n"
      "}\n" ::"r"(sa(bar)),
      "r"(parity)
      : "memory");
}
}  // namespace

// PERM: points ordered last-coordinate-fastest — the kernel's first contraction runs over z (tables arrive in
// contraction order), stage x reads the natural layout with a stride, and the x / z gradient entries swap back.
// BLK: the WARPS pairs of a block iteration are one run of the vector too — ONE bulk load per block (thread 0, one
// mbarrier) into a contiguous landing area instead of one 2 (k+1)^3-entry load per warp.
template <int N1, int M, bool GRAD, bool PERM, int WARPS, bool BLK>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 8 ? 2 : 4)
k_eval_pair(const __grid_constant__ PairTables<N1, M> T, const __grid_constant__ PairArgs R,
            const double* __restrict__ jinvTab, const double* __restrict__ src, double* __restrict__ values,
            double* __restrict__ jac) {
  using Lay = PairLayout<N1, M, GRAD>;
  constexpr int S = Lay::S, NLOC = Lay::NLOC, NQ = Lay::NQ, P1 = Lay::P1, P2 = Lay::P2, P3 = Lay::P3;
  constexpr int T1E = Lay::T1E, T2E = Lay::T2E;
  constexpr int dFirst = PERM ? 2 : 0, dLast = PERM ? 0 : 2;
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // BLK: [WARPS][2 NLOC] landing area first, then the warps' t1 / t2 / barrier slots (same total size)
  double* base = BLK ? smem + (size_t)WARPS * (2 * Lay::NLOC) + (size_t)warp * (Lay::perWarp - 2 * Lay::NLOC) - 2 * Lay::NLOC
                     : smem + (size_t)warp * Lay::perWarp;
  double* pb = BLK ? smem + (size_t)warp * (2 * Lay::NLOC) : base + Lay::oPB;
  double* t1 = base + Lay::oT1;
  double* t2 = base + Lay::oT2;
  unsigned long long* bar = BLK ? reinterpret_cast<unsigned long long*>(smem + (size_t)WARPS * (2 * Lay::NLOC) + (Lay::oBAR - 2 * Lay::NLOC))
                                : reinterpret_cast<unsigned long long*>(base + Lay::oBAR);      // BLK: warp 0's slot, for the block
  double* tileV = smem + (size_t)WARPS * Lay::perWarp;
  double* tileJ = tileV + WARPS * Lay::tileV;
  double* vst = tileV + warp * Lay::tileV;
  double* jst = tileJ + warp * Lay::tileJ;

  const u64 nGroups = (R.nPairs + WARPS - 1) / WARPS;
  const u64 stride = (u64)gridDim.x * WARPS;
  constexpr unsigned pairBytes = (unsigned)(2 * NLOC * sizeof(double));
  u64 pair = (u64)blockIdx.x * WARPS + warp;

  // BLK: coefficients of the group's pairs (the last group may be short), by thread 0
  auto groupLoad = [&](u64 grp) {
    const u64 first = grp * WARPS;
    const unsigned n = (unsigned)min((u64)WARPS, R.nPairs - first);
    pairLoad(smem, src + first * (u64)(2 * NLOC), n * pairBytes, bar);
  };
  if (BLK) {
    if (threadIdx.x == 0) {
      pairBarInit(bar);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if ((u64)blockIdx.x < nGroups) groupLoad(blockIdx.x);
    }
    __syncthreads();
  } else {
    if (lane == 0) {
      pairBarInit(bar);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if (pair < R.nPairs) pairLoad(pb, src + pair * (u64)(2 * NLOC), pairBytes, bar);
    }
    __syncwarp();
  }

  // diag(J^-1) of both elements of a pair (AxisAlignedCubeGeometry::jacobianInverse, tabulated per handle), fetched
  // one pair ahead
  auto loadJinv = [&](u64 pr, double (&j)[2][3]) {
    if (!GRAD) return;
    const unsigned e = R.e0 + (unsigned)(pr * 2);
    unsigned q, c0, c1, c2;
    R.divN0.divmod(e, q, c0);
    R.divN1.divmod(q, c2, c1);
    j[0][0] = __ldg(jinvTab + c0);
    j[0][1] = __ldg(jinvTab + R.n0 + c1);
    j[0][2] = __ldg(jinvTab + R.n0 + R.n1 + c2);
    if (++c0 == R.n0) { c0 = 0; if (++c1 == R.n1) { c1 = 0; ++c2; } }
    j[1][0] = __ldg(jinvTab + c0);
    j[1][1] = __ldg(jinvTab + R.n0 + c1);
    j[1][2] = __ldg(jinvTab + R.n0 + R.n1 + c2);
  };
  double jc[2][3] = {{0, 0, 0}, {0, 0, 0}}, jn[2][3] = {{0, 0, 0}, {0, 0, 0}};
  if (pair < R.nPairs) loadJinv(pair, jc);

  bool pending = false;        // thread 0: the previous iteration's bulk stores may still be reading the block tiles
  unsigned it = 0;
  for (u64 grp = blockIdx.x; grp < nGroups; grp += gridDim.x, pair += stride, ++it) {
    const bool have = pair < R.nPairs;          // the last group may be short
    const u64 pairN = pair + stride;
    if (have) {
      if (pairN < R.nPairs) loadJinv(pairN, jn);
      pairWait(bar, it & 1u);
      __syncwarp();
      // ---- stage x: pencil p = a1 + N1 a2 of both elements ----
#pragma unroll
      for (int p0 = 0; p0 < P1; p0 += 32) {
        const int p = p0 + lane;
        if (P1 - p0 >= 32 || p < P1) {
          const int cb = PERM ? (p / N1 + N1 * (p % N1)) : N1 * p;
          constexpr int cs = PERM ? N1 * N1 : 1;
          double in0[N1], in1[N1];
#pragma unroll
          for (int a = 0; a < N1; ++a) { in0[a] = pb[cb + cs * a]; in1[a] = pb[NLOC + cb + cs * a]; }
          const int o = S == N1 ? p : (p / N1) * S + p % N1;
#pragma unroll
          for (int q = 0; q < M; ++q) {
            double s0 = T.A[0][q][0] * in0[0], s1 = T.A[0][q][0] * in1[0];
#pragma unroll
            for (int a = 1; a < N1; ++a) { s0 = fma(T.A[0][q][a], in0[a], s0); s1 = fma(T.A[0][q][a], in1[a], s1); }
            t1[q * (N1 * S) + o] = s0;
            t1[T1E + q * (N1 * S) + o] = s1;
          }
        }
      }
      __syncwarp();
      // pb has been read: the warp's next pair can land while this one goes through stages y, z and the lines
      if (!BLK && lane == 0 && pairN < R.nPairs) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        pairLoad(pb, src + pairN * (u64)(2 * NLOC), pairBytes, bar);
        // ... and the pair after that one starts its way from DRAM to the L2 now: under the write stream of this kernel
        // a bulk load issued one iteration ahead still arrived late (ncu: 12 % of the stall samples on the mbarrier wait)
        const u64 pairP = pair + (u64)R.l2ahead * stride;
        if (R.l2ahead > 1 && pairP < R.nPairs)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + pairP * (u64)(2 * NLOC)), "r"(pairBytes) : "memory");
      }
      // ---- stage y: pencil p = a2 + N1 q0; t1[(q0 N1 + a2) S + a1] = t1[S p + a1] ----
#pragma unroll
      for (int p0 = 0; p0 < P2; p0 += 32) {
        const int p = p0 + lane;
        if (P2 - p0 >= 32 || p < P2) {
          double in0[N1], in1[N1];
#pragma unroll
          for (int a = 0; a < N1; ++a) { in0[a] = t1[S * p + a]; in1[a] = t1[T1E + S * p + a]; }
          const int o = S == N1 ? p : (p / N1) * S + p % N1;          // (q0, a2) -> q0 S + a2
#pragma unroll
          for (int q = 0; q < M; ++q) {
            double s0 = T.A[1][q][0] * in0[0], s1 = T.A[1][q][0] * in1[0];
#pragma unroll
            for (int a = 1; a < N1; ++a) { s0 = fma(T.A[1][q][a], in0[a], s0); s1 = fma(T.A[1][q][a], in1[a], s1); }
            t2[q * (M * S) + o] = s0;
            t2[T2E + q * (M * S) + o] = s1;
          }
        }
      }
    }
    // the previous iteration's stores must have read the block tiles before they are overwritten.  EARLY (one pass of
    // stage-z pencils): the values tile — the older, four times smaller bulk group — is awaited here, the Jacobian
    // tile only behind stage z, whose derivative along the last direction waits in registers meanwhile; the 48 KB of
    // Jacobians then drain under stages x, y AND z of the next iteration instead of x and y only
    constexpr bool EARLY = GRAD && P3 <= 32;
    if (threadIdx.x == 0 && pending) {
      if (EARLY) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); pending = false; }
    }
    __syncthreads();
    if (BLK && threadIdx.x == 0) {
      // every warp is past stage x (it has passed the barrier): the landing area can take the next group
      const u64 grpN = grp + gridDim.x;
      if (grpN < nGroups) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        groupLoad(grpN);
        const u64 grpP = grp + (u64)R.l2ahead * gridDim.x;
        if (R.l2ahead > 1 && grpP < nGroups) {
          const u64 first = grpP * WARPS;
          const unsigned n = (unsigned)min((u64)WARPS, R.nPairs - first);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + first * (u64)(2 * NLOC)), "r"(n * pairBytes) : "memory");
        }
      }
    }
    double gz0[EARLY ? M : 1], gz1[EARLY ? M : 1];
    if (have) {
      // ---- stage z: pencil p = q0 + M q1; t2[(q1 M + q0) S + a2] = t2[S p + a2]; point p + P3 q2 ----
#pragma unroll
      for (int p0 = 0; p0 < P3; p0 += 32) {
        const int p = p0 + lane;
        if (P3 - p0 >= 32 || p < P3) {
          double in0[N1], in1[N1];
#pragma unroll
          for (int a = 0; a < N1; ++a) { in0[a] = t2[S * p + a]; in1[a] = t2[T2E + S * p + a]; }
          double u0[M], u1[M];
#pragma unroll
          for (int q = 0; q < M; ++q) {
            double s0 = T.A[2][q][0] * in0[0], s1 = T.A[2][q][0] * in1[0];
#pragma unroll
            for (int a = 1; a < N1; ++a) { s0 = fma(T.A[2][q][a], in0[a], s0); s1 = fma(T.A[2][q][a], in1[a], s1); }
            vst[p + q * P3] = s0;
            vst[NQ + p + q * P3] = s1;
            u0[q] = s0; u1[q] = s1;
          }
          if (GRAD) {
            // derivative along the last contracted direction: this lane holds both whole lines
#pragma unroll
            for (int q = 0; q < M; ++q) {
              double g0 = T.C[2][q][0] * u0[0], g1 = T.C[2][q][0] * u1[0];
#pragma unroll
              for (int r = 1; r < M; ++r) { g0 = fma(T.C[2][q][r], u0[r], g0); g1 = fma(T.C[2][q][r], u1[r], g1); }
              if (EARLY) { gz0[q] = g0 * jc[0][dLast]; gz1[q] = g1 * jc[1][dLast]; }
              else {
                jst[(p + q * P3) * 3 + dLast] = g0 * jc[0][dLast];
                jst[(NQ + p + q * P3) * 3 + dLast] = g1 * jc[1][dLast];
              }
            }
          }
        }
      }
    }
    if (EARLY) {
      if (threadIdx.x == 0 && pending) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); pending = false; }
      __syncthreads();
    }
    if (have) {
      if (EARLY) {
        if (lane < P3) {
#pragma unroll
          for (int q = 0; q < M; ++q) {
            jst[(lane + q * P3) * 3 + dLast] = gz0[q];
            jst[(NQ + lane + q * P3) * 3 + dLast] = gz1[q];
          }
        }
      }
      if (GRAD) {
        __syncwarp();
        // ---- lines along the first contracted direction: pencil p = q1 + M q2, points r + M p ----
#pragma unroll
        for (int p0 = 0; p0 < P3; p0 += 32) {
          const int p = p0 + lane;
          if (P3 - p0 >= 32 || p < P3) {
            double u0[M], u1[M];
#pragma unroll
            for (int r = 0; r < M; ++r) { u0[r] = vst[M * p + r]; u1[r] = vst[NQ + M * p + r]; }
#pragma unroll
            for (int q = 0; q < M; ++q) {
              double g0 = T.C[0][q][0] * u0[0], g1 = T.C[0][q][0] * u1[0];
#pragma unroll
              for (int r = 1; r < M; ++r) { g0 = fma(T.C[0][q][r], u0[r], g0); g1 = fma(T.C[0][q][r], u1[r], g1); }
              jst[(M * p + q) * 3 + dFirst] = g0 * jc[0][dFirst];
              jst[(NQ + M * p + q) * 3 + dFirst] = g1 * jc[1][dFirst];
            }
          }
        }
        // ---- lines along the middle direction: pencil p = q0 + M q2, points q0 + M r + M^2 q2 ----
#pragma unroll
        for (int p0 = 0; p0 < P3; p0 += 32) {
          const int p = p0 + lane;
          if (P3 - p0 >= 32 || p < P3) {
            const int o = p % M + (p / M) * P3;
            double u0[M], u1[M];
#pragma unroll
            for (int r = 0; r < M; ++r) { u0[r] = vst[o + r * M]; u1[r] = vst[NQ + o + r * M]; }
#pragma unroll
            for (int q = 0; q < M; ++q) {
              double g0 = T.C[1][q][0] * u0[0], g1 = T.C[1][q][0] * u1[0];
#pragma unroll
              for (int r = 1; r < M; ++r) { g0 = fma(T.C[1][q][r], u0[r], g0); g1 = fma(T.C[1][q][r], u1[r], g1); }
              jst[(o + q * M) * 3 + 1] = g0 * jc[0][1];
              jst[(NQ + o + q * M) * 3 + 1] = g1 * jc[1][1];
            }
          }
        }
      }
      if (GRAD) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int j = 0; j < 3; ++j) jc[e][j] = jn[e][j];
      }
    }
    // ---- the block's tiles leave in output order: WARPS consecutive pairs, one bulk store per array ----
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const u64 first = grp * WARPS;
      const unsigned n = (unsigned)min((u64)WARPS, R.nPairs - first);
      if (values)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(values + first * (u64)(2 * NQ)),
                     "r"(sa(tileV)), "r"((unsigned)(n * 2 * NQ * sizeof(double))) : "memory");
      if (GRAD && P3 <= 32) asm volatile("cp.async.bulk.commit_group;" ::: "memory");     // EARLY: values in their own, older group
      if (GRAD)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(jac + first * (u64)(6 * NQ)),
                     "r"(sa(tileJ)), "r"((unsigned)(n * 6 * NQ * sizeof(double))) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      pending = true;
    }
  }
  // shared memory must outlive the last bulk stores' reads
  if (threadIdx.x == 0 && pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int ensureJinvTable(dfx_basis* b, cudaStream_t st);

template <int N1, int M, bool GRAD, bool PERM, int WARPS, bool BLK>
static int launchPairW(dfx_basis* b, const FastPlan& plan, const PairArgs& R, const double* src, double* values, double* jac,
                       cudaStream_t st) {
  using Lay = PairLayout<N1, M, GRAD>;
  PairTables<N1, M> T;
  for (int j = 0; j < 3; ++j) {
    const int js = PERM ? 2 - j : j;             // tables follow the contraction order
    for (int q = 0; q < M; ++q) {
      for (int a = 0; a < N1; ++a) T.A[j][q][a] = plan.A[js][q][a];
      for (int r = 0; r < M; ++r) T.C[j][q][r] = plan.Cm[js][q][r];
    }
  }
  const size_t smem = Lay::bytes(WARPS);
  auto kern = k_eval_pair<N1, M, GRAD, PERM, WARPS, BLK>;
  static std::atomic<unsigned long long> configured{0};
  if (smem > 48 * 1024) { int rc = optInSharedMemory(kern, b->device, configured); if (rc) return rc; }
  int perSM = 0, sms = 0;
  DFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, WARPS * 32, smem));
  DFX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
  const u64 want = (R.nPairs + WARPS - 1) / WARPS;
  u64 blocks = std::min<u64>(want, (u64)std::max(1, perSM) * (u64)std::max(1, sms));
  if (g_smemMaxBlocks > 0) blocks = std::min<u64>(blocks, (u64)g_smemMaxBlocks);
  kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(T, R, (const double*)b->dJinv, src, values, jac);
  return postLaunch("k_eval_pair");
}

// warps per block: 8 when two such blocks fit on an SM (DG4 with gradients: 2 x 112 KB; the block tile is then a
// whole number of 128-byte lines for every m), else 4 with at least two blocks, else not this kernel
template <int N1, int M, bool GRAD, bool PERM>
static int launchPair(dfx_basis* b, const FastPlan& plan, const PairArgs& R, const double* src, double* values, double* jac,
                      cudaStream_t st) {
  using Lay = PairLayout<N1, M, GRAD>;
  const size_t budget = 227 * 1024;
  const bool fits8 = 2 * (Lay::bytes(8) + 1024) <= budget, fits4 = 2 * (Lay::bytes(4) + 1024) <= budget;
  if ((g_smemPair == 4 || !fits8) && fits4) return launchPairW<N1, M, GRAD, PERM, 4, false>(b, plan, R, src, values, jac, st);
  // 8 warps: one bulk load per block (g_smemPair == 2: one per warp, the first version — 4-8 % slower)
  if (fits8 && g_smemPair == 2) return launchPairW<N1, M, GRAD, PERM, 8, false>(b, plan, R, src, values, jac, st);
  if (fits8) return launchPairW<N1, M, GRAD, PERM, 8, true>(b, plan, R, src, values, jac, st);
  return -1;
}

// -1: not a configuration of this kernel (the caller goes on to k_eval_smem)
int runPairEval(dfx_basis* b, const FastPlan& plan, int leafId, const double* coeff, u64 e0, u64 e1, double* values,
                double* jac, cudaStream_t st) {
  if (g_smemPair == 0 || g_smemGrad == 1) return -1;
  // values only: the one-element-at-a-time kernel with TMA pair loads is bound by the shared-memory pipe, not by its
  // stores, and runs at 24 warps per SM — it stays ahead (DG4 192^3: 2.43-2.49 ms against 2.54-2.62 ms here)
  if (!plan.jac && g_smemPair < 0) return -1;
  const DevBasis& D = b->dev;
  const DevLeaf& lf = D.leaves[leafId];
  if (D.layouts[lf.layout].kind != DFX_NODE_LAGRANGE_DG || lf.posA != 1) return -1;
  if (e1 <= e0 || (e1 - e0) % 2 != 0) return -1;
  const int n1 = plan.k + 1, m = plan.m;
  if (plan.dim != 3 || m < n1 || (plan.jac && !plan.colloc)) return -1;
  const double* src = coeff + lf.posB + e0 * (u64)(n1 * n1 * n1);
  if (((uintptr_t)src & 15) || ((uintptr_t)values & 15) || ((uintptr_t)jac & 15)) return -1;
  if ((u64)D.grid.N[0] * D.grid.N[1] * D.grid.N[2] > 0xffffffffull) return -1;
  if (plan.jac) { int rc = ensureJinvTable(b, st); if (rc) return rc; }
  PairArgs R;
  R.l2ahead = g_pairL2Ahead;
  R.nPairs = (e1 - e0) / 2;
  R.e0 = (unsigned)e0;
  R.n0 = (unsigned)D.grid.N[0]; R.n1 = (unsigned)D.grid.N[1];
  R.divN0 = FastDiv(R.n0); R.divN1 = FastDiv(R.n1);
  const bool perm = !plan.firstFastest;
#define DFX_PAIR(N1_, M_)                                                                                        \
  if (n1 == N1_ && m == M_) {                                                                                    \
    if (plan.jac) return perm ? launchPair<N1_, M_, true, true>(b, plan, R, src, values, jac, st)                \
                              : launchPair<N1_, M_, true, false>(b, plan, R, src, values, jac, st);              \
    return perm ? launchPair<N1_, M_, false, true>(b, plan, R, src, values, jac, st)                             \
                : launchPair<N1_, M_, false, false>(b, plan, R, src, values, jac, st);                           \
  }
  DFX_PAIR(3, 3) DFX_PAIR(3, 4)
  DFX_PAIR(4, 4) DFX_PAIR(4, 5)
  DFX_PAIR(5, 5) DFX_PAIR(5, 6)
  DFX_PAIR(6, 6) DFX_PAIR(6, 7)
#undef DFX_PAIR
  return -1;
}

}  // namespace dfx
