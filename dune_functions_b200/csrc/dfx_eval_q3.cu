// dfx_eval_q3.cu — continuous Lagrange order 3 in three dimensions at 4^3 tensor points, values only: the
// x-blocked thread-per-element mapping of k_eval_reg (dfx_eval_fast.cu) carried to 64 coefficients per element.
//
// k_eval_smem (one element per warp, one pencil per lane) runs Q3 on 128^3 elements at 0.33 of the copy peak: 16
// pencils keep half of the lanes busy, every lane computes the position of its coefficients in one of the eight
// index-set components at run time, and a pair of elements leaves as a 1 KB bulk store.  Here
//  * consecutive threads own consecutive elements in x, so a warp's load of local node (a0, a1, a2) reads one run
//    of one component — vertices with unit stride, the two DOFs of an edge, the four of a facet and the eight of
//    the cell at strides of 2, 4 and 8 doubles, the neighbouring in-entity loads filling the same sectors;
//  * every offset (component, entity size, in-entity index — LagrangeCube LocalKey::index in base k-1) is a
//    compile-time function of the local node, the component table sits in the constant bank;
//  * a thread issues its 64 loads back to back (64 x 8 bytes x 64 threads x the resident blocks in flight) and
//    contracts direction by direction in registers (3 x 256 FMA instead of 4 096);
//  * the block's 64 x 64 results go through a shared-memory tile with an odd row stride (conflict-free for the
//    per-thread rows as well as for the copy-out) and leave as coalesced 8-byte stores, 256 contiguous bytes per warp
//    instruction, 32 KB contiguous per block.
//
// Replaces LocalFunctionBase::bind + LocalFunction::operator() (discreteglobalbasisfunction.hh:124-148, :336-371)
// for LagrangePreBasis<GV, 3> on a structured grid without permuted face DOFs (lagrangebasis.hh:843-873 reduces to
// the closed form of DESIGN.md section 3 when the vertex ids are monotone).
#include <algorithm>

#include "dfx_device.cuh"
#include "dfx_eval_fast.h"
#include "dfx_launch.h"
#include "dfx_sumfactor.cuh"

namespace dfx {

int g_evalQ3 = 1;        // DFX_OP_TUNE_EVAL_Q3: 0 = order 3 stays on the shared-memory kernel

namespace {

constexpr int kQ3Threads = 64;

template <int M>
struct Q3Tables { double A[3][M][4]; };

struct Q3Args {
  int qs[3];                   // stride of q_j in the point index
  unsigned long long e0, ne;
  FastDiv divN0, divN1;
  unsigned long long posA, posB;
};

// UNIT: flat container position = scalar index + posB (a root leaf, or a blocked-lexicographic child)
template <int M, bool UNIT>
__global__ void __launch_bounds__(kQ3Threads, 6)
k_eval_q3(const __grid_constant__ DevLayout L, const __grid_constant__ Q3Tables<M> T, const __grid_constant__ Q3Args R,
          const double* __restrict__ coeff, double* __restrict__ values) {
  constexpr int K = 3, N1 = 4, NLOC = 64, NQ = M * M * M;
  constexpr int STR = NQ | 1;                              // odd row stride of the staging tile
  extern __shared__ __align__(16) double tile[];           // [kQ3Threads][STR]
  const int tid = threadIdx.x;
  const u64 blk0 = (u64)blockIdx.x * kQ3Threads;
  const u64 el = blk0 + tid;
  const bool active = el < R.ne;
  const u64 e = R.e0 + (active ? el : 0);
  unsigned ec0, ec1, ec2;
  {
    unsigned q;
    R.divN0.divmod((unsigned)e, q, ec0);                   // numElements < 2^32 checked on the host
    R.divN1.divmod(q, ec2, ec1);
  }
  double c[NLOC];
  const double* src = coeff + R.posB;
#pragma unroll
  for (int a2 = 0; a2 < N1; ++a2)
#pragma unroll
    for (int a1 = 0; a1 < N1; ++a1)
#pragma unroll
      for (int a0 = 0; a0 < N1; ++a0) {
        // closed form of MCMGMapper::subIndex + LocalKey::index on the structured grid
        const bool i0 = a0 > 0 && a0 < K, i1 = a1 > 0 && a1 < K, i2 = a2 > 0 && a2 < K;
        const int s = (i0 ? 1 : 0) | (i1 ? 2 : 0) | (i2 ? 4 : 0);
        const int blk = (i0 ? 2 : 1) * (i1 ? 2 : 1) * (i2 ? 2 : 1);
        int fr = 0, frs = 1;
        if (i0) { fr += (a0 - 1) * frs; frs *= 2; }
        if (i1) { fr += (a1 - 1) * frs; frs *= 2; }
        if (i2) { fr += (a2 - 1) * frs; }
        const unsigned c0 = ec0 + (a0 == K ? 1u : 0u), c1 = ec1 + (a1 == K ? 1u : 0u), c2 = ec2 + (a2 == K ? 1u : 0u);
        const unsigned ent = c0 + (unsigned)L.csize[s][0] * (c1 + (unsigned)L.csize[s][1] * c2);
        if (UNIT) {
          c[a0 + N1 * (a1 + N1 * a2)] = __ldg(src + L.compOffset[s] + fr + (u64)(ent * (unsigned)blk));
        } else {
          const u64 j = L.compOffset[s] + (u64)ent * (u64)blk + (u64)fr;
          c[a0 + N1 * (a1 + N1 * a2)] = __ldg(src + R.posA * j);
        }
      }
  double u[NQ];
  sumFactor<3, N1, M>(T.A[0], T.A[1], T.A[2], c, u);
  double* row = tile + tid * STR;
#pragma unroll
  for (int q2 = 0; q2 < M; ++q2)
#pragma unroll
    for (int q1 = 0; q1 < M; ++q1)
#pragma unroll
      for (int q0 = 0; q0 < M; ++q0) row[q0 * R.qs[0] + q1 * R.qs[1] + q2 * R.qs[2]] = u[q0 + M * (q1 + M * q2)];
  __syncthreads();
  // the block's results in output order: element-major, 256 contiguous bytes per warp instruction
  const unsigned nAct = (unsigned)(R.ne - blk0 < (u64)kQ3Threads ? R.ne - blk0 : (u64)kQ3Threads);
  double* dst = values + blk0 * (u64)NQ;
#pragma unroll 8
  for (unsigned i = tid; i < nAct * NQ; i += kQ3Threads) {
    const unsigned r = i / NQ, q = i - r * NQ;
    __stcs(dst + i, tile[r * STR + q]);
  }
}

}  // namespace

// -1: not a configuration of this kernel (the caller goes on to k_eval_smem)
int runQ3Eval(dfx_basis* b, const FastPlan& plan, int leafId, const double* coeff, u64 e0, u64 e1, double* values,
              cudaStream_t st) {
  if (!g_evalQ3 || plan.dim != 3 || plan.k != 3 || plan.m != 4 || plan.jac || !values) return -1;
  const DevBasis& D = b->dev;
  const DevLeaf& lf = D.leaves[leafId];
  const DevLayout& L = D.layouts[lf.layout];
  if (L.kind != DFX_NODE_LAGRANGE || L.k != 3) return -1;
  if ((u64)D.grid.N[0] * D.grid.N[1] * D.grid.N[2] > 0xffffffffull) return -1;
  for (int s = 0; s < 8; ++s)                                   // entity counts of one component in 32 bits
    if ((u64)L.csize[s][0] * L.csize[s][1] * L.csize[s][2] > 0xffffffffull) return -1;
  if (e1 == e0) return DFX_OK;
  constexpr int M = 4;
  Q3Tables<M> T;
  for (int j = 0; j < 3; ++j)
    for (int q = 0; q < M; ++q)
      for (int a = 0; a < 4; ++a) T.A[j][q][a] = plan.A[j][q][a];
  Q3Args R;
  if (plan.firstFastest) { R.qs[0] = 1; R.qs[1] = M; R.qs[2] = M * M; }
  else { R.qs[2] = 1; R.qs[1] = M; R.qs[0] = M * M; }
  R.e0 = e0; R.ne = e1 - e0;
  R.divN0 = FastDiv((unsigned)D.grid.N[0]); R.divN1 = FastDiv((unsigned)D.grid.N[1]);
  R.posA = lf.posA; R.posB = lf.posB;
  const size_t smem = (size_t)kQ3Threads * ((M * M * M) | 1) * sizeof(double);
  const u64 blocks = (R.ne + kQ3Threads - 1) / kQ3Threads;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  // UNIT multiplies entity index and block size in 32 bits: every component's DOF count must fit
  bool unit = lf.posA == 1;
  for (int s = 0; s < 8; ++s) if (L.compCount[s] > 0xffffffffull) unit = false;
  if (unit) k_eval_q3<M, true><<<(unsigned)blocks, kQ3Threads, smem, st>>>(L, T, R, coeff, values);
  else k_eval_q3<M, false><<<(unsigned)blocks, kQ3Threads, smem, st>>>(L, T, R, coeff, values);
  return postLaunch("k_eval_q3");
}

}  // namespace dfx
