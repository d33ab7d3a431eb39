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
  const unsigned char* orient;   // ORIENT: orientation code per edge / facet of the local box (DevBasis::orient)
  unsigned ooff[8];
};

// Orientation codes of every edge and quadrilateral facet of the local box from the vertex ids (orientCode,
// dfx_internal.h; FaceOrientations, lagrangebasis.hh:207-331): one thread per entity, one byte each.
struct OrientArgs { unsigned off[8], cnt[8][3]; unsigned total; };
__global__ void __launch_bounds__(256)
k_orient_codes(const __grid_constant__ DevGrid g, const __grid_constant__ OrientArgs A, const u64* __restrict__ vid,
               unsigned char* __restrict__ tab) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.total) return;
  unsigned s = 0;                     // offsets ascend with the component number over the components that exist
#pragma unroll
  for (unsigned q = 1; q < 7; ++q)
    if (A.cnt[q][0] != 0 && t >= A.off[q]) s = q;
  const unsigned r = t - A.off[s];
  int c[3];
  c[0] = (int)(r % A.cnt[s][0]);
  const unsigned r2 = r / A.cnt[s][0];
  c[1] = (int)(r2 % A.cnt[s][1]);
  c[2] = (int)(r2 / A.cnt[s][1]);
  tab[t] = (unsigned char)orientCode(g, vid, s, c);
}

// all eight codes of a facet (edges: two) applied to the in-entity indices of order 3, two bits each: the
// permuted index of `fr` under `code` is (word >> (8 code + 2 fr)) & 3
constexpr unsigned long long q3OrientWord(int nfree) {
  unsigned long long w = 0;
  for (unsigned code = 0; code < 8; ++code)
    for (unsigned fr = 0; fr < (nfree == 2 ? 4u : 2u); ++fr)
      w |= (unsigned long long)applyOrientCode<false>(nfree == 1 ? (code & 1u) : code, nfree, 3, fr) << (8 * code + 2 * fr);
  return w;
}
constexpr unsigned long long kQ3WordEdge = q3OrientWord(1), kQ3WordFace = q3OrientWord(2);

// UNIT: flat container position = scalar index + posB (a root leaf, or a blocked-lexicographic child)
// ORIENT (dfx_basis_set_vertex_ids): the DOFs of an edge / facet sit in the entity's globally unique orientation
// (lagrangebasis.hh:863-869); the thread reads the entity's code and permutes the in-entity index of its load.
template <int M, bool UNIT, bool ORIENT>
__global__ void __launch_bounds__(kQ3Threads, ORIENT ? 5 : 6)
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
        if (ORIENT && blk > 1 && blk < 8) {
          const unsigned long long word = blk == 4 ? kQ3WordFace : kQ3WordEdge;
          const unsigned code = __ldg(R.orient + R.ooff[s] + ent);
          fr = (int)((word >> (8u * code + 2u * (unsigned)fr)) & 3ull);
        }
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

// Builds (once per set of vertex ids) the orientation codes of the local box and publishes them in the handle's
// DevBasis.  Stream-ordered on `st` like every other per-handle cache (one stream per handle, include/dfx.h).
int ensureOrientTable(dfx_basis* b, cudaStream_t st) {
  if (b->orientValid && b->dev.orient) return DFX_OK;
  if (!b->dev.vid) return fail(DFX_ERR_INVALID, "orientation table without vertex ids");
  const DevGrid& g = b->dev.grid;
  OrientArgs A{};
  u64 total = 0;
  for (unsigned s = 1; s < 7; ++s) {
    const int nfree = (int)((s & 1u) + ((s >> 1) & 1u) + ((s >> 2) & 1u));
    bool exists = (nfree == 1 && g.dim > 1) || (nfree == 2 && g.dim == 3);
    for (int j = g.dim; j < 3; ++j) if ((s >> j) & 1u) exists = false;       // no free direction beyond the dimension
    if (!exists) continue;
    u64 n = 1;
    for (int j = 0; j < 3; ++j) { A.cnt[s][j] = (unsigned)(j < g.dim ? g.N[j] + (((s >> j) & 1u) ? 0 : 1) : 1); n *= A.cnt[s][j]; }
    A.off[s] = (unsigned)total;
    total += n;
  }
  if (total == 0 || total > 0xffffffffull) return fail(DFX_ERR_RANGE, "orientation table: entity count out of range");
  A.total = (unsigned)total;
  if (b->dOrientBytes < total) {
    if (b->dOrient) DFX_CUDA(cudaFree(b->dOrient));
    b->dOrient = nullptr; b->dOrientBytes = 0;
    DFX_CUDA(cudaMalloc(&b->dOrient, total));
    b->dOrientBytes = total;
  }
  k_orient_codes<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, A, b->dev.vid, (unsigned char*)b->dOrient);
  int rc = postLaunch("k_orient_codes");
  if (rc) return rc;
  b->dev.orient = (const unsigned char*)b->dOrient;
  for (int s = 0; s < 8; ++s) b->dev.orientOff[s] = A.off[s];
  b->orientValid = true;
  if (b->dDev) DFX_CUDA(cudaMemcpyAsync(b->dDev, &b->dev, sizeof(DevBasis), cudaMemcpyHostToDevice, st));
  return DFX_OK;
}

// -1: not a configuration of this kernel (the caller goes on to k_eval_smem, or to the gather + companion route
// when the handle has permuted face DOFs)
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
  Q3Args R{};
  if (plan.firstFastest) { R.qs[0] = 1; R.qs[1] = M; R.qs[2] = M * M; }
  else { R.qs[2] = 1; R.qs[1] = M; R.qs[0] = M * M; }
  R.e0 = e0; R.ne = e1 - e0;
  R.divN0 = FastDiv((unsigned)D.grid.N[0]); R.divN1 = FastDiv((unsigned)D.grid.N[1]);
  R.posA = lf.posA; R.posB = lf.posB;
  if (b->oriented) {
    int rc = ensureOrientTable(b, st);
    if (rc) return rc;
    // the table numbers entities like the DOF layout does
    for (unsigned s = 1; s < 7; ++s)
      for (int j = 0; j < 3; ++j)
        if ((int)(D.grid.N[j] + (((s >> j) & 1u) ? 0 : 1)) != L.csize[s][j]) return -1;
    R.orient = b->dev.orient;
    for (int s = 0; s < 8; ++s) R.ooff[s] = b->dev.orientOff[s];
  }
  const size_t smem = (size_t)kQ3Threads * ((M * M * M) | 1) * sizeof(double);
  const u64 blocks = (R.ne + kQ3Threads - 1) / kQ3Threads;
  if (blocks > 0x7fffffffull) return fail(DFX_ERR_RANGE, "element range too large for one launch");
  // UNIT multiplies entity index and block size in 32 bits: every component's DOF count must fit
  bool unit = lf.posA == 1;
  for (int s = 0; s < 8; ++s) if (L.compCount[s] > 0xffffffffull) unit = false;
  const unsigned nb = (unsigned)blocks;
  if (b->oriented) {
    if (unit) k_eval_q3<M, true, true><<<nb, kQ3Threads, smem, st>>>(L, T, R, coeff, values);
    else k_eval_q3<M, false, true><<<nb, kQ3Threads, smem, st>>>(L, T, R, coeff, values);
  } else {
    if (unit) k_eval_q3<M, true, false><<<nb, kQ3Threads, smem, st>>>(L, T, R, coeff, values);
    else k_eval_q3<M, false, false><<<nb, kQ3Threads, smem, st>>>(L, T, R, coeff, values);
  }
  return postLaunch("k_eval_q3");
}

}  // namespace dfx
