// dgstream2.cu — development probe, second round: why do 2 KB + 6 KB bulk stores per warp reach only 5.3 TB/s when a
// memset reaches 7.4?  Variants of the STORE pattern of the DG4 values + gradients evaluation (loads optional):
//   A  per-warp stores (as k_eval_pair), grid-stride items
//   B  per-warp stores, chunk padded to a multiple of 1024 B (diagnostic: alignment)
//   C  block-cooperative: the W warps of a block fill one block tile, ONE bulk store per array per block iteration
//   D  per-warp stores, every block walks a contiguous run of items
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dgstream2 dgstream2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulkStore(double* g, const double* s, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(sa(s)), "r"(bytes) : "memory");
}

// VAR 0 = A, 1 = B (padded), 3 = D (contiguous run per block); LOAD: also TMA-load the coefficients
template <int VAR, bool LOAD>
__global__ void k_warp(const double* __restrict__ src, double* __restrict__ val, double* __restrict__ jac, u64 nItems, int warps) {
  constexpr int NL = 250;                       // one pair
  constexpr int NV = VAR == 1 ? 256 : 250, NJ = VAR == 1 ? 768 : 750;
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int perWarp = 256 + 256 + 768 + 2;
  double* pb = smem + (size_t)warp * perWarp;
  double* vst = pb + 256;
  double* jst = vst + 256;
  unsigned long long* bar = (unsigned long long*)(jst + 768);
  u64 item, step, end;
  if (VAR == 3) {
    const u64 per = (nItems + gridDim.x - 1) / gridDim.x;
    item = (u64)blockIdx.x * per + warp; step = warps; end = min(nItems, (u64)(blockIdx.x + 1) * per);
  } else {
    item = (u64)blockIdx.x * warps + warp; step = (u64)gridDim.x * warps; end = nItems;
  }
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  unsigned it = 0;
  bool pending = false;
  for (; item < end; item += step, ++it) {
    if (LOAD) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(bar)), "r"(NL * 8) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(pb)),
                     "l"(src + item * NL), "r"(NL * 8), "r"(sa(bar)) : "memory");
      }
      asm volatile(
          "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(sa(bar)),
          "r"(it & 1u) : "memory");
    }
    if (lane == 0 && pending) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); pending = false; }
    __syncwarp();
    for (int i = lane; i < NL; i += 32) { const double c = LOAD ? pb[i] : 1.0; vst[i] = c; jst[3 * i] = c; jst[3 * i + 1] = c; jst[3 * i + 2] = c; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      bulkStore(val + item * NV, vst, NV * 8);
      bulkStore(jac + item * NJ, jst, NJ * 8);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      pending = true;
    }
  }
  if (lane == 0 && pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// C: block tile of W pairs; BUF staging buffers per block (1: wait for the store's read before refilling)
template <bool LOAD, int BUF>
__global__ void k_block(const double* __restrict__ src, double* __restrict__ val, double* __restrict__ jac, u64 nItems, int warps) {
  constexpr int NL = 250;
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* pbAll = smem;                               // [W][250]
  double* vstAll = pbAll + (size_t)warps * NL;        // [BUF][W][250]
  double* jstAll = vstAll + (size_t)BUF * warps * NL; // [BUF][W][750]
  unsigned long long* bars = (unsigned long long*)(jstAll + (size_t)BUF * warps * 3 * NL);
  double* pb = pbAll + warp * NL;
  unsigned long long* bar = bars + warp;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const u64 nGroups = nItems / warps;                 // probe: nItems divisible
  unsigned it = 0;
  int pend = 0;
  for (u64 grp = blockIdx.x; grp < nGroups; grp += gridDim.x, ++it) {
    const u64 item = grp * warps + warp;
    const int buf = BUF == 1 ? 0 : (int)(it % BUF);
    double* vst = vstAll + ((size_t)buf * warps + warp) * NL;
    double* jst = jstAll + ((size_t)buf * warps + warp) * 3 * NL;
    if (LOAD) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(bar)), "r"(NL * 8) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(pb)),
                     "l"(src + item * NL), "r"(NL * 8), "r"(sa(bar)) : "memory");
      }
      asm volatile(
          "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(sa(bar)),
          "r"(it & 1u) : "memory");
    }
    // the store that last used this buffer must have read it
    if (threadIdx.x == 0 && pend >= BUF) {
      if (BUF == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      pend = BUF - 1;
    }
    __syncthreads();
    for (int i = lane; i < NL; i += 32) { const double c = LOAD ? pb[i] : 1.0; vst[i] = c; jst[3 * i] = c; jst[3 * i + 1] = c; jst[3 * i + 2] = c; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      bulkStore(val + grp * warps * NL, vstAll + (size_t)buf * warps * NL, warps * NL * 8);
      bulkStore(jac + grp * warps * 3 * NL, jstAll + (size_t)buf * warps * 3 * NL, warps * 3 * NL * 8);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      ++pend;
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <class K>
void timeKernel(const char* name, K kern, size_t smem, int warps, int blocksPerSM, bool load, const double* src, double* val,
                double* jac, u64 nItems, double bytes) {
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int perSM = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, warps * 32, smem));
  if (blocksPerSM > 0 && blocksPerSM < perSM) perSM = blocksPerSM;
  if (perSM < 1) { printf("%s: does not fit\n", name); return; }
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f, sum = 0;
  for (int r = 0; r < 8; ++r) {
    CK(cudaEventRecord(a));
    kern<<<148u * perSM, warps * 32, smem>>>(src, val, jac, nItems, warps);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (r >= 3) { sum += ms; if (ms < best) best = ms; }
  }
  CK(cudaGetLastError());
  printf("{\"probe\": \"%s\", \"loads\": %d, \"warps_per_block\": %d, \"blocks_per_sm\": %d, \"ms_mean\": %.4f, \"ms_best\": %.4f, \"GBs\": %.0f}\n",
         name, (int)load, warps, perSM, sum / 5, best, bytes / (sum / 5) / 1e6);
  fflush(stdout);
}

int main() {
  const u64 ne = 192ull * 192 * 192, nItems = ne / 2;
  double *src, *val, *jac;
  CK(cudaMalloc(&src, ne * 1000)); CK(cudaMalloc(&val, ne * 1100)); CK(cudaMalloc(&jac, ne * 3300));
  CK(cudaMemset(src, 0, ne * 1000));
  const double bS = (double)ne * 4000, bLS = (double)ne * 5000;
  const size_t wsm = (256 + 256 + 768 + 2) * 8;
  timeKernel("A per-warp stores", k_warp<0, false>, 4 * wsm, 4, 5, false, src, val, jac, nItems, bS);
  timeKernel("B per-warp stores padded to 2048/6144 B", k_warp<1, false>, 4 * wsm, 4, 5, false, src, val, jac, nItems, bS);
  timeKernel("D per-warp stores, contiguous run per block", k_warp<3, false>, 4 * wsm, 4, 5, false, src, val, jac, nItems, bS);
  timeKernel("A per-warp stores", k_warp<0, true>, 4 * wsm, 4, 5, true, src, val, jac, nItems, bLS);
  timeKernel("B per-warp stores padded to 2048/6144 B", k_warp<1, true>, 4 * wsm, 4, 5, true, src, val, jac, nItems, bLS);
  timeKernel("D per-warp stores, contiguous run per block", k_warp<3, true>, 4 * wsm, 4, 5, true, src, val, jac, nItems, bLS);
  for (int w : {4, 8, 16}) {
    const size_t s1 = ((size_t)w * 250 * (1 + 4) + w) * 8, s2 = ((size_t)w * 250 * (1 + 8) + w) * 8;
    timeKernel("C block tile, 1 buffer", k_block<false, 1>, s1, w, 0, false, src, val, jac, nItems, bS);
    timeKernel("C block tile, 1 buffer", k_block<true, 1>, s1, w, 0, true, src, val, jac, nItems, bLS);
    timeKernel("C block tile, 2 buffers", k_block<true, 2>, s2, w, 0, true, src, val, jac, nItems, bLS);
  }
  timeKernel("C block tile, 1 buffer, 2 blocks/SM", k_block<true, 1>, ((size_t)8 * 250 * 5 + 8) * 8, 8, 2, true, src, val, jac, nItems, bLS);
  timeKernel("C block tile, 1 buffer, 1 block/SM", k_block<true, 1>, ((size_t)16 * 250 * 5 + 16) * 8, 16, 1, true, src, val, jac, nItems, bLS);
  return 0;
}
