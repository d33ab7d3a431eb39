// dgstream.cu — development probe (not part of the library): the DATA MOVEMENT of the DG4 values + gradients
// evaluation without its arithmetic.  Per item of E elements a warp loads E*1000 B of coefficients (TMA bulk load),
// and stores E*1000 B of values and E*3000 B of Jacobians (TMA bulk stores, or plain 16-byte stores), in the same
// persistent-warp schedule as k_eval_smem / k_eval_pair.  Tells what the memory system gives this access pattern —
// the ceiling the contraction kernels can approach — and which chunk size / warp count / cache policy it prefers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dgstream dgstream.cu && ./dgstream
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// MODE 0: TMA load + TMA stores; 1: TMA load + plain st.global.v2.f64 from shared; 2: like 0, stores with L2 evict_first;
// 3: like 0 but loads with evict_first too; 4: stores only (no loads); 5: loads only
template <int E, int MODE>
__global__ void k_stream(const double* __restrict__ src, double* __restrict__ val, double* __restrict__ jac, u64 nItems, int warps) {
  constexpr int NL = 125 * E;
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int perWarp = NL + NL + 3 * NL + 2;
  double* pb = smem + (size_t)warp * perWarp;
  double* vst = pb + NL;
  double* jst = vst + NL;
  unsigned long long* bar = (unsigned long long*)(jst + 3 * NL);
  const u64 warpId = (u64)blockIdx.x * warps + warp, nWarps = (u64)gridDim.x * warps;
  if (warpId >= nItems) return;
  u64 pol = 0;
  if (MODE == 2 || MODE == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  unsigned it = 0;
  bool pending = false;
  for (u64 item = warpId; item < nItems; item += nWarps, ++it) {
    if (MODE != 4) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(bar)), "r"(NL * 8) : "memory");
        if (MODE == 3)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(sa(pb)),
                       "l"(src + item * NL), "r"(NL * 8), "r"(sa(bar)), "l"(pol) : "memory");
        else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(pb)),
                       "l"(src + item * NL), "r"(NL * 8), "r"(sa(bar)) : "memory");
      }
      asm volatile(
          "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(sa(bar)),
          "r"(it & 1u) : "memory");
    }
    if (MODE == 5) continue;
    if (lane == 0 && pending) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); pending = false; }
    __syncwarp();
    // a token amount of work: touch the tiles so that they depend on the load
    for (int i = lane; i < NL; i += 32) { const double c = MODE == 4 ? 1.0 : pb[i]; vst[i] = c; jst[3 * i] = c; jst[3 * i + 1] = c; jst[3 * i + 2] = c; }
    if (MODE == 1) {
      __syncwarp();
      double2* gv = (double2*)(val + item * NL);
      double2* gj = (double2*)(jac + item * 3 * NL);
      for (int i = lane; i < NL / 2; i += 32) gv[i] = ((double2*)vst)[i];
      for (int i = lane; i < 3 * NL / 2; i += 32) gj[i] = ((double2*)jst)[i];
    } else {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (MODE == 2 || MODE == 3) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(val + item * NL), "r"(sa(vst)), "r"(NL * 8), "l"(pol) : "memory");
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(jac + item * 3 * NL), "r"(sa(jst)), "r"(3 * NL * 8), "l"(pol) : "memory");
        } else {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(val + item * NL), "r"(sa(vst)), "r"(NL * 8) : "memory");
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(jac + item * 3 * NL), "r"(sa(jst)), "r"(3 * NL * 8) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        pending = true;
      }
    }
  }
  if (lane == 0 && pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int E, int MODE>
void run(const double* src, double* val, double* jac, u64 ne, int warps, int blocksPerSM) {
  const size_t smem = (size_t)warps * (5 * 125 * E + 2) * 8;
  auto kern = k_stream<E, MODE>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int perSM = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, warps * 32, smem));
  if (blocksPerSM > 0 && blocksPerSM < perSM) perSM = blocksPerSM;
  if (perSM < 1) { printf("E=%d MODE=%d warps=%d: does not fit\n", E, MODE, warps); return; }
  const u64 nItems = ne / E;
  const unsigned blocks = 148u * perSM;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f, sum = 0;
  for (int r = 0; r < 8; ++r) {
    CK(cudaEventRecord(a));
    kern<<<blocks, warps * 32, smem>>>(src, val, jac, nItems, warps);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (r >= 3) { sum += ms; if (ms < best) best = ms; }
  }
  CK(cudaGetLastError());
  const double bytes = (double)ne * 125 * 8 * (MODE == 4 ? 4 : (MODE == 5 ? 1 : 5));
  printf("{\"probe\": \"dgstream\", \"elements_per_item\": %d, \"mode\": %d, \"warps_per_block\": %d, \"blocks_per_sm\": %d, \"warps_per_sm\": %d, \"ms_mean\": %.4f, \"ms_best\": %.4f, \"GBs\": %.0f}\n",
         E, MODE, warps, perSM, warps * perSM, sum / 5, best, bytes / (sum / 5) / 1e6);
  fflush(stdout);
}

int main() {
  const u64 ne = 192ull * 192 * 192;
  double *src, *val, *jac;
  CK(cudaMalloc(&src, ne * 1000)); CK(cudaMalloc(&val, ne * 1000)); CK(cudaMalloc(&jac, ne * 3000));
  CK(cudaMemset(src, 0, ne * 1000));
  // the reference point: plain device-to-device copy of the same volume
  {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(a));
      CK(cudaMemcpyAsync(jac, val, ne * 1000, cudaMemcpyDeviceToDevice));
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (r == 2) printf("{\"probe\": \"cudaMemcpy d2d\", \"ms\": %.4f, \"GBs\": %.0f}\n", ms, 2.0 * ne * 1000 / ms / 1e6);
    }
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(a));
      CK(cudaMemsetAsync(jac, 0, ne * 3000));
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (r == 2) printf("{\"probe\": \"cudaMemset\", \"ms\": %.4f, \"GBs\": %.0f}\n", ms, 1.0 * ne * 3000 / ms / 1e6);
    }
  }
  run<2, 0>(src, val, jac, ne, 4, 4);     // the pair kernel's occupancy
  run<2, 0>(src, val, jac, ne, 4, 5);
  run<2, 0>(src, val, jac, ne, 4, 0);
  run<2, 0>(src, val, jac, ne, 8, 0);
  run<2, 1>(src, val, jac, ne, 4, 5);
  run<2, 1>(src, val, jac, ne, 4, 0);
  run<2, 2>(src, val, jac, ne, 4, 5);
  run<2, 2>(src, val, jac, ne, 4, 0);
  run<2, 3>(src, val, jac, ne, 4, 5);
  run<2, 4>(src, val, jac, ne, 4, 5);
  run<2, 5>(src, val, jac, ne, 4, 5);
  run<4, 0>(src, val, jac, ne, 4, 2);
  run<4, 0>(src, val, jac, ne, 4, 0);
  run<4, 2>(src, val, jac, ne, 4, 0);
  run<8, 0>(src, val, jac, ne, 4, 0);
  run<8, 2>(src, val, jac, ne, 4, 0);
  run<16, 0>(src, val, jac, ne, 2, 0);
  return 0;
}
