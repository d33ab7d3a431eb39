// user_functor.cu — form (3) of f (include/dfx/device_functor.cuh): a caller-defined __device__
// functor interpolated into a Taylor-Hood basis (velocity <- rotation field, pressure <- a
// polynomial), checked against a host evaluation at the library's node positions and against
// the registry path.  Compiled with nvcc, needs a GPU.
#include <cmath>
#include <cstdio>
#include <vector>

#include <dfx/device_functor.cuh>

struct Rotation {                     // v(x) = (-x1, x0, x2 * x0)
  __device__ double operator()(const double* x, int, int comp) const {
    return comp == 0 ? -x[1] : (comp == 1 ? x[0] : x[2] * x[0]);
  }
};
struct Poly {                         // p(x) = 1 + 2 x0 - x1 x2
  __host__ __device__ double operator()(const double* x, int, int) const { return 1 + 2 * x[0] - x[1] * x[2]; }
};

#define CK(call) do { int rc_ = (call); if (rc_ != 0) { std::printf("FAILED %s -> %d (%s)\n", #call, rc_, dfx_last_error()); return 1; } } while (0)

int main() {
  dfx_grid_desc g{};
  g.dim = 3;
  const int N[3] = {6, 5, 4};
  for (int j = 0; j < 3; ++j) { g.n_global[j] = g.local_n[j] = N[j]; g.local_begin[j] = 0; g.lower[j] = 0; g.upper[j] = 1 + j; }
  dfx_tree_node tree[1] = {{DFX_NODE_TAYLOR_HOOD, 0, 0, DFX_IMS_NONE}};
  dfx_basis* b = nullptr;
  CK(dfx_basis_create(&g, tree, 1, 0, &b));
  const size_t n = dfx_basis_dimension(b);
  double *xyz, *coeff; int32_t* leaf;
  cudaMalloc(&xyz, n * 3 * sizeof(double)); cudaMalloc(&coeff, n * sizeof(double)); cudaMalloc(&leaf, n * sizeof(int32_t));
  cudaMemset(coeff, 0, n * sizeof(double));
  const int velocity = dfx_basis_node_child(b, 0, 0), pressure = dfx_basis_node_child(b, 0, 1);
  CK(dfx::interpolate(b, velocity, Rotation{}, false, coeff, nullptr, xyz, leaf, false, 3, 0));
  CK(dfx::interpolate(b, pressure, Poly{}, true, coeff, nullptr, xyz, leaf, true, 3, 0));
  std::vector<double> hx(n * 3), hc(n); std::vector<int32_t> hl(n);
  cudaMemcpy(hx.data(), xyz, n * 3 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(hc.data(), coeff, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(hl.data(), leaf, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (size_t p = 0; p < n; ++p) {
    const double* x = &hx[3 * p];
    double expect = hl[p] == 0 ? -x[1] : hl[p] == 1 ? x[0] : hl[p] == 2 ? x[2] * x[0] : Poly{}(x, 3, 0);
    if (std::abs(hc[p] - expect) > 1e-15 * (1 + std::abs(expect))) ++bad;
  }
  if (bad) { std::printf("FAILED %d coefficients differ\n", bad); return 1; }
  // the velocity DOFs interpolated with the registry's identity must equal x at the same nodes
  dfx_function id{}; id.id = DFX_FN_IDENTITY; id.n_comp = 3;
  CK(dfx_interpolate(b, velocity, &id, coeff, nullptr, nullptr));
  cudaMemcpy(hc.data(), coeff, n * sizeof(double), cudaMemcpyDeviceToHost);
  for (size_t p = 0; p < n; ++p)
    if (hl[p] < 3 && hc[p] != hx[3 * p + hl[p]]) ++bad;
  if (bad) { std::printf("FAILED registry identity differs at %d nodes\n", bad); return 1; }
  dfx_basis_destroy(b);
  std::printf("user_functor passed (%zu coefficients)\n", n);
  return 0;
}
