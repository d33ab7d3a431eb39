// dfx_device.cuh — device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>

#include "dfx_internal.h"

namespace dfx {

// Runtime divisor with precomputed magic (Granlund-Montgomery / libdivide "branchfree"
// scheme for 32-bit unsigned): q = (((n - t) >> 1) + t) >> s with t = mulhi(n, m).
struct FastDiv {
  unsigned d, m, s;
  FastDiv() : d(1), m(0), s(0) {}
  explicit FastDiv(unsigned div) : d(div) {
    if (div == 1) { m = 0; s = 0; return; }
    unsigned l = 32 - __builtin_clz(div - 1);           // ceil(log2(d))
    unsigned long long pw = 1ull << l;
    m = (unsigned)(((pw - div) << 32) / div + 1);
    s = l - 1;
  }
  DFX_HD unsigned div(unsigned n) const {
    if (d == 1) return n;
#if defined(__CUDA_ARCH__)
    unsigned t = __umulhi(n, m);
#else
    unsigned t = (unsigned)(((unsigned long long)n * m) >> 32);
#endif
    return (((n - t) >> 1) + t) >> s;
  }
  DFX_HD void divmod(unsigned n, unsigned& q, unsigned& r) const { q = div(n); r = n - q * d; }
};

// coordinate(j, i) = origin + i*h, EquidistantOffsetCoordinates of YaspGrid; written with
// explicit round-to-nearest intrinsics so that no FMA contraction changes the bits.
__device__ __forceinline__ double gridCoordinate(const DevGrid& g, int j, int iGlobal) {
  return __dadd_rn(g.lower[j], __dmul_rn((double)iGlobal, g.h[j]));
}

// AxisAlignedCubeGeometry::global on element iGlobal: lower + xhat*(upper-lower)
__device__ __forceinline__ double elementGlobal(const DevGrid& g, int j, int iGlobal, double xhat) {
  const double lo = gridCoordinate(g, j, iGlobal);
  const double up = gridCoordinate(g, j, iGlobal + 1);
  return __dadd_rn(lo, __dmul_rn(xhat, __dsub_rn(up, lo)));
}

// AxisAlignedCubeGeometry::jacobianInverse diagonal entry: 1/(upper-lower)
__device__ __forceinline__ double elementJacInv(const DevGrid& g, int j, int iGlobal) {
  const double lo = gridCoordinate(g, j, iGlobal);
  const double up = gridCoordinate(g, j, iGlobal + 1);
  return __ddiv_rn(1.0, __dsub_rn(up, lo));
}

// local index -> lattice multi-index alpha, LagrangeCube numbering: i = sum alpha_j (k+1)^j
__device__ __forceinline__ void localAlpha(int k, int i, int a[3]) {
  const int n1 = k + 1;
  a[0] = i % n1; i /= n1;
  a[1] = i % n1; i /= n1;
  a[2] = i;
}

// ---- registry of functions f (include/dfx.h, dfx_fn_id) ----------------------
// Functor interface used by the interpolation kernels (also what a user-supplied device
// functor implements): `common(x0,x1,x2)` computes the part shared by all range entries
// (the transcendental), `comp(c, common, x0,x1,x2)` returns range entry c.  No arrays, so
// everything stays in registers.
struct RegistryFn {
  dfx_function f;
  int dim;
  __device__ __forceinline__ int ncomp() const { return f.n_comp; }
  __device__ __forceinline__ double norm2(double x0, double x1, double x2) const {
    double s = __dmul_rn(x0, x0);                                   // x.two_norm2(): sum in index order
    if (dim > 1) s = __dadd_rn(s, __dmul_rn(x1, x1));
    if (dim > 2) s = __dadd_rn(s, __dmul_rn(x2, x2));
    return s;
  }
  __device__ __forceinline__ double common(double x0, double x1, double x2) const {
    switch (f.id) {
      case DFX_FN_GAUSSIAN:
      case DFX_FN_GAUSSIAN_VEC:
        return exp(__dmul_rn(-f.p[0], norm2(x0, x1, x2)));
      case DFX_FN_QUADRATIC: {
        const double xx[3] = {x0, dim > 1 ? x1 : 0.0, dim > 2 ? x2 : 0.0};
        double r = f.p[0];
#pragma unroll
        for (int j = 0; j < 3; ++j) r = __dadd_rn(r, __dmul_rn(f.p[1 + j], xx[j]));
        int q = 4;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = i; j < 3; ++j) { r = __dadd_rn(r, __dmul_rn(__dmul_rn(f.p[q], xx[i]), xx[j])); ++q; }
        return r; }
      case DFX_FN_SINPROD: {
        double r = __dmul_rn(1.0, sin(__dmul_rn(f.p[0], x0)));
        if (dim > 1) r = __dmul_rn(r, sin(__dmul_rn(f.p[0], x1)));
        if (dim > 2) r = __dmul_rn(r, sin(__dmul_rn(f.p[0], x2)));
        return r; }
      case DFX_FN_COORD: {
        const int j = (int)f.p[0];
        return j == 0 ? x0 : (j == 1 ? x1 : x2); }
    }
    return 0.0;
  }
  __device__ __forceinline__ double comp(int c, double cm, double x0, double x1, double x2) const {
    switch (f.id) {
      case DFX_FN_CONSTANT: return f.p[c];
      case DFX_FN_IDENTITY: return c >= dim ? 0.0 : (c == 0 ? x0 : (c == 1 ? x1 : x2));
      case DFX_FN_GAUSSIAN_VEC: return __dmul_rn(c >= dim ? 1.0 : (c == 0 ? x0 : (c == 1 ? x1 : x2)), cm);
    }
    return cm;
  }
};

}  // namespace dfx
