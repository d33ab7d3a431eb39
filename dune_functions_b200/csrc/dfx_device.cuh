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

// ---- exp for the registry's Gaussians -----------------------------------------------
// interpolate() evaluates f at EVERY Lagrange node (interpolate.hh:151-173), so exp(-a|x|^2) runs
// once per DOF and the FP64 pipe, not HBM, bounds the kernel with the library exp (about 35 FP64
// issues per call).  This one is 10 FP64 issues and one cached table load: x = (64 k + j) ln2/64 + r,
// |r| <= ln2/128, exp(x) = 2^k * 2^(j/64) * exp(r) with a Cody-Waite two-step reduction (k ln2/64
// exact in the high part), the degree-5 Taylor polynomial of expm1(r) (truncation 3.5e-17) and a
// correctly rounded 64-entry table.  Worst error 1.3 ulp over [-50, 50] (checked against 50-digit
// arithmetic, tests/test_host_cpu.py), the same class as the CUDA library's 1 ulp and glibc's < 1 ulp;
// arguments beyond +-700 (sub-normal / overflow range) take the library routine.
static __device__ const double kExp2Tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// constants that do not fit a DFMA immediate come from the constant bank as direct operands (no
// per-use UMOV pairs in the node loop)
static __constant__ double kExpC[6] = {0x1.71547652b82fep+6 /* 64/ln2 */, -0x1.62e42fee00000p-7 /* -ln2/64, high 32 bits */,
                                       -0x1.a39ef35793c76p-39 /* -ln2/64, rest */, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};

// CHECKED = false: the caller guarantees |x| <= 700 (the host bounds the argument over the grid box)
template <bool CHECKED>
__device__ __forceinline__ double expTabT(double x) {
  if (CHECKED && !(fabs(x) <= 700.0)) return exp(x);            // NaN, infinities, sub-normal / overflow range
  constexpr double kMagic = 6755399441055744.0;                 // 1.5 * 2^52: rounds to nearest integer
  const double t = fma(x, kExpC[0], kMagic);
  const int ki = __double2loint(t);                             // 64 k + j
  const double kd = t - kMagic;
  double r = fma(kd, kExpC[1], x);                              // the product is exact
  r = fma(kd, kExpC[2], r);
  const double w = r * r;
  double q = fma(r, kExpC[3], kExpC[4]);
  q = fma(q, r, kExpC[5]);
  q = fma(q, r, 0.5);
  const double pm1 = fma(q, w, r);                              // expm1(r)
  const double tj = __ldg(kExp2Tab + (ki & 63));
  const double y = fma(tj, pm1, tj);                            // in [1, 2): scaling by 2^k is an exponent add
  return __hiloint2double(__double2hiint(y) + ((ki >> 6) << 20), __double2loint(y));
}
__device__ __forceinline__ double expTab(double x) { return expTabT<true>(x); }

// ---- registry of functions f (include/dfx.h, dfx_fn_id) ----------------------
// Functor interface used by the interpolation kernels (also what a user-supplied device
// functor implements): `common(x0,x1,x2)` computes the part shared by all range entries
// (the transcendental), `comp(c, common, x0,x1,x2)` returns range entry c.  No arrays, so
// everything stays in registers.
// ID >= 0: the function id is a compile-time constant (the switches below fold away: no indirect branch
// and no dead transcendental code in the node loop); ID = -1 dispatches on f.id at run time.
// LEAN (structured per-node kernels only): the caller passes 0.0 for coordinates beyond the grid's
// dimension — adding their squares / products changes no bit, so the dimension tests go — and the
// host has bounded the Gaussian's argument over the grid box by 700 (expTabT<false>).
template <int ID, bool LEAN = false>
struct RegistryFnT {
  dfx_function f;
  int dim;
  __device__ __forceinline__ int id() const { return ID >= 0 ? ID : f.id; }
  __device__ __forceinline__ int ncomp() const { return f.n_comp; }
  __device__ __forceinline__ double norm2(double x0, double x1, double x2) const {
    double s = __dmul_rn(x0, x0);                                   // x.two_norm2(): sum in index order
    if (LEAN || dim > 1) s = __dadd_rn(s, __dmul_rn(x1, x1));
    if (LEAN || dim > 2) s = __dadd_rn(s, __dmul_rn(x2, x2));
    return s;
  }
  __device__ __forceinline__ double common(double x0, double x1, double x2) const {
    switch (id()) {
      case DFX_FN_GAUSSIAN:
      case DFX_FN_GAUSSIAN_VEC:
        return expTabT<!LEAN>(__dmul_rn(-f.p[0], norm2(x0, x1, x2)));
      case DFX_FN_QUADRATIC: {
        const double xx[3] = {x0, (LEAN || dim > 1) ? x1 : 0.0, (LEAN || dim > 2) ? x2 : 0.0};
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
  // `common` from the SQUARES of the coordinates, for the functions that only need those (the Gaussians): the lattice
  // sweeps keep x0^2 per x-entity and x1^2, x2^2 per row, so a node costs the two additions, not three more
  // multiplications — the same operations in the same order as `common`, hence the same bits.
  static constexpr bool kFromSquares = ID == DFX_FN_GAUSSIAN || ID == DFX_FN_GAUSSIAN_VEC;
  __device__ __forceinline__ double commonSq(double q0, double q1, double q2) const {
    double s = q0;                                                  // x.two_norm2(): sum in index order
    if (LEAN || dim > 1) s = __dadd_rn(s, q1);
    if (LEAN || dim > 2) s = __dadd_rn(s, q2);
    return expTabT<!LEAN>(__dmul_rn(-f.p[0], s));
  }
  __device__ __forceinline__ double comp(int c, double cm, double x0, double x1, double x2) const {
    switch (id()) {
      case DFX_FN_CONSTANT: return f.p[c];
      case DFX_FN_IDENTITY: return c >= dim ? 0.0 : (c == 0 ? x0 : (c == 1 ? x1 : x2));
      case DFX_FN_GAUSSIAN_VEC: return __dmul_rn(c >= dim ? 1.0 : (c == 0 ? x0 : (c == 1 ? x1 : x2)), cm);
    }
    return cm;
  }
};
typedef RegistryFnT<-1, false> RegistryFn;

}  // namespace dfx
