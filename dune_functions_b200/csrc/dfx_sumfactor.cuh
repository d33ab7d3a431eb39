// dfx_sumfactor.cuh — register-resident sum factorisation shared by the evaluation kernel
// (dfx_eval_fast.cu) and the matrix-free operator (dfx_operators.cu).
#pragma once

namespace dfx {

// contraction of one direction: out[q] = sum_a T[q][a] * in[a]
template <int N1, int M>
__device__ __forceinline__ void contract1(const double (&T)[M][N1], const double (&in)[N1], double (&out)[M]) {
#pragma unroll
  for (int q = 0; q < M; ++q) {
    double s = T[q][0] * in[0];
#pragma unroll
    for (int a = 1; a < N1; ++a) s = fma(T[q][a], in[a], s);
    out[q] = s;
  }
}

// u[q0][q1][q2] = sum_a T0[q0][a0] T1[q1][a1] T2[q2][a2] c[a0][a1][a2]   (DIM = 3; DIM = 2 drops a2)
template <int DIM, int N1, int M>
__device__ __forceinline__ void sumFactor(const double (&T0)[M][N1], const double (&T1)[M][N1],
                                          const double (&T2)[M][N1], const double* __restrict__ c,
                                          double* __restrict__ u) {
  constexpr int NZ = DIM == 3 ? N1 : 1;
  constexpr int MZ = DIM == 3 ? M : 1;
  // x: t1[q0][a1][a2]
  double t1[M][N1][NZ];
#pragma unroll
  for (int a2 = 0; a2 < NZ; ++a2)
#pragma unroll
    for (int a1 = 0; a1 < N1; ++a1) {
      double in[N1], out[M];
#pragma unroll
      for (int a0 = 0; a0 < N1; ++a0) in[a0] = c[a0 + N1 * (a1 + N1 * a2)];
      contract1<N1, M>(T0, in, out);
#pragma unroll
      for (int q = 0; q < M; ++q) t1[q][a1][a2] = out[q];
    }
  // y: t2[q0][q1][a2]
  double t2[M][M][NZ];
#pragma unroll
  for (int q0 = 0; q0 < M; ++q0)
#pragma unroll
    for (int a2 = 0; a2 < NZ; ++a2) {
      double in[N1], out[M];
#pragma unroll
      for (int a1 = 0; a1 < N1; ++a1) in[a1] = t1[q0][a1][a2];
      contract1<N1, M>(T1, in, out);
#pragma unroll
      for (int q = 0; q < M; ++q) t2[q0][q][a2] = out[q];
    }
  if (DIM == 2) {
#pragma unroll
    for (int q0 = 0; q0 < M; ++q0)
#pragma unroll
      for (int q1 = 0; q1 < M; ++q1) u[q0 + M * q1] = t2[q0][q1][0];
    return;
  }
  // z
#pragma unroll
  for (int q0 = 0; q0 < M; ++q0)
#pragma unroll
    for (int q1 = 0; q1 < M; ++q1) {
      double in[N1], out[M];
#pragma unroll
      for (int a2 = 0; a2 < NZ; ++a2) in[a2] = t2[q0][q1][a2];
      contract1<N1, M>(T2, in, out);
#pragma unroll
      for (int q = 0; q < MZ; ++q) u[q0 + M * (q1 + M * q)] = out[q];
    }
}

}  // namespace dfx
