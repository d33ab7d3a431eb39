// dfx_eval_fast.h — specialised evaluation kernels (sum factorisation, TMA bulk
// stores, FP64 tensor-core contraction) and the planner that picks one.
#pragma once
#include <cuda_runtime.h>

#include "dfx_internal.h"

namespace dfx {

constexpr int kMaxQ1 = 8;     // quadrature points per direction handled by the fast paths
constexpr int kMaxN1 = 8;     // (order+1) handled by the fast paths

// A tensor-product point set: xhat[q] = (x1[0][q0], x1[1][q1], x1[2][q2]) with
// q = q0 + m0*(q1 + m1*q2) when firstFastest, q = q2 + m2*(q1 + m1*q0) otherwise.
struct FastPlan {
  int path = 0;               // 0 generic, 1 register sum-factorisation, 2 shared-memory sum-factorisation, 3 DMMA
  int dim = 0, k = 0, m = 0;  // order and points per direction (same in every direction)
  bool firstFastest = false;
  bool jac = false;
  int nq = 0;
  // 1-d tables A[q][a] = l_a(x_q), Dm[q][a] = l_a'(x_q) per direction
  double A[3][kMaxQ1][kMaxN1];
  double Dm[3][kMaxQ1][kMaxN1];
  // collocation derivative on the points, Cm[q][r] = l_r'(x_q) with l_r the Lagrange polynomials
  // through the m points of the direction; valid (colloc) when the points are distinct
  double Cm[3][kMaxQ1][kMaxQ1];
  bool colloc = false;
  std::vector<double> xhat;   // the points (dense DMMA path builds its table from them)
};

// returns the path that will run (0 = generic) and fills `plan` when > 0
int planFastEval(const dfx_basis* b, int firstLeaf, int nLeaves, const double* xhat, int nq, bool wantJac, int forced,
                 FastPlan& plan);
bool smemPathHas(int dim, int k, int m);
extern int g_smemGrad;
extern int g_smemTma;
extern int g_smemPair;
extern int g_smemMaxBlocks;
extern int g_pairL2Ahead;
extern int g_evalQ3;
extern int g_evalSplitByGroup;
extern int g_evalSwizzle;
extern int g_evalStreamHint;
extern int g_evalBlock;
bool dmmaPathHas(const dfx_basis* b, int firstLeaf, int nLeaves);
int runDmmaEval(dfx_basis* b, int node, int firstLeaf, int nLeaves, const double* xhat, int nq, const double* coeff,
                u64 e0, u64 e1, double* values, double* jac, cudaStream_t st);
// compBase / compTotal: position of the group's first range entry in the output and the total
// number of range entries per point (a mixed-order subtree is evaluated run by run)
int runSmemEval(dfx_basis* b, const FastPlan& plan, int firstLeaf, int nLeaves, const double* coeff, u64 e0, u64 e1,
                double* values, double* jac, cudaStream_t st, int compBase = 0, int compTotal = -1);
// LagrangeDG leaf, whole element pairs, both elements of a pair contracted at once; -1 = not a configuration of it
int runPairEval(dfx_basis* b, const FastPlan& plan, int leafId, const double* coeff, u64 e0, u64 e1, double* values,
                double* jac, cudaStream_t st);
// continuous Lagrange order 3 at 4^3 points, values only, one leaf; -1 = not a configuration of it
int runQ3Eval(dfx_basis* b, const FastPlan& plan, int leafId, const double* coeff, u64 e0, u64 e1, double* values,
              cudaStream_t st);
// orientation codes of the local box's edges / facets (handles with permuted face DOFs)
int ensureOrientTable(dfx_basis* b, cudaStream_t st);
int runFastEval(dfx_basis* b, const FastPlan& plan, int firstLeaf, int nLeaves, const double* coeff, u64 e0, u64 e1,
                double* values, double* jac, cudaStream_t st, int compBase = 0, int compTotal = -1);

// one thread per (element, leaf): a run of >= 2 leaves on one lattice, optionally followed by an order-1 run
// behind an order-2 run (Taylor-Hood) in the same launch; -1 = not a configuration of this kernel
int runSplitEval(dfx_basis* b, const FastPlan& planA, int firstA, int nA, const FastPlan* planB, int firstB, int nB,
                 const double* coeff, u64 e0, u64 e1, double* values, double* jac, cudaStream_t st, int compBase = 0,
                 int compTotal = -1);

}  // namespace dfx
