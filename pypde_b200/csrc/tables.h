// Host-side constant tables of the ADER-WENO scheme (runs once per solve).
//
// Replaces, for the GPU path, the per-solver-object table construction of the
// reference: poly/basis.cpp:7-76 (nodes, weights, Lagrange basis, ENDVALS,
// DERVALS), solvers/weno/weno_matrices.cpp:8-51 (stencil matrices, oscillation
// indicator) and solvers/dg/dg_matrices.cpp:27-55 + dg.cpp:36-39 (DG_MAT).
// Everything is evaluated in long double from closed forms and rounded once to
// double; the constant N x N systems are inverted here so the kernels only do
// small mat-vecs.
#pragma once
#include <string>
#include <vector>

namespace pypde {

struct BasisTables {
  int N = 0;
  std::vector<double> nodes;    // [N]      Gauss-Legendre nodes on [0,1]
  std::vector<double> wghts;    // [N]      weights (sum to 1)
  std::vector<double> derv;     // [N][N]   derv[i][j]  = psi_j'(x_i)
  std::vector<double> endv;     // [2][N]   endv[e][j]  = psi_j(e)
  std::vector<double> dgmat;    // [N][N]   DG_END - DG_DER^T  (dg.cpp:39)
  std::vector<double> dginv;    // [N][N]   inverse of dgmat
  std::vector<double> wm[4];    // [N][N]   mL, mR, mCL, mCR  (weno_matrices.cpp)
  std::vector<double> wminv[4]; // [N][N]   their inverses
  std::vector<double> sig;      // [N][N]   oscillation indicator
  int nstencils = 0;            // 2 (N==2), 3 (N odd > 2), 4 (N even > 2)
  int stencil_off[4] = {0, 0, 0, 0}; // first row of each stencil window
  double stencil_lam[4] = {0, 0, 0, 0};
};

BasisTables make_tables(int N);

// Emits the tables as CUDA source (`__constant__` arrays with hex-float
// initialisers) that is prepended to the kernel translation unit.
std::string tables_cuda_source(const BasisTables &T);

} // namespace pypde
