// Test infrastructure, not product code.
//
// Per-stage entry points over the UNMODIFIED reference classes (compiled from
// /root/reference/src by oracle/Makefile into oracle/_ref/libpypde_stages.so),
// so that each GPU kernel can be checked against the reference function it
// replaces with identical inputs: boundaries(), WenoSolver::reconstruction,
// TimeStepper::step, DGSolver::predictor, FVSolver::apply, max_abs_eigs and
// the basis tables.  Only tests/ and bench.py's cpu_baseline leg load this.
#include "eigs/system.h"
#include "grid/boundaries.h"
#include "poly/basis.h"
#include "solvers/dg/dg.h"
#include "solvers/dg/dg_matrices.h"
#include "solvers/fv/fv.h"
#include "solvers/stepper.h"
#include "solvers/weno/weno.h"
#include "solvers/weno/weno_matrices.h"
#include "types.h"

#include <cmath>
#include <cstring>

typedef void (*Ffn)(double *, double *, double *, int);
typedef void (*Bfn)(double *, double *, int);
typedef void (*Sfn)(double *, double *);

static void put(double *dst, const Mat &m) {
  for (int i = 0; i < m.rows(); i++)
    for (int j = 0; j < m.cols(); j++)
      dst[i * m.cols() + j] = m(i, j);
}

extern "C" {

// basis.cpp / weno_matrices.cpp / dg_matrices.cpp tables for order N
void ref_tables(int N, double *nodes, double *wghts, double *derv, double *endv, double *dgmat,
                double *sig, double *mL, double *mR, double *mCL, double *mCR) {
  std::vector<poly> basis = basis_polys(N);
  Vec NODES = scaled_nodes(N);
  Vec WGHTS = scaled_weights(N);
  for (int i = 0; i < N; i++) {
    nodes[i] = NODES(i);
    wghts[i] = WGHTS(i);
  }
  put(derv, derivative_values(basis, NODES));
  put(endv, end_values(basis));
  Mat DG_END = end_value_products(basis);
  Mat DG_DER = derivative_products(basis, NODES, WGHTS);
  Mat DG_MAT = DG_END - DG_DER.transpose();
  put(dgmat, DG_MAT);
  put(sig, oscillation_indicator(basis));
  int FN2 = (int)floor((N - 1) / 2.);
  int CN2 = (int)ceil((N - 1) / 2.);
  std::vector<Mat> cm = coefficient_matrices(basis, FN2, CN2);
  put(mL, cm[0]);
  put(mR, cm[1]);
  put(mCL, cm[2]);
  put(mCR, cm[3]);
}

// grid/boundaries.cpp:25-53
void ref_boundaries(double *out, double *_u, int *_nX, int ndim, int V, int *_bt, int N) {
  iVecMap nX(_nX, ndim);
  iVecMap bt(_bt, ndim);
  int ncell = nX.prod();
  MatMap u(_u, ncell, V, OuterStride(V));
  Mat ub = boundaries(u, nX, bt, N);
  std::memcpy(out, ub.data(), sizeof(double) * ub.size());
}

// stepper.cpp:35-76 on a w array of (prod(nXw), N^ndim, V)
double ref_step(Ffn F, Bfn B, double *_w, long nrows, double *_dX, int ndim, int N, int V,
                double CFL, double tf, int secondOrder, double t, int count) {
  VecMap dX(_dX, ndim);
  TimeStepper ts(F, B, dX, N, V, CFL, tf, secondOrder != 0);
  MatMap w(_w, nrows, V, OuterStride(V));
  return ts.step(w, t, count);
}

// dg.cpp:204-224
void ref_predictor(double *out, Ffn F, Bfn B, Sfn S, double *_w, long nrows, double *_dX,
                   int ndim, int STIFF, int N, int V, double dt) {
  VecMap dX(_dX, ndim);
  DGSolver dg(F, B, S, dX, STIFF != 0, N, V);
  MatMap w(_w, nrows, V, OuterStride(V));
  Mat qh = dg.predictor(w, dt);
  std::memcpy(out, qh.data(), sizeof(double) * qh.size());
}

// fv.cpp:200-207 (u updated in place)
void ref_fv(double *_u, Ffn F, Bfn B, Sfn S, double *_qh, long qhrows, int *_nX, double *_dX,
            int ndim, int FLUX, int N, int V, int secondOrder, double dt) {
  iVecMap nX(_nX, ndim);
  VecMap dX(_dX, ndim);
  FVSolver fv(F, B, S, nX, dX, FLUX, N, V, secondOrder != 0);
  MatMap u(_u, nX.prod(), V, OuterStride(V));
  MatMap qh(_qh, qhrows, V, OuterStride(V));
  fv.apply(u, qh, dt);
}

// eigs/system.cpp:45-51
double ref_max_abs_eigs(Ffn F, Bfn B, double *_q, double *_dq, int d, int V, int ndim) {
  VecMap q(_q, V);
  MatMap dq(_dq, ndim, V, OuterStride(V));
  return max_abs_eigs(F, B, q, dq, d);
}

// spectral radius of a given matrix through the same Eigen path (system.cpp:28-43 for V<6)
double ref_spectral_radius(double *_M, int V) {
  MatMap M(_M, V, V, OuterStride(V));
  Mat A = M;
  return A.eigenvalues().array().abs().maxCoeff();
}
}
