/* PDE systems as plain C, one text for both sides of the parity tests:
 *   - NVRTC compiles it to device functions user_F / user_B / user_S that are
 *     LTO-linked into the kernels (pypde_b200.systems.cuda_source);
 *   - gcc compiles the same text (oracle/systems.c, -ffp-contract=off) into CPU
 *     callbacks with the reference's cfunc signatures (cfuncs.py:6-8) for the
 *     reference library.
 * Only + - * / sqrt and exp are used, in a fixed order, so the two sides agree
 * bit for bit except where exp() is involved.  The fluxes divide once (1/rho) and
 * multiply: on the GPU a double division whose quotient is 0 (momentum / rho in
 * gas at rest) takes div.rn.f64's slow path, ~10x the cost of the fast one.
 *
 * No include guard on purpose: the file is included once per system with
 *   SYS_<NAME>   selecting the system,
 *   SYS_NDIM     the number of space dimensions,
 *   SYS_F/SYS_B/SYS_S  the names to give the functions.
 * The systems mirror the reference's example problems (pypde/tests/...):
 *   EULER           tests/euler/system.py:10-21 generalised to ndim velocities
 *   REACTIVE_EULER  tests/reactive_euler/system.py:13-61, Arrhenius kinetics as
 *                   docs/pages/example_pdes.rst:35-61
 *   NAVIER_STOKES   tests/navier_stokes/system.py:22-52
 *   ADVECT_NC       a small non-conservative + source system for the B/S paths
 *   BURGERS         scalar conservation law (V = 1): the 1 x 1 eigen path
 *   GPR             tests/gpr/system.py + misc/*.py + params.py (Godunov-Peshkov-Romenski
 *                   continuum model, V = 17, stiffened-gas EOS with pINF = 0)
 */
#ifndef PDE_FN
#ifdef __CUDACC__
#define PDE_FN extern "C" __device__
#else
#define PDE_FN
#include <math.h>
#endif
#endif

#if defined(SYS_EULER)
/* Q = [rho, rho E, rho v_0 .. rho v_{ndim-1}],  V = 2 + ndim,  gamma = 1.4 */
PDE_FN void SYS_F(double *out, const double *Q, const double *dQ, int d) {
  const double g = 1.4;
  double r = Q[0];
  double ir = 1. / r;
  double E = Q[1] * ir;
  double v[SYS_NDIM];
  double vv = 0.;
  for (int i = 0; i < SYS_NDIM; i++) {
    v[i] = Q[2 + i] * ir;
    vv += v[i] * v[i];
  }
  double e = E - vv / 2.;
  double p = (g - 1.) * r * e;
  double vd = v[d];
  out[0] = r * vd;
  out[1] = r * E * vd + p * vd;
  for (int i = 0; i < SYS_NDIM; i++)
    out[2 + i] = r * v[i] * vd;
  out[2 + d] += p;
  (void)dQ;
}
#endif

#if defined(SYS_REACTIVE_EULER)
/* Q = [rho, rho E, rho v_0.., rho lambda],  V = 3 + ndim */
#ifndef SYS_RE_K0
#define SYS_RE_K0 250.
#endif
#ifndef SYS_RE_EA
#define SYS_RE_EA 2.
#endif
PDE_FN void SYS_F(double *out, const double *Q, const double *dQ, int d) {
  const double g = 1.4, Qc = 1.;
  double r = Q[0];
  double ir = 1. / r;
  double E = Q[1] * ir;
  double v[SYS_NDIM];
  double vv = 0.;
  for (int i = 0; i < SYS_NDIM; i++) {
    v[i] = Q[2 + i] * ir;
    vv += v[i] * v[i];
  }
  double lam = Q[2 + SYS_NDIM] * ir;
  double e = E - vv / 2. - Qc * (lam - 1.);
  double p = (g - 1.) * r * e;
  double vd = v[d];
  for (int i = 0; i < 3 + SYS_NDIM; i++)
    out[i] = vd * Q[i];
  out[1] += p * vd;
  out[2 + d] += p;
  (void)dQ;
}
PDE_FN void SYS_S(double *out, const double *Q) {
  const double Qc = 1., cv = 2.5;
  double r = Q[0];
  double ir = 1. / r;
  double E = Q[1] * ir;
  double vv = 0.;
  for (int i = 0; i < SYS_NDIM; i++) {
    double vi = Q[2 + i] * ir;
    vv += vi * vi;
  }
  double lam = Q[2 + SYS_NDIM] * ir;
  double e = E - vv / 2. - Qc * (lam - 1.);
  double T = e / cv;
  for (int i = 0; i < 3 + SYS_NDIM; i++)
    out[i] = 0.;
  out[2 + SYS_NDIM] = -r * lam * SYS_RE_K0 * exp(-SYS_RE_EA / T);
}
#endif

#if defined(SYS_NAVIER_STOKES)
/* Q = [rho, rho E, rho v_0, rho v_1, rho v_2], V = 5, second-order flux F(Q, dQ, d);
 * as the reference example, only the x-gradient row dQ[0] enters the stress. */
#ifndef SYS_NS_MU
#define SYS_NS_MU 1e-2
#endif
PDE_FN void SYS_F(double *out, const double *Q, const double *dQ, int d) {
  const double g = 1.4, mu = SYS_NS_MU;
  double r = Q[0];
  double ir = 1. / r;
  double E = Q[1] * ir;
  double v[3];
  for (int i = 0; i < 3; i++)
    v[i] = Q[2 + i] * ir;
  double dr_dx = dQ[0];
  double dv_dx[3];
  for (int i = 0; i < 3; i++)
    dv_dx[i] = (dQ[2 + i] - dr_dx * v[i]) * ir;
  /* dv[0][:] = dv_dx, other rows zero */
  double p = r * (g - 1.) * (E - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / 2.);
  double tr = dv_dx[0];
  /* sigma = mu (dv + dv^T - 2/3 tr I), row d */
  double sd[3];
  for (int j = 0; j < 3; j++) {
    double dv_dj = d == 0 ? dv_dx[j] : 0.;
    double dv_jd = j == 0 ? dv_dx[d] : 0.;
    double I = d == j ? 1. : 0.;
    sd[j] = mu * (dv_dj + dv_jd - 2. / 3. * tr * I);
  }
  double vd = v[d];
  double rvd = r * vd;
  out[0] = rvd;
  out[1] = rvd * E + p * vd;
  for (int i = 0; i < 3; i++)
    out[2 + i] = rvd * v[i];
  out[2 + d] += p;
  out[1] -= sd[0] * v[0] + sd[1] * v[1] + sd[2] * v[2];
  for (int i = 0; i < 3; i++)
    out[2 + i] -= sd[i];
}
#endif

#if defined(SYS_ADVECT_NC)
/* V = 3: q0 conserved with flux a_d q0 (1 + q0/4); q1, q2 advected
 * non-conservatively with a state-dependent matrix; linear relaxation source. */
PDE_FN void SYS_F(double *out, const double *Q, const double *dQ, int d) {
  double a = 1. - 0.35 * d;
  out[0] = a * Q[0] * (1. + Q[0] / 4.);
  out[1] = 0.;
  out[2] = 0.;
  (void)dQ;
}
PDE_FN void SYS_B(double *out, const double *Q, int d) {
  double a = 0.6 + 0.3 * d;
  for (int i = 0; i < 9; i++)
    out[i] = 0.;
  out[1 * 3 + 1] = a * (1. + 0.2 * Q[0]);
  out[1 * 3 + 2] = 0.1 * Q[1];
  out[2 * 3 + 0] = 0.05;
  out[2 * 3 + 2] = a + 0.1 * Q[2];
}
PDE_FN void SYS_S(double *out, const double *Q) {
  out[0] = -0.5 * (Q[0] - 1.);
  out[1] = 0.3 * Q[2] - 0.2 * Q[1];
  out[2] = -0.1 * Q[2] * Q[0];
}
#endif

#if defined(SYS_BURGERS)
/* V = 1: F_d = a_d q^2 / 2 */
PDE_FN void SYS_F(double *out, const double *Q, const double *dQ, int d) {
  double a = 1. - 0.4 * d;
  out[0] = a * Q[0] * Q[0] / 2.;
  (void)dQ;
}
#endif

#if defined(SYS_GPR)
/* Q = [rho, rho E, rho v(3), A(3x3 row-major), rho J(3)], V = 17.
 * Parameters of reference tests/gpr/params.py. */
#define GPR_G 1.4
#define GPR_CV 2.5
#define GPR_CS2 25.
#define GPR_CA2 25.
#define GPR_MU 2e-2
#define GPR_PR 0.75
#define GPR_RHO0 1.
#define GPR_P0 (1. / GPR_G)
#define GPR_KAPPA (GPR_MU * GPR_G * GPR_CV / GPR_PR)
#define GPR_T0 (GPR_P0 / (GPR_RHO0 * (GPR_G - 1.) * GPR_CV))
#define GPR_TAU1 (6. * GPR_MU / (GPR_RHO0 * GPR_CS2))
#define GPR_TAU2 (GPR_KAPPA * GPR_RHO0 / (GPR_T0 * GPR_CA2))

#ifndef GPR_HELPERS
#define GPR_HELPERS
#ifdef __CUDACC__
#define GPR_INL static __device__ inline
#else
#define GPR_INL static inline
#endif
/* psi = dE/dA = cs^2 A dev(A^T A); returns sum(dev(G)^2) for E_2A (misc/state.py, eos.py) */
GPR_INL double gpr_psi(const double *A, double *psi) {
  double G[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0.;
      for (int k = 0; k < 3; k++)
        acc += A[k * 3 + i] * A[k * 3 + j];
      G[i * 3 + j] = acc;
    }
  double tr3 = (G[0] + G[4] + G[8]) / 3.;
  G[0] -= tr3;
  G[4] -= tr3;
  G[8] -= tr3;
  double s2 = 0.;
  for (int i = 0; i < 9; i++)
    s2 += G[i] * G[i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0.;
      for (int k = 0; k < 3; k++)
        acc += A[i * 3 + k] * G[k * 3 + j];
      psi[i * 3 + j] = GPR_CS2 * acc;
    }
  return s2;
}
/* pressure and temperature of the state (misc/state.py, mg.py) */
GPR_INL void gpr_pT(double r, double E, const double *v, double devG2, const double *J, double *p,
                    double *T) {
  double E3 = (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / 2.;
  double E2A = GPR_CS2 / 4. * devG2;
  double E2J = GPR_CA2 / 2. * (J[0] * J[0] + J[1] * J[1] + J[2] * J[2]);
  double E1 = E - E3 - E2A - E2J;
  *p = E1 * r * (GPR_G - 1.);
  *T = *p / (r * (GPR_G - 1.) * GPR_CV);
}
#endif

PDE_FN void SYS_F(double *out, const double *Q, const double *dQ, int d) {
  double r = Q[0];
  double ir = 1. / r;
  double E = Q[1] * ir;
  double v[3], J[3], psi[9], sig[9];
  const double *A = Q + 5;
  for (int i = 0; i < 3; i++) {
    v[i] = Q[2 + i] * ir;
    J[i] = Q[14 + i] * ir;
  }
  double devG2 = gpr_psi(A, psi);
  double p, T;
  gpr_pT(r, E, v, devG2, J, &p, &T);
  /* sigma = -rho A^T psi */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0.;
      for (int k = 0; k < 3; k++)
        acc += A[k * 3 + i] * psi[k * 3 + j];
      sig[i * 3 + j] = -r * acc;
    }
  double vd = v[d];
  double rvd = r * vd;
  for (int i = 0; i < 17; i++)
    out[i] = 0.;
  out[0] = rvd;
  out[1] = rvd * E + p * vd;
  for (int i = 0; i < 3; i++)
    out[2 + i] = rvd * v[i];
  out[2 + d] += p;
  out[1] -= sig[d * 3 + 0] * v[0] + sig[d * 3 + 1] * v[1] + sig[d * 3 + 2] * v[2];
  for (int i = 0; i < 3; i++)
    out[2 + i] -= sig[d * 3 + i];
  for (int i = 0; i < 3; i++)
    out[5 + 3 * i + d] = A[i * 3 + 0] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
  out[1] += GPR_CA2 * J[d] * T;
  for (int i = 0; i < 3; i++)
    out[14 + i] = rvd * J[i];
  out[14 + d] += T;
  (void)dQ;
}
PDE_FN void SYS_B(double *out, const double *Q, int d) {
  double ir = 1. / Q[0];
  double v[3];
  for (int i = 0; i < 3; i++)
    v[i] = Q[2 + i] * ir;
  /* straight-line zero fill: every entry that stays zero is then a compile-time constant
     where this function is inlined, and the kernels, which skip zero entries of B, never
     hold B as an array in local memory */
#define GPR_Z(r)                                                                                 \
  out[(r) * 17 + 0] = 0.; out[(r) * 17 + 1] = 0.; out[(r) * 17 + 2] = 0.; out[(r) * 17 + 3] = 0.;  \
  out[(r) * 17 + 4] = 0.; out[(r) * 17 + 5] = 0.; out[(r) * 17 + 6] = 0.; out[(r) * 17 + 7] = 0.;  \
  out[(r) * 17 + 8] = 0.; out[(r) * 17 + 9] = 0.; out[(r) * 17 + 10] = 0.;                        \
  out[(r) * 17 + 11] = 0.; out[(r) * 17 + 12] = 0.; out[(r) * 17 + 13] = 0.;                      \
  out[(r) * 17 + 14] = 0.; out[(r) * 17 + 15] = 0.; out[(r) * 17 + 16] = 0.;
  GPR_Z(0) GPR_Z(1) GPR_Z(2) GPR_Z(3) GPR_Z(4) GPR_Z(5) GPR_Z(6) GPR_Z(7) GPR_Z(8) GPR_Z(9)
  GPR_Z(10) GPR_Z(11) GPR_Z(12) GPR_Z(13) GPR_Z(14) GPR_Z(15) GPR_Z(16)
#undef GPR_Z
  const double vd = d == 0 ? v[0] : (d == 1 ? v[1] : v[2]);
  out[5 * 17 + 5] = vd; out[6 * 17 + 6] = vd; out[7 * 17 + 7] = vd;
  out[8 * 17 + 8] = vd; out[9 * 17 + 9] = vd; out[10 * 17 + 10] = vd;
  out[11 * 17 + 11] = vd; out[12 * 17 + 12] = vd; out[13 * 17 + 13] = vd;
  /* ret[5+d, 5+d:8+d] -= v etc., exactly as the reference example's slices; one
     statically indexed block per direction */
#define GPR_BD(D)                                                                                \
  out[(5 + D) * 17 + 5 + D + 0] -= v[0]; out[(5 + D) * 17 + 5 + D + 1] -= v[1];                  \
  out[(5 + D) * 17 + 5 + D + 2] -= v[2];                                                          \
  out[(8 + D) * 17 + 8 + D + 0] -= v[0]; out[(8 + D) * 17 + 8 + D + 1] -= v[1];                  \
  out[(8 + D) * 17 + 8 + D + 2] -= v[2];                                                          \
  out[(11 + D) * 17 + 11 + D + 0] -= v[0]; out[(11 + D) * 17 + 11 + D + 1] -= v[1];              \
  out[(11 + D) * 17 + 11 + D + 2] -= v[2];
  if (d == 0) {
    GPR_BD(0)
  } else if (d == 1) {
    GPR_BD(1)
  } else {
    GPR_BD(2)
  }
#undef GPR_BD
}
PDE_FN void SYS_S(double *out, const double *Q) {
  double r = Q[0];
  double ir = 1. / r;
  double E = Q[1] * ir;
  double v[3], J[3], psi[9];
  const double *A = Q + 5;
  for (int i = 0; i < 3; i++) {
    v[i] = Q[2 + i] * ir;
    J[i] = Q[14 + i] * ir;
  }
  double devG2 = gpr_psi(A, psi);
  double p, T;
  gpr_pT(r, E, v, devG2, J, &p, &T);
  double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
               A[2] * (A[3] * A[7] - A[4] * A[6]);
  double th1 = 3. * pow(det, 5. / 3.) / (GPR_CS2 * GPR_TAU1);
  double th2 = 1. / (GPR_CA2 * GPR_TAU2 * (r / GPR_RHO0) * (GPR_T0 / T));
  for (int i = 0; i < 5; i++)
    out[i] = 0.;
  for (int i = 0; i < 9; i++)
    out[5 + i] = -psi[i] * th1;
  for (int i = 0; i < 3; i++)
    out[14 + i] = -r * (GPR_CA2 * J[i]) * th2;
}
#endif
