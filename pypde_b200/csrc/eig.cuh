// pypde_b200 device eigen-solver (D1), shared text:
//   * prepended to kernels.cuh in the JIT translation unit (device code), and
//   * compiled by g++ into libpypde.so as pypde_b200_host_spectral_radius for the
//     CPU unit tests of exactly this code (tests/test_eig.py).
// Replaces Eigen's EigenSolver / Spectra in the reference's eigs/system.cpp:28-43.
#ifndef PYPDE_B200_EIG_CUH
#define PYPDE_B200_EIG_CUH

#ifdef __CUDACC__
#define EIG_FN __device__ __forceinline__
#define EIG_FN_NOINLINE __device__ __noinline__
#else
#include <cmath>
#define EIG_FN inline
#define EIG_FN_NOINLINE inline
using std::cbrt;
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
using std::hypot;
using std::pow;
using std::sqrt;
#endif

#define DBL_EPS 2.2204460492503131e-16

// max of two non-NaN values as one compare + select (fmax's NaN semantics cost extra
// instructions in the hot path)
EIG_FN double sel_max(double a, double b) { return a > b ? a : b; }
#ifndef PDE_EIG_QR_ONLY
#define PDE_EIG_QR_ONLY 0 // 1: always use the general QR iteration for spectral radii
#endif

// ---------------------------------------------------------------------------
// D1: spectral radius of a general real V x V matrix (replaces Eigen
// EigenSolver / Spectra at eigs/system.cpp:28-43): elimination to Hessenberg
// form followed by the Francis double-shift QR iteration, eigenvalues only.
// a is row-major n x n in thread-local memory and is destroyed.
// ---------------------------------------------------------------------------
template <int n> EIG_FN_NOINLINE double spectral_radius_qr(double *a) {
#define A_(i, j) a[(i) * n + (j)]
  if (n == 1)
    return fabs(A_(0, 0));
  // --- Hessenberg reduction by stabilised elementary transformations
  for (int m = 1; m < n - 1; m++) {
    double x = 0.;
    int i = m;
    for (int j = m; j < n; j++)
      if (fabs(A_(j, m - 1)) > fabs(x)) {
        x = A_(j, m - 1);
        i = j;
      }
    if (i != m) {
      for (int j = m - 1; j < n; j++) {
        double tmp = A_(i, j);
        A_(i, j) = A_(m, j);
        A_(m, j) = tmp;
      }
      for (int j = 0; j < n; j++) {
        double tmp = A_(j, i);
        A_(j, i) = A_(j, m);
        A_(j, m) = tmp;
      }
    }
    if (x != 0.) {
      for (i = m + 1; i < n; i++) {
        double y = A_(i, m - 1);
        if (y != 0.) {
          y /= x;
          A_(i, m - 1) = y;
          for (int j = m; j < n; j++)
            A_(i, j) -= y * A_(m, j);
          for (int j = 0; j < n; j++)
            A_(j, m) += y * A_(j, i);
        }
      }
    }
  }
  for (int i = 2; i < n; i++)
    for (int j = 0; j < i - 1; j++)
      A_(i, j) = 0.;

  // --- QR iteration
  double rad = 0.;
  double anorm = 0.;
  for (int i = 0; i < n; i++)
    for (int j = (i > 0 ? i - 1 : 0); j < n; j++)
      anorm += fabs(A_(i, j));
  int nn = n - 1;
  double t = 0.;
  double p = 0., q = 0., r = 0., s, w, x, y, z;
  while (nn >= 0) {
    int its = 0;
    int l;
    do {
      for (l = nn; l > 0; l--) {
        s = fabs(A_(l - 1, l - 1)) + fabs(A_(l, l));
        // both diagonal entries negligible against the matrix (zero, or denormal
        // leftovers such as a fully burnt mass fraction ~1e-308): measure the
        // subdiagonal against the norm instead — a backward error of eps ||A||
        if (s <= DBL_EPS * anorm)
          s = anorm;
        if (fabs(A_(l, l - 1)) <= DBL_EPS * s) {
          A_(l, l - 1) = 0.;
          break;
        }
      }
      x = A_(nn, nn);
      if (l == nn) { // one real root
        rad = fmax(rad, fabs(x + t));
        nn--;
      } else {
        y = A_(nn - 1, nn - 1);
        w = A_(nn, nn - 1) * A_(nn - 1, nn);
        if (l == nn - 1) { // two roots
          p = 0.5 * (y - x);
          q = p * p + w;
          z = sqrt(fabs(q));
          x += t;
          if (q >= 0.) { // real pair
            z = p + (p >= 0. ? fabs(z) : -fabs(z));
            double r1 = x + z;
            double r2 = r1;
            if (z != 0.)
              r2 = x - w / z;
            rad = fmax(rad, fmax(fabs(r1), fabs(r2)));
          } else { // complex pair
            rad = fmax(rad, hypot(x + p, z));
          }
          nn -= 2;
        } else { // no root yet: QR step
          if (its >= 60) { // no convergence: fall back to a norm bound
            return anorm;
          }
          if (its == 10 || its == 20) { // exceptional shift
            t += x;
            for (int i = 0; i <= nn; i++)
              A_(i, i) -= x;
            s = fabs(A_(nn, nn - 1)) + fabs(A_(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          for (m = nn - 2; m >= l; m--) {
            z = A_(m, m);
            r = x - z;
            s = y - z;
            p = (r * s - w) / A_(m + 1, m) + A_(m, m + 1);
            q = A_(m + 1, m + 1) - z - r - s;
            r = A_(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            {
              const double rs_ = 1. / s; // (one reciprocal for the three quotients: the slow
              p *= rs_;                  //  path of div.rn.f64 — zero quotients, all over a
              q *= rs_;                  //  sparse matrix — was 5-11 % of the QR kernels)
              r *= rs_;
            }
            if (m == l)
              break;
            double uu = fabs(A_(m, m - 1)) * (fabs(q) + fabs(r));
            double vv = fabs(p) * (fabs(A_(m - 1, m - 1)) + fabs(z) +
                                   fabs(A_(m + 1, m + 1)));
            if (uu <= DBL_EPS * vv)
              break;
          }
          for (int i = m; i < nn - 1; i++) {
            A_(i + 2, i) = 0.;
            if (i != m)
              A_(i + 2, i - 1) = 0.;
          }
          for (int k = m; k < nn; k++) {
            if (k != m) {
              p = A_(k, k - 1);
              q = A_(k + 1, k - 1);
              r = 0.;
              if (k + 1 != nn)
                r = A_(k + 2, k - 1);
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.) {
                const double rx_ = 1. / x;
                p *= rx_;
                q *= rx_;
                r *= rx_;
              }
            }
            double sq = sqrt(p * p + q * q + r * r);
            s = p >= 0. ? sq : -sq;
            if (s != 0.) {
              if (k == m) {
                if (l != m)
                  A_(k, k - 1) = -A_(k, k - 1);
              } else
                A_(k, k - 1) = -s * x;
              p += s;
              {
                const double rs_ = 1. / s, rp_ = 1. / p;
                x = p * rs_;
                y = q * rs_;
                z = r * rs_;
                q *= rp_;
                r *= rp_;
              }
              for (int j = k; j <= nn; j++) {
                p = A_(k, j) + q * A_(k + 1, j);
                if (k + 1 != nn) {
                  p += r * A_(k + 2, j);
                  A_(k + 2, j) -= p * z;
                }
                A_(k + 1, j) -= p * y;
                A_(k, j) -= p * x;
              }
              int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; i++) {
                p = x * A_(i, k) + y * A_(i, k + 1);
                if (k + 1 != nn) {
                  p += z * A_(i, k + 2);
                  A_(i, k + 2) -= p * r;
                }
                A_(i, k + 1) -= p * q;
                A_(i, k) -= p;
              }
            }
          }
        }
      }
    } while (l + 1 < nn);
  }
  return rad;
#undef A_
}

// The same iteration on the leading n x n block of a matrix with row pitch ld, n a run-time
// value: the active block that remains after the permutation step below.
// Part 1: Hessenberg reduction by stabilised elementary transformations (the multipliers stay
// below the subdiagonal; hqr_rt clears them).
#ifndef PDE_EIG_HESS_SKIP
#define PDE_EIG_HESS_SKIP 0 // (measured slower on B200: the branches break the pipelining of the loads)
#endif
EIG_FN void hessenberg_rt(double *a, const int ld, const int n) {
#define A_(i, j) a[(i) * ld + (j)]
  for (int m = 1; m < n - 1; m++) {
    double x = 0.;
    int i = m;
    for (int j = m; j < n; j++)
      if (fabs(A_(j, m - 1)) > fabs(x)) {
        x = A_(j, m - 1);
        i = j;
      }
    if (i != m) {
      for (int j = m - 1; j < n; j++) {
        double tmp = A_(i, j);
        A_(i, j) = A_(m, j);
        A_(m, j) = tmp;
      }
      for (int j = 0; j < n; j++) {
        double tmp = A_(j, i);
        A_(j, i) = A_(j, m);
        A_(j, m) = tmp;
      }
    }
    if (x != 0.) {
      // (any multiplier gives an exact similarity as long as the row and the column operation
      //  use the same one: a reciprocal per column instead of a quotient per row)
      const double rx = 1. / x;
      for (i = m + 1; i < n; i++) {
        double y = A_(i, m - 1);
        if (y != 0.) {
          y *= rx;
          A_(i, m - 1) = y;
#if PDE_EIG_HESS_SKIP
          // (sparse matrices: a zero factor leaves the target as it is — x - y * 0 = x for
          //  finite values — and saves its load and store; the kernels that run this are
          //  bound by the latency of exactly these local-memory accesses)
          for (int j = m; j < n; j++) {
            const double v = A_(m, j);
            if (v != 0.)
              A_(i, j) -= y * v;
          }
          for (int j = 0; j < n; j++) {
            const double v = A_(j, i);
            if (v != 0.)
              A_(j, m) += y * v;
          }
#else
          for (int j = m; j < n; j++)
            A_(i, j) -= y * A_(m, j);
          for (int j = 0; j < n; j++)
            A_(j, m) += y * A_(j, i);
#endif
        }
      }
    }
  }
#undef A_
}

// Part 2: the double-shift QR iteration on the Hessenberg matrix hessenberg_rt leaves.
EIG_FN_NOINLINE double hqr_rt(double *a, const int ld, const int n) {
#define A_(i, j) a[(i) * ld + (j)]
  for (int i = 2; i < n; i++)
    for (int j = 0; j < i - 1; j++)
      A_(i, j) = 0.;

  double rad = 0.;
  double anorm = 0.;
  for (int i = 0; i < n; i++)
    for (int j = (i > 0 ? i - 1 : 0); j < n; j++)
      anorm += fabs(A_(i, j));
  int nn = n - 1;
  double t = 0.;
  double p = 0., q = 0., r = 0., s, w, x, y, z;
  while (nn >= 0) {
    int its = 0;
    int l;
    do {
      for (l = nn; l > 0; l--) {
        s = fabs(A_(l - 1, l - 1)) + fabs(A_(l, l));
        // both diagonal entries negligible against the matrix (zero, or denormal
        // leftovers such as a fully burnt mass fraction ~1e-308): measure the
        // subdiagonal against the norm instead — a backward error of eps ||A||
        if (s <= DBL_EPS * anorm)
          s = anorm;
        if (fabs(A_(l, l - 1)) <= DBL_EPS * s) {
          A_(l, l - 1) = 0.;
          break;
        }
      }
      x = A_(nn, nn);
      if (l == nn) { // one real root
        rad = fmax(rad, fabs(x + t));
        nn--;
      } else {
        y = A_(nn - 1, nn - 1);
        w = A_(nn, nn - 1) * A_(nn - 1, nn);
        if (l == nn - 1) { // two roots
          p = 0.5 * (y - x);
          q = p * p + w;
          z = sqrt(fabs(q));
          x += t;
          if (q >= 0.) { // real pair
            z = p + (p >= 0. ? fabs(z) : -fabs(z));
            double r1 = x + z;
            double r2 = r1;
            if (z != 0.)
              r2 = x - w / z;
            rad = fmax(rad, fmax(fabs(r1), fabs(r2)));
          } else { // complex pair
            rad = fmax(rad, hypot(x + p, z));
          }
          nn -= 2;
        } else { // no root yet: QR step
          if (its >= 60) { // no convergence: fall back to a norm bound
            return anorm;
          }
          if (its == 10 || its == 20) { // exceptional shift
            t += x;
            for (int i = 0; i <= nn; i++)
              A_(i, i) -= x;
            s = fabs(A_(nn, nn - 1)) + fabs(A_(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          for (m = nn - 2; m >= l; m--) {
            z = A_(m, m);
            r = x - z;
            s = y - z;
            p = (r * s - w) / A_(m + 1, m) + A_(m, m + 1);
            q = A_(m + 1, m + 1) - z - r - s;
            r = A_(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            {
              const double rs_ = 1. / s; // (one reciprocal for the three quotients: the slow
              p *= rs_;                  //  path of div.rn.f64 — zero quotients, all over a
              q *= rs_;                  //  sparse matrix — was 5-11 % of the QR kernels)
              r *= rs_;
            }
            if (m == l)
              break;
            double uu = fabs(A_(m, m - 1)) * (fabs(q) + fabs(r));
            double vv = fabs(p) * (fabs(A_(m - 1, m - 1)) + fabs(z) +
                                   fabs(A_(m + 1, m + 1)));
            if (uu <= DBL_EPS * vv)
              break;
          }
          for (int i = m; i < nn - 1; i++) {
            A_(i + 2, i) = 0.;
            if (i != m)
              A_(i + 2, i - 1) = 0.;
          }
          for (int k = m; k < nn; k++) {
            if (k != m) {
              p = A_(k, k - 1);
              q = A_(k + 1, k - 1);
              r = 0.;
              if (k + 1 != nn)
                r = A_(k + 2, k - 1);
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.) {
                const double rx_ = 1. / x;
                p *= rx_;
                q *= rx_;
                r *= rx_;
              }
            }
            double sq = sqrt(p * p + q * q + r * r);
            s = p >= 0. ? sq : -sq;
            if (s != 0.) {
              if (k == m) {
                if (l != m)
                  A_(k, k - 1) = -A_(k, k - 1);
              } else
                A_(k, k - 1) = -s * x;
              p += s;
              {
                const double rs_ = 1. / s, rp_ = 1. / p;
                x = p * rs_;
                y = q * rs_;
                z = r * rs_;
                q *= rp_;
                r *= rp_;
              }
              for (int j = k; j <= nn; j++) {
                p = A_(k, j) + q * A_(k + 1, j);
                if (k + 1 != nn) {
                  p += r * A_(k + 2, j);
                  A_(k + 2, j) -= p * z;
                }
                A_(k + 1, j) -= p * y;
                A_(k, j) -= p * x;
              }
              int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; i++) {
                p = x * A_(i, k) + y * A_(i, k + 1);
                if (k + 1 != nn) {
                  p += z * A_(i, k + 2);
                  A_(i, k + 2) -= p * r;
                }
                A_(i, k + 1) -= p * q;
                A_(i, k) -= p;
              }
            }
          }
        }
      }
    } while (l + 1 < nn);
  }
  return rad;
#undef A_
}

EIG_FN double spectral_radius_qr_rt(double *a, const int ld, const int n) {
  if (n == 1)
    return fabs(a[0]);
  hessenberg_rt(a, ld, n);
  return hqr_rt(a, ld, n);
}

// ---------------------------------------------------------------------------
// D1 fast path for n <= 5 (the Euler / reactive Euler / Navier-Stokes sizes):
// all eigenvalues from the characteristic polynomial of the trace-shifted
// matrix B = A - (tr A / n) I, held entirely in registers:
//   * coefficients by the Faddeev-LeVerrier recurrence (n-1 small mat-mats),
//   * real roots peeled alternately from the right and from the left by
//     Laguerre's iteration started outside the spectrum (||B||_inf bounds every
//     root; from outside the real roots the iteration is monotone), deflating
//     until a quadratic remains, which is solved in closed form.
// The spectral radius is accepted only if it is attained by one of the two
// outermost real roots (each polished on the undeflated polynomial) and that
// root is well conditioned (kappa * eps < ~1e-13); in every other case —
// a dominant complex pair, a multiple outer root, slow or non-monotone
// convergence, NaNs — the general QR iteration above decides.  For hyperbolic
// systems the accepted case is the rule; the result agrees with the QR
// iteration to rounding (tests/test_gpu_parity.py::test_spectral_radius_*).
// ---------------------------------------------------------------------------
// Warm start for a sequence of nearby matrices (consecutive time nodes of one
// face trace, left/right state of a face): the outer roots of the previous
// shifted characteristic polynomial.
struct EigGuess {
  double yp, ym; // largest / smallest real root of the previous solve (shifted)
  double dy;     // change of the outer roots between the last two solves
  int valid;
};

template <int n> struct PolyCoef { // coefficients by value: for the out-of-line cold paths
  double v[n];
};
struct PolyIterResult {
  double root;
  int ok;
};
#ifndef PDE_EIG_ITER_NOINLINE
#define PDE_EIG_ITER_NOINLINE 0 // (1: measured neutral on B200 at C2)
#endif

template <int m> struct PolyRoots {
  // One monotone step towards the largest real root from a point to its right:
  // Laguerre's iteration in the one-division form
  //   a = m p / (p' + sqrt((m-1)((m-1) p'^2 - m p p'')))
  // (cubic), or Newton's a = p / p' once `newton` is set.  state: 0 iterating,
  // 1 converged (root written), -1 failed (left the monotone regime / NaN).
  static EIG_FN void step(const double *c, double &x, bool &newton, int &state, double &root) {
    double p = 1., dp = 0., d2 = 0.; // p, p', p''/2 by Horner
#pragma unroll
    for (int k = m - 1; k >= 0; k--) {
      d2 = fma(d2, x, dp);
      dp = fma(dp, x, p);
      p = fma(p, x, c[k]);
    }
    if (!(p > 0.) || !(dp > 0.)) {
      // on (or a rounding error past) the root, or outside the regime
      double ab = 1.;
      const double ax = fabs(x);
#pragma unroll
      for (int k = m - 1; k >= 0; k--)
        ab = fma(ab, ax, fabs(c[k]));
      if (fabs(p) <= 64. * DBL_EPS * ab && dp > 0.) {
        root = x - p / dp;
        state = 1;
      } else {
        state = -1;
      }
      return;
    }
    double a;
    if (newton) {
      a = p / dp;
    } else {
      const double disc = (m - 1) * ((m - 1) * dp * dp - 2. * m * p * d2);
      a = disc > 0. ? m * p / (dp + sqrt(disc)) : p / dp;
    }
    if (!(a >= 0.) || !(a <= 1e300)) { // NaN / inf
      state = -1;
      return;
    }
    x -= a;
    const double ax = fabs(x);
    if (a <= 1e-5 * ax) {
      // the error is now ~1e-10 |x| (quadratic) or smaller: one Newton step ends it
      p = 1.;
      dp = 0.;
#pragma unroll
      for (int k = m - 1; k >= 0; k--) {
        dp = fma(dp, x, p);
        p = fma(p, x, c[k]);
      }
      if (dp > 0.) {
        root = x - p / dp;
        state = 1;
      } else {
        state = -1;
      }
      return;
    }
    newton = a <= 0.02 * ax;
  }

  // The monotone iteration from x, right of the largest real root.  false = fall back to QR.
  static EIG_FN bool iterate(const double *c, double x, bool newton, double &root) {
    int state = 0;
    for (int it = 0; it < 30 && state == 0; it++)
      step(c, x, newton, state, root);
    return state == 1;
  }
  // The same out of line (arguments by value, so the caller's coefficients stay in
  // registers): searches that the warm Halley step did not finish.
  static EIG_FN_NOINLINE PolyIterResult iterate_call(PolyCoef<m> pc, double x, bool newton) {
    PolyIterResult r;
    r.root = 0.;
    r.ok = iterate(pc.v, x, newton, r.root) ? 1 : 0;
    return r;
  }
};

// complex pair of y^2 + b1 y + b0 (disc < 0): |mu + y|^2
EIG_FN double pair_modulus2(double mu, double b1, double b0) {
  const double re = mu - 0.5 * b1;
  return re * re + (b0 - 0.25 * b1 * b1);
}

// Phase 1: mean shift mu = tr A / n and the coefficients of the characteristic
// polynomial p(y) = y^n + c[n-1] y^(n-1) + .. + c[0] of B = A - mu I (c[n-1] = 0).
template <int n> EIG_FN void poly_setup(const double *A, double &mu, double *c) {
  mu = 0.;
#pragma unroll
  for (int i = 0; i < n; i++)
    mu += A[i * n + i];
  mu *= (1. / n);
  double B[n * n];
#pragma unroll
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int j = 0; j < n; j++)
      B[i * n + j] = A[i * n + j] - (i == j ? mu : 0.);
  c[n - 1] = 0.;
  if (n == 3) {
    // p(y) = y^3 + c1 y + c0: c1 = sum of principal 2 x 2 minors, c0 = -det B
    const double m01 = fma(B[0], B[4], -(B[1] * B[3]));
    const double m02 = fma(B[0], B[8], -(B[2] * B[6]));
    const double m12 = fma(B[4], B[8], -(B[5] * B[7]));
    c[1] = m01 + m02 + m12;
    const double d0 = fma(B[4], B[8], -(B[5] * B[7]));
    const double d1 = fma(B[3], B[8], -(B[5] * B[6]));
    const double d2 = fma(B[3], B[7], -(B[4] * B[6]));
    c[0] = -(fma(B[0], d0, fma(-B[1], d1, B[2] * d2)));
  } else if (n == 4) {
    // p(y) = y^4 + c2 y^2 + c1 y + c0 from the principal minors of B (about half the
    // work of the general recurrence below):
    //   c2 = sum of principal 2 x 2 minors, c1 = -sum of principal 3 x 3 minors, c0 = det B
#define B_(i, j) B[(i) * 4 + (j)]
#define MIN2(i, j) fma(B_(i, i), B_(j, j), -(B_(i, j) * B_(j, i)))
    c[2] = ((MIN2(0, 1) + MIN2(0, 2)) + (MIN2(0, 3) + MIN2(1, 2))) + (MIN2(1, 3) + MIN2(2, 3));
#define DET3(i, j, k)                                                                             \
  fma(B_(i, i), fma(B_(j, j), B_(k, k), -(B_(j, k) * B_(k, j))),                                  \
      fma(-B_(i, j), fma(B_(j, i), B_(k, k), -(B_(j, k) * B_(k, i))),                             \
          B_(i, k) * fma(B_(j, i), B_(k, j), -(B_(j, j) * B_(k, i)))))
    c[1] = -((DET3(0, 1, 2) + DET3(0, 1, 3)) + (DET3(0, 2, 3) + DET3(1, 2, 3)));
    // det by complementary 2 x 2 minors of rows (0,1) and (2,3)
#define M01(p, q) fma(B_(0, p), B_(1, q), -(B_(0, q) * B_(1, p)))
#define M23(p, q) fma(B_(2, p), B_(3, q), -(B_(2, q) * B_(3, p)))
    c[0] = fma(M01(0, 1), M23(2, 3),
               fma(-M01(0, 2), M23(1, 3),
                   fma(M01(0, 3), M23(1, 2),
                       fma(M01(1, 2), M23(0, 3), fma(-M01(1, 3), M23(0, 2), M01(2, 3) * M23(0, 1))))));
#undef M01
#undef M23
#undef DET3
#undef MIN2
#undef B_
  } else {
  // Faddeev-LeVerrier: M_1 = B, c_{n-1} = -tr M_1 (= 0); M_k = B (M_{k-1} + c_{n-k+1} I),
  // c_{n-k} = -tr(M_k)/k
  double M[n * n];
#pragma unroll
  for (int i = 0; i < n * n; i++)
    M[i] = B[i];
#pragma unroll
  for (int k = 2; k <= n; k++) {
#pragma unroll
    for (int i = 0; i < n; i++)
      M[i * n + i] += c[n - k + 1];
    if (k < n) {
      // M <- B M, column by column in place
#pragma unroll
      for (int j = 0; j < n; j++) {
        double col[n];
#pragma unroll
        for (int i = 0; i < n; i++) {
          double acc = 0.;
#pragma unroll
          for (int l = 0; l < n; l++)
            acc = fma(B[i * n + l], M[l * n + j], acc);
          col[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < n; i++)
          M[i * n + j] = col[i];
      }
      double tr = 0.;
#pragma unroll
      for (int i = 0; i < n; i++)
        tr += M[i * n + i];
      c[n - k] = -tr * (1. / k);
    } else {
      // only the trace of B M is needed
      double tr = 0.;
#pragma unroll
      for (int i = 0; i < n; i++)
#pragma unroll
        for (int l = 0; l < n; l++)
          tr = fma(B[i * n + l], M[l * n + i], tr);
      c[0] = -tr * (1. / n);
    }
  }
  }
}

// Phase 3: given the outermost real roots yp >= ym of p, certify that the spectral
// radius |mu + y| is attained by one of them and update the warm-start record.
template <int n>
EIG_FN bool poly_finish(double mu, const double *c, double yp, double ym, double &rho,
                        EigGuess *guess) {
  // both searches ending on the same root means a single real root (n odd) or a
  // multiple one
  const double ayp = fabs(yp), aym = fabs(ym);
  const double ymax = ayp > aym ? ayp : aym;
  const bool single = !(yp - ym > 1e-7 * ymax);
  if (single && n != 3)
    return false;

  // the larger of the two outer roots in |mu + y| is the candidate; certify its
  // conditioning: kappa = sum |c_k| |y|^k / (|y| |p'(y)|)
  const double yw = fabs(mu + yp) >= fabs(mu + ym) ? yp : ym;
  const double best = fabs(mu + yw);
  {
    double dp = 0., p = 1., ab = 1.;
    const double ay = fabs(yw);
#pragma unroll
    for (int k = n - 1; k >= 0; k--) {
      dp = fma(dp, yw, p);
      p = fma(p, yw, c[k]);
      ab = fma(ab, ay, fabs(c[k]));
    }
    // a root this ill-conditioned (multiple or nearly so) is not certified here
    if (!(ab <= 400. * ay * fabs(dp)))
      return false;
  }

  // Remaining roots.  Real ones lie between ym and yp and cannot exceed the
  // outer roots in |mu + y|; complex pairs must be looked at.
  double b[n]; // quotient of p by (y - yp): y^(n-1) + b[n-2] y^(n-2) + ... + b[0]
  {
    double carry = 1.;
#pragma unroll
    for (int k = n - 1; k >= 1; k--) {
      carry = fma(carry, yp, c[k]);
      b[k - 1] = carry;
    }
  }
  if (n == 3) {
    // y^2 + b[1] y + b[0]
    const double disc = b[1] * b[1] - 4. * b[0];
    if (single) {
      if (!(disc < 0.))
        return false; // multiple real root: not certified here
      if (pair_modulus2(mu, b[1], b[0]) > 0.96 * best * best)
        return false;
    } else if (disc < 0.) {
      return false; // inconsistent with two distinct outer real roots
    }
    rho = best;
    if (guess) {
      guess->dy = guess->valid ? sel_max(fabs(yp - guess->yp), fabs(ym - guess->ym)) : 2.5e-4 * ymax;
      guess->yp = yp;
      guess->ym = ym;
      guess->valid = 1;
    }
    return true;
  }
  double e[n]; // quotient of that by (y - ym): y^(n-2) + e[n-3] y^(n-3) + ... + e[0]
  {
    double carry = 1.;
#pragma unroll
    for (int k = n - 2; k >= 1; k--) {
      carry = fma(carry, ym, b[k]);
      e[k - 1] = carry;
    }
  }
  if (n == 4) {
    const double disc = e[1] * e[1] - 4. * e[0];
    if (disc < 0. && pair_modulus2(mu, e[1], e[0]) > 0.96 * best * best)
      return false;
  } else { // n == 5: cubic y^3 + e[2] y^2 + e[1] y + e[0]
    // cheap certificate first (Fujiwara's bound on every root of the cubic):
    // for Euler-type spectra (v-c, v, v, v, v+c) the cubic is ~ y^3
    const double fb = 2. * fmax(fabs(e[2]), fmax(sqrt(fabs(e[1])), cbrt(0.5 * fabs(e[0]))));
    if (!(fabs(mu) + fb <= 0.98 * best)) {
      double ce[3] = {e[0], e[1], e[2]};
      double yr;
      if (!PolyRoots<3>::iterate(ce, fb * (1. + 1e-12), false, yr))
        return false;
      const double g1 = ce[2] + yr;
      const double g0 = fma(g1, yr, ce[1]);
      const double disc = g1 * g1 - 4. * g0;
      if (disc < 0. && pair_modulus2(mu, g1, g0) > 0.96 * best * best)
        return false;
    }
  }
  rho = best;
  if (guess) {
    guess->dy = guess->valid ? sel_max(fabs(yp - guess->yp), fabs(ym - guess->ym)) : 2.5e-4 * ymax;
    guess->yp = yp;
    guess->ym = ym;
    guess->valid = 1;
  }
  return true;
}

// A bound on every root of p from its coefficients alone (Fujiwara):
// |y| <= 2 max_k |c_{n-k}|^(1/k), with c_0 halved.  Only the rare searches whose
// warm start is not certified start from here.  0 means p = y^n.
template <int n> EIG_FN_NOINLINE double poly_root_bound(PolyCoef<n> pc) {
  const double *c = pc.v;
  double b = 0.;
#pragma unroll
  for (int k = 1; k <= n; k++) {
    const double ck = fabs(c[n - k]) * (k == n ? 0.5 : 1.);
    const double r = k == 1 ? ck : (k == 2 ? sqrt(ck) : (k == 3 ? cbrt(ck) : (k == 4 ? sqrt(sqrt(ck)) : pow(ck, 1. / k))));
    b = r > b ? r : b; // (a NaN coefficient: caught by the caller's range check on p)
    if (!(ck <= 1e300))
      b = ck;
  }
  return 2. * b;
}

// Phase 2 + 3 for NS independent polynomials at once (NS = 2: the left and the right
// state of a face point).  The 2 NS root searches (largest root of p and of
// (-1)^n p(-y) per side) take their certified warm-start step in lockstep —
// straight-line code over independent dependency chains, which is what the FP64
// pipe needs with four warps per scheduler — and only the searches that have not
// converged by then continue one by one.  ok[s] = false: side s is left to the
// general QR iteration.
template <int n, int NS>
EIG_FN void poly_solve(const double *mu, const double (*c)[n], EigGuess *const *guess, double *rho,
                       bool *ok) {
  constexpr int NC = 2 * NS;
  double cc[NC][n], xg[NC];
#pragma unroll
  for (int s = 0; s < NS; s++) {
#pragma unroll
    for (int k = 0; k < n; k++) {
      cc[2 * s][k] = c[s][k];
      cc[2 * s + 1][k] = ((n - k) & 1) ? -c[s][k] : c[s][k];
    }
    // Starting points.  Warm: the previous solve's outer roots moved outwards by four
    // times the last observed change of the roots (a sequence of nearby states), within
    // [1e-7, 1e-3] relative; first solve of a sequence: the Laguerre-Samuelson bound
    // sqrt((n-1)/n sum y_i^2) = sqrt(-2 c_{n-2} (n-1)/n) of a real zero-mean spectrum.
    // Either is used only under the Budan-Fourier certificate below.
    xg[2 * s] = xg[2 * s + 1] = -1.;
    const EigGuess *g = guess[s];
    if (g && g->valid) {
      const double gp = g->yp, gm = -g->ym, d4 = 4. * g->dy;
      const double lp = 1e-7 * gp, hp = 1e-3 * gp, lm = 1e-7 * gm, hm = 1e-3 * gm;
      xg[2 * s] = gp + (d4 > hp ? hp : (d4 > lp ? d4 : lp)); // (plain selects: no NaNs here)
      xg[2 * s + 1] = gm + (d4 > hm ? hm : (d4 > lm ? d4 : lm));
    } else if (c[s][n - 2] < 0.) {
      xg[2 * s] = xg[2 * s + 1] = sqrt(-2. * c[s][n - 2] * ((n - 1.) / n)) * (1. + 1e-3);
    }
  }
  // The Taylor coefficients of p at xg serve twice: all positive certifies that no real
  // root lies to the right of xg (Budan-Fourier), and t0, t1, t2 = p, p', p''/2 give the
  // first step for free — Halley's a = p p' / (p'^2 - p p''/2) (cubic, one reciprocal, no
  // square root; it lies between the Newton and the Laguerre step, so it is monotone in
  // the same regime).  A step below 4e-7 |x| leaves an error ~ (step / |x|)^3 times the
  // root's conditioning, i.e. rounding level for every root poly_finish certifies.
  double t[NC][n + 1];
  bool cert[NC];
#pragma unroll
  for (int k = 0; k < NC; k++) {
#pragma unroll
    for (int j = 0; j < n; j++)
      t[k][j] = cc[k][j];
    t[k][n] = 1.;
    cert[k] = xg[k] > 0.;
  }
#pragma unroll
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int j = n - 1; j >= i; j--)
#pragma unroll
      for (int k = 0; k < NC; k++)
        t[k][j] = fma(xg[k], t[k][j + 1], t[k][j]);
#pragma unroll
    for (int k = 0; k < NC; k++)
      cert[k] = cert[k] && (t[k][i] > 0.);
  }
  double den[NC], step[NC];
#pragma unroll
  for (int k = 0; k < NC; k++)
    den[k] = fma(t[k][1], t[k][1], -(t[k][0] * t[k][2]));
#pragma unroll
  for (int k = 0; k < NC; k++)
    step[k] = (t[k][0] * t[k][1]) * (1. / den[k]);
  double x[NC], root[NC];
  bool newton[NC];
  int st[NC]; // 1 root found, 0 iterate from x, -1 start from a bound, 2 failed
#pragma unroll
  for (int k = 0; k < NC; k++) {
    x[k] = xg[k];
    root[k] = 0.;
    newton[k] = false;
    st[k] = cert[k] ? 0 : -1;
    // (t0, t1 > 0 when certified; den <= 0, an overflow or a NaN leave it to iterate())
    if (cert[k] && den[k] > 0. && step[k] <= 1e300) {
      x[k] = xg[k] - step[k];
      const double ax = fabs(x[k]);
      if (step[k] <= 4e-7 * ax) {
        root[k] = x[k];
        st[k] = 1;
      }
      newton[k] = step[k] <= 0.02 * ax;
    }
  }
  bool zero[NS];
#pragma unroll
  for (int s = 0; s < NS; s++)
    zero[s] = false;
#pragma unroll
  for (int k = 0; k < NC; k++) {
    if (st[k] < 0) {
      // p = y^n (a zero matrix up to the shift, e.g. a flux without a gradient term in
      // this direction): every eigenvalue equals mu
      bool allzero = true;
#pragma unroll
      for (int j = 0; j < n; j++)
        allzero = allzero && cc[k][j] == 0.;
      double bnd = 0.;
      if (!allzero) {
        PolyCoef<n> pc;
#pragma unroll
        for (int j = 0; j < n; j++)
          pc.v[j] = cc[k][j];
        bnd = poly_root_bound<n>(pc);
      }
      if (allzero) {
        zero[k / 2] = true;
        st[k] = 1;
      } else if (!(bnd <= 1e150)) {
        st[k] = 2; // NaN or huge: leave it to the general routine
      } else {
        x[k] = bnd * (1. + 1e-12);
        newton[k] = false;
        st[k] = 0;
      }
    }
    if (st[k] == 0) {
#if PDE_EIG_ITER_NOINLINE
      PolyCoef<n> pc;
#pragma unroll
      for (int j = 0; j < n; j++)
        pc.v[j] = cc[k][j];
      const PolyIterResult ir = PolyRoots<n>::iterate_call(pc, x[k], newton[k]);
      root[k] = ir.root;
      st[k] = ir.ok ? 1 : 2;
#else
      st[k] = PolyRoots<n>::iterate(cc[k], x[k], newton[k], root[k]) ? 1 : 2;
#endif
    }
  }
#pragma unroll
  for (int s = 0; s < NS; s++) {
    if (zero[s]) {
      rho[s] = fabs(mu[s]);
      ok[s] = true;
    } else {
      ok[s] = st[2 * s] == 1 && st[2 * s + 1] == 1 &&
              poly_finish<n>(mu[s], c[s], root[2 * s], -root[2 * s + 1], rho[s], guess[s]);
    }
  }
}

template <int n>
EIG_FN bool spectral_radius_poly(const double *A, double &rho, EigGuess *guess = nullptr) {
  double mu[1], c[1][n], r[1];
  bool ok[1];
  EigGuess *g[1] = {guess};
  poly_setup<n>(A, mu[0], c[0]);
  poly_solve<n, 1>(mu, c, g, r, ok);
  rho = r[0];
  return ok[0];
}

// Two matrices at once (the left and the right state of a face point): ok[s] = false
// leaves side s to the caller's general routine.
template <int n>
EIG_FN void spectral_radius_poly_pair(const double *A0, const double *A1, double *rho, bool *ok,
                                      EigGuess *g0, EigGuess *g1) {
  double mu[2], c[2][n];
  EigGuess *g[2] = {g0, g1};
  poly_setup<n>(A0, mu[0], c[0]);
  poly_setup<n>(A1, mu[1], c[1]);
  poly_solve<n, 2>(mu, c, g, rho, ok);
}

// Parlett-Reinsch balancing by powers of two (an exact similarity): brings row
// and column norms together so that conserved-variable Jacobians, whose entries
// span orders of magnitude at high Mach number, lose nothing in either path.
template <int n> EIG_FN void balance(double *a) {
  for (int sweep = 0; sweep < 6; sweep++) {
    bool done = true;
#pragma unroll
    for (int i = 0; i < n; i++) {
      double c = 0., r = 0.;
#pragma unroll
      for (int j = 0; j < n; j++)
        if (j != i) {
          c += fabs(a[j * n + i]);
          r += fabs(a[i * n + j]);
        }
      if (c > 0. && r > 0. && c <= 1e300 && r <= 1e300) {
        double g = 0.5 * r, f = 1.;
        const double s0 = c + r;
        int guard = 0;
        while (c < g && guard++ < 600) {
          f *= 2.;
          c *= 4.;
        }
        g = 2. * r;
        while (c >= g && guard++ < 1200) {
          f *= 0.5;
          c *= 0.25;
        }
        if ((c + r) / f < 0.95 * s0) {
          done = false;
          const double fi = 1. / f;
#pragma unroll
          for (int j = 0; j < n; j++)
            a[i * n + j] *= fi;
#pragma unroll
          for (int j = 0; j < n; j++)
            a[j * n + i] *= f;
        }
      }
    }
    if (done)
      break;
  }
}

// Run-time sized balancing of the leading m x m block (row pitch ld); same steps as balance<n>.
// (three sweeps: the GPR matrices are done after three — the fourth only confirms it, at 3.5 %
//  of the wave-speed kernel; a matrix that would need more is left slightly less balanced, which
//  costs the Hessenberg reduction a little accuracy and nothing else: scaling is an exact
//  similarity whenever it stops)
#ifndef PDE_EIG_BAL_SWEEPS
#define PDE_EIG_BAL_SWEEPS 3
#endif
EIG_FN void balance_rt(double *a, const int ld, const int m) {
  for (int sweep = 0; sweep < PDE_EIG_BAL_SWEEPS; sweep++) {
    bool done = true;
    for (int i = 0; i < m; i++) {
      double c = 0., r = 0.;
      // (two branch-free loops, the same sums in the same order)
      for (int j = 0; j < i; j++) {
        c += fabs(a[j * ld + i]);
        r += fabs(a[i * ld + j]);
      }
      for (int j = i + 1; j < m; j++) {
        c += fabs(a[j * ld + i]);
        r += fabs(a[i * ld + j]);
      }
      if (c > 0. && r > 0. && c <= 1e300 && r <= 1e300) {
        double g = 0.5 * r, f = 1.;
        const double s0 = c + r;
        int guard = 0;
        while (c < g && guard++ < 600) {
          f *= 2.;
          c *= 4.;
        }
        g = 2. * r;
        while (c >= g && guard++ < 1200) {
          f *= 0.5;
          c *= 0.25;
        }
        if ((c + r) / f < 0.95 * s0) {
          done = false;
          const double fi = 1. / f;
          for (int j = 0; j < m; j++)
            a[i * ld + j] *= fi;
          for (int j = 0; j < m; j++)
            a[j * ld + i] *= f;
        }
      }
    }
    if (done)
      break;
  }
}

// ---------------------------------------------------------------------------
// n > 5, after the Hessenberg reduction: the spectral radius from the characteristic
// polynomial instead of the QR iteration, under a certificate.
//
// For an upper Hessenberg H the leading principal minors p_k(y) = det(y I - H_k) obey
//   p_k = (y - h_kk) p_(k-1) - sum_{i<k} h_ik (h_(i+1,i) .. h_(k,k-1)) p_(i-1),
// a division-free recurrence that is indifferent to zero subdiagonal entries; carried out on
// coefficient arrays it gives the characteristic polynomial of the m x m block in ~m^3/6
// multiply-adds (222 for m = 11, where the QR iteration needs ~10 m^3).  Then, as for n <= 5:
//   1. the largest and the smallest real root by the monotone Laguerre iteration from outside
//      (warm-started from the previous solve of a sequence), on the mean-shifted polynomial;
//   2. each polished by one Newton step on H itself — the same recurrence evaluated at a
//      number, O(m^2) — whose size also bounds what the coefficient form lost;
//   3. the certificate: p divided by both roots, moved back to the unshifted variable and
//      scaled to the disk of radius 0.999 rho, must pass the Schur-Cohn test (all reflection
//      coefficients below one in modulus <=> every other eigenvalue, real or complex, lies
//      strictly inside that disk).  Then rho = max |outer root| is the spectral radius.
// Anything else — a dominant complex pair, a multiple or clustered outer root, a search that
// leaves the monotone regime, non-finite numbers — returns false and the caller continues with
// the QR iteration on the same (untouched) Hessenberg matrix.  For hyperbolic systems, whose
// fastest waves are simple and real, the certificate holds as a rule.
// guess (optional): outer eigenvalues of the previous matrix, unshifted.
// ---------------------------------------------------------------------------
EIG_FN bool poly_rt_iterate(const double *c, const int m, double x, double &root) {
  bool newton = false;
  for (int it = 0; it < 40; it++) {
    double p = 1., dp = 0., d2 = 0.; // p, p', p''/2 by Horner
    for (int k = m - 1; k >= 0; k--) {
      d2 = fma(d2, x, dp);
      dp = fma(dp, x, p);
      p = fma(p, x, c[k]);
    }
    if (!(p > 0.) || !(dp > 0.)) {
      // on (or a rounding error past) the root, or outside the monotone regime
      double ab = 1.;
      const double ax = fabs(x);
      for (int k = m - 1; k >= 0; k--)
        ab = fma(ab, ax, fabs(c[k]));
      if (fabs(p) <= 64. * DBL_EPS * ab && dp > 0.) {
        root = x - p / dp;
        return true;
      }
      return false;
    }
    double a;
    if (newton) {
      a = p / dp;
    } else {
      const double disc = (m - 1.) * ((m - 1.) * dp * dp - 2. * m * p * d2);
      a = disc > 0. ? m * p / (dp + sqrt(disc)) : p / dp;
    }
    if (!(a >= 0.) || !(a <= 1e300))
      return false;
    x -= a;
    const double ax = fabs(x);
    if (a <= 1e-5 * ax) {
      p = 1.;
      dp = 0.;
      for (int k = m - 1; k >= 0; k--) {
        dp = fma(dp, x, p);
        p = fma(p, x, c[k]);
      }
      if (!(dp > 0.))
        return false;
      root = x - p / dp;
      return true;
    }
    newton = a <= 0.02 * ax;
  }
  return false;
}

// p(lam) / p'(lam) for p = det(lam I - H), H the leading m x m Hessenberg block of a, at two
// arguments in one pass over H (the products h_ik h_(i+1,i) .. h_(k,k-1) do not depend on lam)
template <int NMAX>
EIG_FN void hess_newton_correction2(const double *a, const int ld, const int m, const double lam0,
                                    const double lam1, double &c0, double &c1) {
#define A_(i, j) a[(i) * ld + (j)]
  double pv0[NMAX + 1], dv0[NMAX + 1], pv1[NMAX + 1], dv1[NMAX + 1];
  pv0[0] = pv1[0] = 1.;
  dv0[0] = dv1[0] = 0.;
  for (int k = 1; k <= m; k++) {
    const double hkk = A_(k - 1, k - 1);
    const double s0 = lam0 - hkk, s1 = lam1 - hkk;
    double p0 = s0 * pv0[k - 1], d0 = fma(s0, dv0[k - 1], pv0[k - 1]);
    double p1 = s1 * pv1[k - 1], d1 = fma(s1, dv1[k - 1], pv1[k - 1]);
    double t = 1.;
    for (int i = k - 1; i >= 1; i--) {
      t *= A_(i, i - 1);
      if (t == 0.)
        break;
      const double g = A_(i - 1, k - 1) * t;
      if (g != 0.) {
        p0 = fma(-g, pv0[i - 1], p0);
        d0 = fma(-g, dv0[i - 1], d0);
        p1 = fma(-g, pv1[i - 1], p1);
        d1 = fma(-g, dv1[i - 1], d1);
      }
    }
    pv0[k] = p0;
    dv0[k] = d0;
    pv1[k] = p1;
    dv1[k] = d1;
  }
  c0 = pv0[m] / dv0[m];
  c1 = pv1[m] / dv1[m];
#undef A_
}

#ifndef PDE_EIG_HESS_POLY
#define PDE_EIG_HESS_POLY 1 // 0: n > 5 always through the QR iteration
#endif
template <int NMAX>
EIG_FN_NOINLINE bool spectral_radius_hess_poly(const double *a, const int ld, const int m,
                                               double &rho, EigGuess *guess) {
#define A_(i, j) a[(i) * ld + (j)]
  if (m < 2 || m > NMAX)
    return false;
  double mu = 0.;
  for (int i = 0; i < m; i++)
    mu += A_(i, i);
  mu *= 1. / m;
  // --- coefficients: P_k (degree k, ascending powers, leading 1) at P[k (k + 1) / 2]
  double P[(NMAX + 1) * (NMAX + 2) / 2];
  P[0] = 1.;
  for (int k = 1; k <= m; k++) {
    double *pk = P + k * (k + 1) / 2;
    const double *pp = P + (k - 1) * k / 2;
    const double hkk = A_(k - 1, k - 1) - mu;
    pk[k] = 1.;
    for (int j = k - 1; j >= 1; j--)
      pk[j] = fma(-hkk, pp[j], pp[j - 1]);
    pk[0] = -hkk * pp[0];
    double t = 1.;
    for (int i = k - 1; i >= 1; i--) {
      t *= A_(i, i - 1);
      if (t == 0.)
        break;
      const double g = A_(i - 1, k - 1) * t;
      if (g != 0.) {
        const double *pi = P + (i - 1) * i / 2;
        for (int j = 0; j < i; j++)
          pk[j] = fma(-g, pi[j], pk[j]);
      }
    }
  }
  const double *c = P + m * (m + 1) / 2;
  double cn[NMAX + 1]; // (-1)^m p(-y): its largest root is minus the smallest root of p
  for (int k = 0; k <= m; k++)
    cn[k] = ((m - k) & 1) ? -c[k] : c[k];

  // --- the outer real roots
  double cold = -1.;
  if (c[m - 2] < 0.) // Laguerre-Samuelson: bounds a real zero-mean spectrum
    cold = sqrt(-2. * c[m - 2] * ((m - 1.) / m)) * (1. + 1e-3);
  double yp = 0., ym = 0.;
  bool okp = false, okm = false;
  if (guess && guess->valid) {
    const double gp = guess->yp - mu, gm = -(guess->ym - mu), d4 = 4. * guess->dy;
    const double sc = sel_max(fabs(gp), fabs(gm));
    const double lo = 1e-7 * sc, hi = 1e-3 * sc;
    const double mg = d4 > hi ? hi : (d4 > lo ? d4 : lo);
    okp = poly_rt_iterate(c, m, gp + mg, yp);
    okm = poly_rt_iterate(cn, m, gm + mg, ym);
  }
  if (!okp) {
    if (!(cold > 0.) || !poly_rt_iterate(c, m, cold, yp))
      return false;
  }
  if (!okm) {
    if (!(cold > 0.) || !poly_rt_iterate(cn, m, cold, ym))
      return false;
  }
  ym = -ym;
  const double ymax = sel_max(fabs(yp), fabs(ym));
  if (!(yp - ym > 1e-7 * ymax)) // one real root only, or a multiple one
    return false;
  // --- polish on H (unshifted variable)
  double lp = mu + yp, lm = mu + ym;
  {
    double cp, cm;
    hess_newton_correction2<NMAX>(a, ld, m, lp, lm, cp, cm);
    if (!(fabs(cp) <= 1e-9 * ymax) || !(fabs(cm) <= 1e-9 * ymax))
      return false;
    lp -= cp;
    lm -= cm;
  }
  const double best = sel_max(fabs(lp), fabs(lm));
  if (!(best > 0.) || !(best <= 1e150))
    return false;
  // --- the other m - 2 roots: inside the disk of radius 0.999 best?
  const int dg = m - 2;
  if (dg > 0) {
    double f[NMAX + 1], g2[NMAX + 1];
    // p / (y - yp) / (y - ym), backward deflation (stable for the roots of largest modulus):
    // c_0 = -r q_0, c_k = q_(k-1) - r q_k
    if (yp == 0. || ym == 0.)
      return false;
    {
      const double ri = 1. / yp;
      double qk = -c[0] * ri;
      f[0] = qk;
      for (int k = 1; k < m - 1; k++) {
        qk = (qk - c[k]) * ri;
        f[k] = qk;
      }
      f[m - 1] = 1.;
    }
    {
      const double ri = 1. / ym;
      double qk = -f[0] * ri;
      g2[0] = qk;
      for (int k = 1; k < m - 2; k++) {
        qk = (qk - f[k]) * ri;
        g2[k] = qk;
      }
      g2[dg] = 1.;
    }
    // Taylor shift to the unshifted variable lam = y + mu: e(lam - mu)
    for (int i = 0; i < dg; i++)
      for (int j = dg - 1; j >= i; j--)
        g2[j] = fma(-mu, g2[j + 1], g2[j]);
    // scale to the unit disk: z = lam / R
    {
      const double Ri = 1. / (0.999 * best);
      double sc = 1.;
      for (int j = dg - 1; j >= 0; j--) {
        sc *= Ri;
        g2[j] *= sc;
      }
    }
    // Schur-Cohn: kappa = f_0 / f_d, f <- (f_(j+1) - kappa f_(d-1-j))_j
    double *cur = g2, *nxt = f;
    for (int d = dg; d >= 1; d--) {
      const double kap = cur[0] / cur[d];
      if (!(fabs(kap) < 1.))
        return false;
      for (int j = 0; j < d; j++)
        nxt[j] = fma(-kap, cur[d - 1 - j], cur[j + 1]);
      double *tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
  }
  rho = best;
  if (guess) {
    guess->dy = guess->valid ? sel_max(fabs(lp - guess->yp), fabs(lm - guess->ym)) : 2.5e-4 * ymax;
    guess->yp = lp;
    guess->ym = lm;
    guess->valid = 1;
  }
  return true;
#undef A_
}

// Larger systems (n > 5; the GPR model has n = 17): most of a conserved-variable system
// matrix dF_d/dQ + B_d is structurally zero — the 2-D GPR matrices have 50-110 non-zeros of
// 289 — and a third to a half of its eigenvalues sit isolated on the diagonal (the
// distortion and thermal-impulse rows that direction d does not couple: eigenvalue v_d).
// The permutation step of the classical balancing algorithm (Parlett & Reinsch; LAPACK
// dgebal, job P) finds them: a row whose off-diagonal entries vanish inside the active
// block is exchanged to its end, then a column whose off-diagonal entries vanish to its
// front, repeatedly — an exact similarity — leaving
//             | T1  X   Y  |
//   P A P^T = |  0   B   Z  |     spec(A) = diag(T1) u spec(B) u diag(T2)
//             |  0   0   T2 |
// with T1, T2 upper triangular.  (Keeping non-zero counts per row and column instead of
// exchanging rows was measured slower on the GPU: 11.4 against 9.7 ms for the GPR wave
// speeds at 256^2 — its gather through an index list is one more level of dynamic
// indexing in local memory.)  Only B (9 x 9 to 11 x 11 of 17 x 17 for GPR) goes through
// scaling, Hessenberg reduction and the QR iteration, whose cost is cubic in its size:
// measured on B200, GPR 256^2: k_wavespeeds 11.0 -> see profiles/.  a is destroyed.
// The m x m active block at the front of a (row pitch m): scaling, Hessenberg form, then the
// certified characteristic polynomial or the QR iteration.
// Row pitch of the active block.  Thread-local arrays are interleaved over the lanes of a warp,
// one element index per 256-byte row: contiguity within a thread buys nothing, whereas the same
// pitch in every lane keeps entry (i, j) of all lanes in one row whatever their block sizes.
#ifndef PDE_EIG_PITCH_FULL
#define PDE_EIG_PITCH_FULL 1 // 0: pitch m (the block contiguous)
#endif
#define EIG_PITCH(n, m) (PDE_EIG_PITCH_FULL ? (n) : (m))
template <int n>
EIG_FN double spectral_radius_active_block(double *a, const int m, int *path, EigGuess *guess) {
  const int ld = EIG_PITCH(n, m);
  balance_rt(a, ld, m);
  hessenberg_rt(a, ld, m);
#if PDE_EIG_HESS_POLY
  {
    double r;
    if (spectral_radius_hess_poly<n>(a, ld, m, r, guess)) {
      if (path)
        *path = 2;
      return r;
    }
  }
#endif
  if (guess)
    guess->valid = 0;
  return hqr_rt(a, ld, m);
}

template <int n>
EIG_FN_NOINLINE double spectral_radius_deflated_qr(double *a, int *path = nullptr,
                                                   EigGuess *guess = nullptr) {
#define A_(i, j) a[(i) * n + (j)]
  int lo = 0, hi = n - 1;
  double rad = 0.;
  auto exchange = [&](int j, int m) {
    if (j == m)
      return;
    for (int i = 0; i < n; i++) {
      const double t = A_(i, j);
      A_(i, j) = A_(i, m);
      A_(i, m) = t;
    }
    for (int i = 0; i < n; i++) {
      const double t = A_(j, i);
      A_(j, i) = A_(m, i);
      A_(m, i) = t;
    }
  };
  // rows isolating an eigenvalue go to the end of the active block
  for (bool found = true; found && hi >= lo;) {
    found = false;
    for (int j = hi; j >= lo; j--) {
      bool zero = true;
      for (int i = lo; i <= hi; i++)
        if (i != j && A_(j, i) != 0.) {
          zero = false;
          break;
        }
      if (zero) {
        exchange(j, hi);
        rad = sel_max(rad, fabs(A_(hi, hi)));
        hi--;
        found = true;
        break;
      }
    }
  }
  // columns isolating an eigenvalue go to its front
  for (bool found = true; found && hi >= lo;) {
    found = false;
    for (int j = lo; j <= hi; j++) {
      bool zero = true;
      for (int i = lo; i <= hi; i++)
        if (i != j && A_(i, j) != 0.) {
          zero = false;
          break;
        }
      if (zero) {
        exchange(j, lo);
        rad = sel_max(rad, fabs(A_(lo, lo)));
        lo++;
        found = true;
        break;
      }
    }
  }
  const int m = hi - lo + 1;
  if (m <= 0)
    return rad;
  if (m == 1)
    return sel_max(rad, fabs(A_(lo, lo)));
  // B moves to the front of the array with row pitch m (destination index <= source index,
  // rows and columns ascending: in place), so that the iteration touches m^2 contiguous
  // doubles of this thread's local memory instead of a window of the n^2
  if (m < n) {
    const int ld = EIG_PITCH(n, m);
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++)
        a[i * ld + j] = A_(lo + i, lo + j);
  }
  return sel_max(rad, spectral_radius_active_block<n>(a, m, path, guess));
#undef A_
}

// Out-of-line entry for callers that assemble the active block themselves (kernels.cuh: the
// two-pass Jacobian): a holds the m x m block with row pitch EIG_PITCH(n, m), rad the isolated
// eigenvalues.
template <int n>
EIG_FN_NOINLINE double spectral_radius_compact(double *a, const int m, const double rad,
                                               EigGuess *guess = nullptr) {
  if (m <= 0)
    return rad;
  if (m == 1)
    return sel_max(rad, fabs(a[0]));
  return sel_max(rad, spectral_radius_active_block<n>(a, m, nullptr, guess));
}

#ifndef PDE_EIG_MASK
#define PDE_EIG_MASK 1 // 0: the permutation step by row / column exchanges in memory (above)
#endif
// The same permutation step on the sparsity pattern alone: rm[i] has bit j set where A(i, j)
// may be non-zero (n <= 32).  The device builds the pattern in registers while it differences
// the Jacobian, so the search for isolated eigenvalues — iteratively removing the sinks, then
// the sources, of the graph of A; the remaining set does not depend on the order — is integer
// work on n words: no scan of the matrix, no row or column exchange in local memory.  The
// active block is then gathered in ascending index order (destination index <= source index:
// in place).  rm == nullptr: the pattern is read off the matrix.
// the active set (bit j: index j stays) of the pattern r; fully unrolled, r stays in registers
template <int n> EIG_FN unsigned mask_active_set(const unsigned *r) {
  unsigned act = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
  // rows that vanish off the diagonal inside the active set
  for (bool found = true; found;) {
    found = false;
#pragma unroll
    for (int j = 0; j < n; j++) {
      const unsigned bj = 1u << j;
      if ((act & bj) && !(r[j] & act & ~bj)) {
        act &= ~bj;
        found = true;
      }
    }
  }
  // columns that do
  for (bool found = true; found && act;) {
    found = false;
    unsigned any = 0;
#pragma unroll
    for (int j = 0; j < n; j++) {
      const unsigned bj = 1u << j;
      if (act & bj)
        any |= r[j] & ~bj;
    }
    const unsigned iso = act & ~any;
    if (iso) {
      act &= ~iso;
      found = true;
    }
  }
  return act;
}

template <int n>
EIG_FN_NOINLINE double spectral_radius_masked(double *a, const unsigned *rm, int *path = nullptr,
                                              EigGuess *guess = nullptr) {
#define A_(i, j) a[(i) * n + (j)]
  unsigned r[n];
  if (rm) {
    for (int i = 0; i < n; i++)
      r[i] = rm[i];
  } else {
    for (int i = 0; i < n; i++) {
      unsigned b = 0;
      for (int j = 0; j < n; j++)
        b |= (A_(i, j) != 0. ? 1u : 0u) << j;
      r[i] = b;
    }
  }
  const unsigned act = mask_active_set<n>(r);
  double rad = 0.;
  for (int j = 0; j < n; j++)
    if (!(act & (1u << j)))
      rad = sel_max(rad, fabs(A_(j, j)));
  int idx[n];
  int m = 0;
  for (int j = 0; j < n; j++)
    if (act & (1u << j))
      idx[m++] = j;
  if (m == 0)
    return rad;
  if (m == 1)
    return sel_max(rad, fabs(A_(idx[0], idx[0])));
  if (m < n) {
    const int ld = EIG_PITCH(n, m);
    for (int i = 0; i < m; i++) {
      const double *row = a + idx[i] * n;
      for (int j = 0; j < m; j++)
        a[i * ld + j] = row[idx[j]];
    }
  }
  return sel_max(rad, spectral_radius_active_block<n>(a, m, path, guess));
#undef A_
}

// The general routine: balancing + QR iteration.  Only here does the matrix need an
// address (the QR iteration indexes it dynamically); copying with static indices
// keeps the caller's `a` in registers.
#ifndef PDE_EIG_DEFLATE
#define PDE_EIG_DEFLATE 1 // 0: n > 5 as in round 1 (scaling + QR iteration on the full matrix)
#endif
template <int n>
EIG_FN_NOINLINE double spectral_radius_balanced_qr(double *a, int *path = nullptr,
                                                   EigGuess *guess = nullptr,
                                                   const unsigned *rm = nullptr) {
#if PDE_EIG_DEFLATE
  if (n > 5) {
#if PDE_EIG_MASK
    if (n <= 32)
      return spectral_radius_masked<n>(a, rm, path, guess);
#endif
    return spectral_radius_deflated_qr<n>(a, path, guess);
  }
#endif
  if (guess)
    guess->valid = 0;
  balance<n>(a);
  return spectral_radius_qr<n>(a);
}
template <int n> EIG_FN double spectral_radius_general(const double *a) {
  // (everything cold stays out of line: ncu showed the hot path's instruction fetches
  //  stalling at the reconvergence points behind inlined cold blocks)
  double tmp[n * n];
#pragma unroll
  for (int i = 0; i < n * n; i++)
    tmp[i] = a[i];
  return spectral_radius_balanced_qr<n>(tmp);
}

// rm (optional, n > 5): the sparsity pattern of a by rows, see spectral_radius_masked
template <int n>
EIG_FN double spectral_radius(double *a, int *path = nullptr, EigGuess *guess = nullptr,
                              const unsigned *rm = nullptr) {
  if (n == 1)
    return fabs(a[0]);
  if (n == 2) {
    // closed form on the shifted matrix: y^2 = b00^2 + b01 b10
    const double mu = 0.5 * (a[0] + a[3]);
    const double b00 = a[0] - mu;
    const double s2 = fma(b00, b00, a[1] * a[2]);
    if (s2 >= 0.)
      return fabs(mu) + sqrt(s2);
    return sqrt(fma(mu, mu, -s2));
  }
#if !PDE_EIG_QR_ONLY
  if (n >= 3 && n <= 5) {
    double rho;
    if (spectral_radius_poly<(n >= 3 && n <= 5) ? n : 3>(a, rho, guess)) {
      if (path)
        *path = 1;
      return rho;
    }
  }
#endif
  if (path)
    *path = 0;
  // Larger matrices live in local memory already (they are indexed dynamically when they are
  // built): iterate in place instead of on a copy — `a` is destroyed, as documented — which
  // halves the local-memory footprint and traffic of the n = 17 wave-speed kernels (ncu:
  // 28 GB of local-memory traffic reaching DRAM per launch at C4).
  if (n > 5)
    return spectral_radius_balanced_qr<n>(a, path, guess, rm);
  if (guess)
    guess->valid = 0;
  return spectral_radius_general<n>(a);
}

// ---------------------------------------------------------------------------
// D2 fast path (3 <= n <= 5): y = |A| x without an eigen-decomposition, for the spectrum of
// an Euler-type system — two simple outer real eigenvalues lam_m < lam_p and everything else
// in one tight cluster around a real centre a (the n - 2 copies of the convective speed, split
// at the 1e-8 level by the finite-difference Jacobian).  With the spectral projectors
//   P_i = adj(lam_i I - A) / p'(lam_i)            (simple eigenvalue lam_i)
// and the cluster component w = x - P_m x - P_p x,
//   |A| x = |lam_m| P_m x + |lam_p| P_p x + |a| w.
// adj(y I - B) x = sum_k y^k v_k with v_(n-1) = x, v_(k-1) = B v_k + c_k x (the
// Faddeev-LeVerrier recurrence applied to a vector: n - 1 matrix-vector products, B = A - mu I
// and c the characteristic polynomial poly_setup gives), so the whole product is ~10 n^2
// multiply-adds in registers where the real-Schur route (elmhes / hqr2 / back-substitution / QR
// solve, in local memory) takes thousands.  It is also the more accurate of the two on exactly
// this spectrum: a triple eigenvalue leaves hqr2's back-substituted vectors numerically dependent
// (sqrt(eps)-level errors, tests/test_eig.py), whereas the cluster never needs its own vectors.
// Certificates — any failure returns false and the caller takes the general routine:
//   * both outer roots found by the monotone iteration, distinct, well conditioned
//     (kappa eps <~ 1e-13);
//   * the remaining n - 2 roots within 1e-4 of the scale around their mean (Fujiwara's bound on
//     the remainder polynomial shifted to its mean), and that cluster separated from both
//     outer roots by more than 1e-3 of the scale;
//   * the cluster term: with the cluster clear of zero (|a| > twice that bound) all its
//     eigenvalues share the sign of a and |A| w = sign(a) A w exactly — no approximation, any
//     spread; around zero (gas at rest: a ~ 1e-9) it is |a| w, accepted only if A acts on w as
//     a I to 2e-7 of the scale, ||A w - a w|| <= 2e-7 scale ||x|| (a direct measurement; the
//     bound from the coefficients cannot resolve a cluster below eps^(1/3)).  The result then
//     differs from R |Lambda| R^-1 x by at most that spread times ||w|| — for a
//     finite-difference Jacobian its own rounding noise, on a term that is itself that small.
// ---------------------------------------------------------------------------
#ifndef PDE_ABS_POLY
#define PDE_ABS_POLY 1
#endif
template <int n> EIG_FN bool abs_matrix_apply_poly(const double *A, const double *x, double *y) {
  double mu, c[n];
  poly_setup<n>(A, mu, c);
  if (!(c[n - 2] < 0.))
    return false;
  const double x0 = sqrt(-2. * c[n - 2] * ((n - 1.) / n)) * (1. + 1e-3);
  double cn[n];
#pragma unroll
  for (int k = 0; k < n; k++)
    cn[k] = ((n - k) & 1) ? -c[k] : c[k];
  double yp, ym;
  if (!PolyRoots<n>::iterate(c, x0, false, yp))
    return false;
  if (!PolyRoots<n>::iterate(cn, x0, false, ym))
    return false;
  ym = -ym;
  const double ayp = fabs(yp), aym = fabs(ym);
  const double sc = fabs(mu) + (ayp > aym ? ayp : aym);
  if (!(sc > 0.) || !(sc <= 1e150))
    return false;
  // p' at both roots and their conditioning
  double dpp = 0., dpm = 0.;
  {
    double pp = 1., pm = 1., abp = 1., abm = 1.;
#pragma unroll
    for (int k = n - 1; k >= 0; k--) {
      dpp = fma(dpp, yp, pp);
      pp = fma(pp, yp, c[k]);
      abp = fma(abp, ayp, fabs(c[k]));
      dpm = fma(dpm, ym, pm);
      pm = fma(pm, ym, c[k]);
      abm = fma(abm, aym, fabs(c[k]));
    }
    if (!(abp <= 400. * ayp * fabs(dpp)) || !(abm <= 400. * aym * fabs(dpm)))
      return false;
  }
  // the remainder p / ((y - yp)(y - ym)): centre and radius of its roots
  double ay = 0., r = 0.;
  {
    double b[n], e[n];
    double carry = 1.;
#pragma unroll
    for (int k = n - 1; k >= 1; k--) {
      carry = fma(carry, yp, c[k]);
      b[k - 1] = carry;
    }
    carry = 1.;
#pragma unroll
    for (int k = n - 2; k >= 1; k--) {
      carry = fma(carry, ym, b[k]);
      e[k - 1] = carry;
    }
    if (n == 3) {
      ay = -e[0];
    } else if (n == 4) { // y^2 + e1 y + e0
      ay = -0.5 * e[1];
      r = sqrt(fabs(fma(-0.25 * e[1], e[1], e[0])));
    } else { // y^3 + e2 y^2 + e1 y + e0 about its mean: z^3 + t1 z + t0
      ay = -e[2] * (1. / 3.);
      const double t1 = fma(-e[2] * (1. / 3.), e[2], e[1]);
      const double t0 = fma(fma(fma(1., ay, e[2]), ay, e[1]), ay, e[0]);
      r = 2. * fmax(sqrt(fabs(t1)), cbrt(0.5 * fabs(t0)));
    }
  }
  if (!(r <= 1e-4 * sc) || !(yp - ay > 1e-3 * sc + r) || !(ay - ym > 1e-3 * sc + r))
    return false;
  // v_k = M_k x and adj(y I - B) x at both roots (Horner over k, leading term v_(n-1) = x)
  double v[n], gp[n], gm[n], xmax = 0.;
#pragma unroll
  for (int i = 0; i < n; i++) {
    v[i] = x[i];
    gp[i] = x[i];
    gm[i] = x[i];
    xmax = fmax(xmax, fabs(x[i]));
  }
#pragma unroll
  for (int k = n - 1; k >= 1; k--) {
    double t[n];
#pragma unroll
    for (int i = 0; i < n; i++) {
      double acc = fma(-mu, v[i], c[k] * x[i]);
#pragma unroll
      for (int j = 0; j < n; j++)
        acc = fma(A[i * n + j], v[j], acc);
      t[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < n; i++) {
      v[i] = t[i];
      gp[i] = fma(gp[i], yp, t[i]);
      gm[i] = fma(gm[i], ym, t[i]);
    }
  }
  const double rp = 1. / dpp, rm = 1. / dpm;
  const double a = mu + ay, lp = fabs(mu + yp), lm = fabs(mu + ym);
  double w[n];
#pragma unroll
  for (int i = 0; i < n; i++) {
    gp[i] *= rp;
    gm[i] *= rm;
    w[i] = (x[i] - gp[i]) - gm[i];
  }
  // A w - a w: how far A is from a I on the cluster component
  double dev = 0., wmax = 0., dv[n];
#pragma unroll
  for (int i = 0; i < n; i++) {
    double acc = -a * w[i];
#pragma unroll
    for (int j = 0; j < n; j++)
      acc = fma(A[i * n + j], w[j], acc);
    dv[i] = acc;
    dev = fmax(dev, fabs(acc));
    wmax = fmax(wmax, fabs(w[i]));
  }
  const double aa = fabs(a);
  // The cluster lies within r of a (a rigorous bound).  Clear of zero (|a| > 2 r), all its
  // eigenvalues have the sign of a and |A| w = sign(a) A w EXACTLY, whatever its spread
  // (a complex pair inside it: to second order in its imaginary part).  Around zero the
  // term is |a| w up to the spread itself, which must then be at rounding-noise level.
  const bool one_signed = aa > 2. * r && aa * wmax >= 8. * dev;
  if (!one_signed && !(dev <= 2e-7 * sc * xmax))
    return false;
  const double sg = a < 0. ? -1. : 1.;
#pragma unroll
  for (int i = 0; i < n; i++) {
    const double yc = one_signed ? sg * fma(a, w[i], dv[i]) : aa * w[i];
    y[i] = fma(lp, gp[i], fma(lm, gm[i], yc));
  }
  return true;
}

// ---------------------------------------------------------------------------
// D2: y = |A| x for a general real n x n matrix, |A| = R |Lambda| R^-1 — the
// dissipation matrix of the Osher and Roe fluxes (reference fluxes.cpp:36-41,
// 64-69: Eigen EigenSolver + complex column-pivoted QR solve, real part taken).
// Real arithmetic throughout: elimination to Hessenberg form with accumulated
// transformations, Francis double-shift QR to real Schur form with accumulation,
// back-substitution for the eigenvectors (the classical EISPACK elmhes / eltran /
// hqr2 sequence), giving A W = W D with W real: column j for a real eigenvalue,
// columns (j, j+1) = (Re v, Im v) for a complex pair, whose 2 x 2 block of D has
// modulus |lambda| times a rotation — so |A| = W diag(|lambda_j|) W^-1 with the
// same modulus on both columns of a pair, and y = W (|lambda| o (W^-1 x)) by one
// real LU solve with partial pivoting.  a is destroyed.  Returns false if the QR
// iteration did not converge (y is then left untouched).
// ---------------------------------------------------------------------------
template <int n> EIG_FN_NOINLINE bool abs_matrix_apply(double *a, const double *x, double *y) {
#define A_(i, j) a[(i) * n + (j)]
#define Z_(i, j) zz[(i) * n + (j)]
  double zz[n * n], wr[n], wi[n];
  int perm[n];
  if (n == 1) {
    y[0] = fabs(a[0]) * x[0];
    return true;
  }
  // --- elmhes with the permutation recorded
  for (int m = 1; m < n - 1; m++) {
    double xx = 0.;
    int i = m;
    for (int j = m; j < n; j++)
      if (fabs(A_(j, m - 1)) > fabs(xx)) {
        xx = A_(j, m - 1);
        i = j;
      }
    perm[m] = i;
    if (i != m) {
      for (int j = m - 1; j < n; j++) {
        double tmp = A_(i, j);
        A_(i, j) = A_(m, j);
        A_(m, j) = tmp;
      }
      for (int j = 0; j < n; j++) {
        double tmp = A_(j, i);
        A_(j, i) = A_(j, m);
        A_(j, m) = tmp;
      }
    }
    if (xx != 0.) {
      for (i = m + 1; i < n; i++) {
        double yy = A_(i, m - 1);
        if (yy != 0.) {
          yy /= xx;
          A_(i, m - 1) = yy;
          for (int j = m; j < n; j++)
            A_(i, j) -= yy * A_(m, j);
          for (int j = 0; j < n; j++)
            A_(j, m) += yy * A_(j, i);
        }
      }
    }
  }
  // --- eltran: accumulate the similarity transformations
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      Z_(i, j) = i == j ? 1. : 0.;
  for (int mp = n - 2; mp > 0; mp--) {
    for (int k = mp + 1; k < n; k++)
      Z_(k, mp) = A_(k, mp - 1);
    const int i = perm[mp];
    if (i != mp) {
      for (int j = mp; j < n; j++) {
        Z_(mp, j) = Z_(i, j);
        Z_(i, j) = 0.;
      }
      Z_(i, mp) = 1.;
    }
  }
  for (int i = 2; i < n; i++)
    for (int j = 0; j < i - 1; j++)
      A_(i, j) = 0.;

  // --- hqr2: real Schur form with accumulation
  double anorm = 0.;
  for (int i = 0; i < n; i++)
    for (int j = (i > 0 ? i - 1 : 0); j < n; j++)
      anorm += fabs(A_(i, j));
  int nn = n - 1;
  double t = 0., p = 0., q = 0., r = 0., s = 0., w, xx, yy, z = 0.;
  while (nn >= 0) {
    int its = 0, l;
    do {
      for (l = nn; l > 0; l--) {
        s = fabs(A_(l - 1, l - 1)) + fabs(A_(l, l));
        // both diagonal entries negligible against the matrix (zero, or denormal
        // leftovers such as a fully burnt mass fraction ~1e-308): measure the
        // subdiagonal against the norm instead — a backward error of eps ||A||
        if (s <= DBL_EPS * anorm)
          s = anorm;
        if (fabs(A_(l, l - 1)) <= DBL_EPS * s) {
          A_(l, l - 1) = 0.;
          break;
        }
      }
      xx = A_(nn, nn);
      if (l == nn) { // one root
        wr[nn] = A_(nn, nn) = xx + t;
        wi[nn] = 0.;
        nn--;
      } else {
        yy = A_(nn - 1, nn - 1);
        w = A_(nn, nn - 1) * A_(nn - 1, nn);
        if (l == nn - 1) { // two roots
          p = 0.5 * (yy - xx);
          q = p * p + w;
          z = sqrt(fabs(q));
          xx += t;
          A_(nn, nn) = xx;
          A_(nn - 1, nn - 1) = yy + t;
          if (q >= 0.) { // real pair
            z = p + (p >= 0. ? fabs(z) : -fabs(z));
            wr[nn - 1] = wr[nn] = xx + z;
            if (z != 0.)
              wr[nn] = xx - w / z;
            wi[nn - 1] = wi[nn] = 0.;
            xx = A_(nn, nn - 1);
            s = fabs(xx) + fabs(z);
            p = xx / s;
            q = z / s;
            r = sqrt(p * p + q * q);
            p /= r;
            q /= r;
            for (int j = nn - 1; j < n; j++) { // row modification
              z = A_(nn - 1, j);
              A_(nn - 1, j) = q * z + p * A_(nn, j);
              A_(nn, j) = q * A_(nn, j) - p * z;
            }
            for (int i = 0; i <= nn; i++) { // column modification
              z = A_(i, nn - 1);
              A_(i, nn - 1) = q * z + p * A_(i, nn);
              A_(i, nn) = q * A_(i, nn) - p * z;
            }
            for (int i = 0; i < n; i++) { // accumulate
              z = Z_(i, nn - 1);
              Z_(i, nn - 1) = q * z + p * Z_(i, nn);
              Z_(i, nn) = q * Z_(i, nn) - p * z;
            }
          } else { // complex pair
            wr[nn - 1] = wr[nn] = xx + p;
            wi[nn - 1] = z;
            wi[nn] = -z;
          }
          nn -= 2;
        } else { // no root yet: QR step
          if (its >= 60)
            return false;
          if (its == 10 || its == 20) { // exceptional shift
            t += xx;
            for (int i = 0; i <= nn; i++)
              A_(i, i) -= xx;
            s = fabs(A_(nn, nn - 1)) + fabs(A_(nn - 1, nn - 2));
            yy = xx = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          for (m = nn - 2; m >= l; m--) {
            z = A_(m, m);
            r = xx - z;
            s = yy - z;
            p = (r * s - w) / A_(m + 1, m) + A_(m, m + 1);
            q = A_(m + 1, m + 1) - z - r - s;
            r = A_(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            {
              const double rs_ = 1. / s; // (one reciprocal for the three quotients: the slow
              p *= rs_;                  //  path of div.rn.f64 — zero quotients, all over a
              q *= rs_;                  //  sparse matrix — was 5-11 % of the QR kernels)
              r *= rs_;
            }
            if (m == l)
              break;
            double u = fabs(A_(m, m - 1)) * (fabs(q) + fabs(r));
            double v = fabs(p) * (fabs(A_(m - 1, m - 1)) + fabs(z) + fabs(A_(m + 1, m + 1)));
            if (u <= DBL_EPS * v)
              break;
          }
          for (int i = m; i < nn - 1; i++) {
            A_(i + 2, i) = 0.;
            if (i != m)
              A_(i + 2, i - 1) = 0.;
          }
          for (int k = m; k < nn; k++) {
            if (k != m) {
              p = A_(k, k - 1);
              q = A_(k + 1, k - 1);
              r = 0.;
              if (k + 1 != nn)
                r = A_(k + 2, k - 1);
              if ((xx = fabs(p) + fabs(q) + fabs(r)) != 0.) {
                const double rx_ = 1. / xx;
                p *= rx_;
                q *= rx_;
                r *= rx_;
              }
            }
            double sq = sqrt(p * p + q * q + r * r);
            s = p >= 0. ? sq : -sq;
            if (s != 0.) {
              if (k == m) {
                if (l != m)
                  A_(k, k - 1) = -A_(k, k - 1);
              } else
                A_(k, k - 1) = -s * xx;
              p += s;
              {
                const double rs_ = 1. / s, rp_ = 1. / p;
                xx = p * rs_;
                yy = q * rs_;
                z = r * rs_;
                q *= rp_;
                r *= rp_;
              }
              for (int j = k; j < n; j++) { // row modification
                p = A_(k, j) + q * A_(k + 1, j);
                if (k + 1 != nn) {
                  p += r * A_(k + 2, j);
                  A_(k + 2, j) -= p * z;
                }
                A_(k + 1, j) -= p * yy;
                A_(k, j) -= p * xx;
              }
              int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = 0; i <= mmin; i++) { // column modification
                p = xx * A_(i, k) + yy * A_(i, k + 1);
                if (k + 1 != nn) {
                  p += z * A_(i, k + 2);
                  A_(i, k + 2) -= p * r;
                }
                A_(i, k + 1) -= p * q;
                A_(i, k) -= p;
              }
              for (int i = 0; i < n; i++) { // accumulate
                p = xx * Z_(i, k) + yy * Z_(i, k + 1);
                if (k + 1 != nn) {
                  p += z * Z_(i, k + 2);
                  Z_(i, k + 2) -= p * r;
                }
                Z_(i, k + 1) -= p * q;
                Z_(i, k) -= p;
              }
            }
          }
        }
      }
    } while (l + 1 < nn);
  }

  // --- back-substitution: eigenvectors of the quasi-triangular matrix
  if (anorm != 0.) {
    for (nn = n - 1; nn >= 0; nn--) {
      p = wr[nn];
      q = wi[nn];
      const int na = nn - 1;
      if (q == 0.) { // real vector
        int m = nn;
        A_(nn, nn) = 1.;
        for (int i = nn - 1; i >= 0; i--) {
          w = A_(i, i) - p;
          r = 0.;
          for (int j = m; j <= nn; j++)
            r += A_(i, j) * A_(j, nn);
          if (wi[i] < 0.) {
            z = w;
            s = r;
          } else {
            m = i;
            if (wi[i] == 0.) {
              t = w;
              if (t == 0.)
                t = DBL_EPS * anorm;
              A_(i, nn) = -r / t;
            } else { // solve the 2 x 2 block
              xx = A_(i, i + 1);
              yy = A_(i + 1, i);
              q = (wr[i] - p) * (wr[i] - p) + wi[i] * wi[i];
              t = (xx * s - z * r) / q;
              A_(i, nn) = t;
              if (fabs(xx) > fabs(z))
                A_(i + 1, nn) = (-r - w * t) / xx;
              else
                A_(i + 1, nn) = (-s - yy * t) / z;
            }
            t = fabs(A_(i, nn)); // overflow control
            if (DBL_EPS * t * t > 1.)
              for (int j = i; j <= nn; j++)
                A_(j, nn) /= t;
          }
        }
      } else if (q < 0.) { // complex vector, last component chosen imaginary
        int m = na;
        if (fabs(A_(nn, na)) > fabs(A_(na, nn))) {
          A_(na, na) = q / A_(nn, na);
          A_(na, nn) = -(A_(nn, nn) - p) / A_(nn, na);
        } else { // (0, -a[na][nn]) / (a[na][na]-p, q)
          const double cr = A_(na, na) - p, ci = q, nr = 0., ni = -A_(na, nn);
          const double den = cr * cr + ci * ci;
          A_(na, na) = (nr * cr + ni * ci) / den;
          A_(na, nn) = (ni * cr - nr * ci) / den;
        }
        A_(nn, na) = 0.;
        A_(nn, nn) = 1.;
        for (int i = nn - 2; i >= 0; i--) {
          w = A_(i, i) - p;
          double ra = 0., sa = 0.;
          for (int j = m; j <= nn; j++) {
            ra += A_(i, j) * A_(j, na);
            sa += A_(i, j) * A_(j, nn);
          }
          if (wi[i] < 0.) {
            z = w;
            r = ra;
            s = sa;
          } else {
            m = i;
            if (wi[i] == 0.) { // (-ra, -sa) / (w, q)
              const double den = w * w + q * q;
              A_(i, na) = (-ra * w - sa * q) / den;
              A_(i, nn) = (-sa * w + ra * q) / den;
            } else { // solve the complex 2 x 2 block
              xx = A_(i, i + 1);
              yy = A_(i + 1, i);
              double vr = (wr[i] - p) * (wr[i] - p) + wi[i] * wi[i] - q * q;
              const double vi = 2. * q * (wr[i] - p);
              if (vr == 0. && vi == 0.)
                vr = DBL_EPS * anorm * (fabs(w) + fabs(q) + fabs(xx) + fabs(yy) + fabs(z));
              { // (x r - z ra + q sa, x s - z sa - q ra) / (vr, vi)
                const double nr = xx * r - z * ra + q * sa, ni = xx * s - z * sa - q * ra;
                const double den = vr * vr + vi * vi;
                A_(i, na) = (nr * vr + ni * vi) / den;
                A_(i, nn) = (ni * vr - nr * vi) / den;
              }
              if (fabs(xx) > fabs(z) + fabs(q)) {
                A_(i + 1, na) = (-ra - w * A_(i, na) + q * A_(i, nn)) / xx;
                A_(i + 1, nn) = (-sa - w * A_(i, nn) - q * A_(i, na)) / xx;
              } else { // (-r - y a[i][na], -s - y a[i][nn]) / (z, q)
                const double nr = -r - yy * A_(i, na), ni = -s - yy * A_(i, nn);
                const double den = z * z + q * q;
                A_(i + 1, na) = (nr * z + ni * q) / den;
                A_(i + 1, nn) = (ni * z - nr * q) / den;
              }
            }
          }
          t = fmax(fabs(A_(i, na)), fabs(A_(i, nn))); // overflow control
          if (DBL_EPS * t * t > 1.)
            for (int j = i; j <= nn; j++) {
              A_(j, na) /= t;
              A_(j, nn) /= t;
            }
        }
      }
    }
    // multiply by the transformation matrix: vectors of the original matrix
    for (int j = n - 1; j >= 0; j--)
      for (int i = 0; i < n; i++) {
        z = 0.;
        for (int k = 0; k <= j; k++)
          z += Z_(i, k) * A_(k, j);
        Z_(i, j) = z;
      }
  }
  // --- c = W^-1 x by Householder QR with column pivoting on a copy of W (in a),
  // rank-revealing as the reference's colPivHouseholderQr().solve(): for a repeated
  // eigenvalue (v = 0 in gas at rest: a triple zero in 3-D) the back-substituted
  // vectors can be numerically dependent; pivots below n eps |largest pivot| are
  // dropped and their coefficients set to zero (the basic solution).
  double c[n], rhs[n];
  int cp[n];
  for (int i = 0; i < n * n; i++)
    a[i] = zz[i];
  for (int i = 0; i < n; i++) {
    rhs[i] = x[i];
    cp[i] = i;
  }
  int rank = 0;
  double maxpiv = 0.;
  for (int k = 0; k < n; k++) {
    // pivot: remaining column of largest norm (recomputed: n is small)
    int piv = k;
    double best2 = -1.;
    for (int j = k; j < n; j++) {
      double sacc = 0.;
      for (int i = k; i < n; i++)
        sacc += A_(i, j) * A_(i, j);
      if (sacc > best2) {
        best2 = sacc;
        piv = j;
      }
    }
    if (piv != k) {
      for (int i = 0; i < n; i++) {
        double tmp = A_(i, k);
        A_(i, k) = A_(i, piv);
        A_(i, piv) = tmp;
      }
      int ti = cp[k];
      cp[k] = cp[piv];
      cp[piv] = ti;
    }
    const double nrm = sqrt(best2);
    if (k == 0)
      maxpiv = nrm;
    if (!(nrm > n * DBL_EPS * maxpiv))
      break;
    rank = k + 1;
    // Householder vector for column k, rows k..n-1
    const double alpha = A_(k, k) >= 0. ? -nrm : nrm;
    const double v0 = A_(k, k) - alpha;
    double vnorm2 = v0 * v0;
    for (int i = k + 1; i < n; i++)
      vnorm2 += A_(i, k) * A_(i, k);
    if (vnorm2 > 0.) {
      const double beta = 2. / vnorm2;
      for (int j = k + 1; j < n; j++) {
        double sacc = v0 * A_(k, j);
        for (int i = k + 1; i < n; i++)
          sacc += A_(i, k) * A_(i, j);
        sacc *= beta;
        A_(k, j) -= sacc * v0;
        for (int i = k + 1; i < n; i++)
          A_(i, j) -= sacc * A_(i, k);
      }
      double sacc = v0 * rhs[k];
      for (int i = k + 1; i < n; i++)
        sacc += A_(i, k) * rhs[i];
      sacc *= beta;
      rhs[k] -= sacc * v0;
      for (int i = k + 1; i < n; i++)
        rhs[i] -= sacc * A_(i, k);
    }
    A_(k, k) = alpha;
  }
  for (int i = 0; i < n; i++)
    c[i] = 0.;
  for (int i = rank - 1; i >= 0; i--) {
    double sacc = rhs[i];
    for (int j = i + 1; j < rank; j++)
      sacc -= A_(i, j) * c[cp[j]];
    c[cp[i]] = sacc / A_(i, i);
  }
  for (int j = 0; j < n; j++)
    c[j] *= hypot(wr[j], wi[j]);
  for (int i = 0; i < n; i++) {
    double sacc = 0.;
    for (int j = 0; j < n; j++)
      sacc += Z_(i, j) * c[j];
    y[i] = sacc;
  }
  return true;
#undef A_
#undef Z_
}

// The dissipation product as the face kernels call it: the projector form where it certifies,
// the real-Schur route otherwise.  a is destroyed.
template <int n> EIG_FN bool abs_matrix_apply_any(double *a, const double *x, double *y) {
#if PDE_ABS_POLY
  if (n >= 3 && n <= 5) {
    if (abs_matrix_apply_poly<(n >= 3 && n <= 5) ? n : 3>(a, x, y))
      return true;
  }
#endif
  return abs_matrix_apply<n>(a, x, y);
}

#endif // PYPDE_B200_EIG_CUH
