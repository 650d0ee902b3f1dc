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
using std::sqrt;
#endif

#define DBL_EPS 2.2204460492503131e-16
#ifndef PDE_EIG_PAIR
#define PDE_EIG_PAIR 0 // 1: iterate the two outer roots in one loop (two dependency chains)
#endif
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
        if (s == 0.)
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
            p /= s;
            q /= s;
            r /= s;
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
                p /= x;
                q /= x;
                r /= x;
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
              x = p / s;
              y = q / s;
              z = r / s;
              q /= p;
              r /= p;
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
  int valid;
};

template <int m> struct PolyRoots {
  // Budan-Fourier certificate: every Taylor coefficient of p at x positive means
  // no real root lies to the right of x.
  static EIG_FN bool right_of_all_roots(const double *c, double x) {
    double t[m + 1];
#pragma unroll
    for (int k = 0; k < m; k++)
      t[k] = c[k];
    t[m] = 1.;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < m; i++) {
#pragma unroll
      for (int k = m - 1; k >= i; k--)
        t[k] = fma(x, t[k + 1], t[k]);
      ok = ok && (t[i] > 0.);
    }
    return ok;
  }

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

  // Largest real root of x^m + c[m-1] x^(m-1) + .. + c[0].  Start: `guess` (> 0
  // means given) nudged right by 1e-3 if certified to be right of every real root,
  // else x_cold, which the caller guarantees to be.  false = fall back to QR.
  static EIG_FN bool rightmost(const double *c, double x_cold, double guess, double &root) {
    double x = x_cold;
    if (guess > 0.) {
      const double xg = guess * (1. + 1e-3);
      if (xg <= x_cold && right_of_all_roots(c, xg))
        x = xg;
    }
    bool newton = false;
    int state = 0;
    for (int it = 0; it < 30 && state == 0; it++)
      step(c, x, newton, state, root);
    return state == 1;
  }

  // Largest root of c and of cm at once: the two iterations are independent, and
  // running them in one loop gives the FP64 pipe two dependency chains to overlap.
  static EIG_FN bool outer_pair(const double *c, const double *cm, double x_cold, double gp,
                                double gm, double &rp, double &rm) {
    double xp = x_cold, xm = x_cold;
    if (gp > 0.) {
      const double xg = gp * (1. + 1e-3);
      if (xg <= x_cold && right_of_all_roots(c, xg))
        xp = xg;
    }
    if (gm > 0.) {
      const double xg = gm * (1. + 1e-3);
      if (xg <= x_cold && right_of_all_roots(cm, xg))
        xm = xg;
    }
    bool np = false, nm = false;
    int sp = 0, sm = 0;
    for (int it = 0; it < 30; it++) {
      if (sp == 0)
        step(c, xp, np, sp, rp);
      if (sm == 0)
        step(cm, xm, nm, sm, rm);
      if ((sp != 0 && sm != 0) || sp < 0 || sm < 0)
        break;
    }
    return sp == 1 && sm == 1;
  }
};

// complex pair of y^2 + b1 y + b0 (disc < 0): |mu + y|^2
EIG_FN double pair_modulus2(double mu, double b1, double b0) {
  const double re = mu - 0.5 * b1;
  return re * re + (b0 - 0.25 * b1 * b1);
}

template <int n>
EIG_FN bool spectral_radius_poly(const double *A, double &rho, EigGuess *guess = nullptr) {
  // shift by the mean eigenvalue
  double mu = 0.;
#pragma unroll
  for (int i = 0; i < n; i++)
    mu += A[i * n + i];
  mu *= (1. / n);
  double B[n * n];
  double R = 0.;
#pragma unroll
  for (int i = 0; i < n; i++) {
    double rs = 0.;
#pragma unroll
    for (int j = 0; j < n; j++) {
      B[i * n + j] = A[i * n + j] - (i == j ? mu : 0.);
      rs += fabs(B[i * n + j]);
    }
    R = fmax(R, rs);
  }
  if (!(R <= 1e150)) // NaN or huge: leave it to the general routine
    return false;
  if (R == 0.) {
    rho = fabs(mu);
    return true;
  }
  // Faddeev-LeVerrier: M_1 = B, c_{n-1} = -tr M_1 (= 0); M_k = B (M_{k-1} + c_{n-k+1} I),
  // c_{n-k} = -tr(M_k)/k
  double c[n];
  c[n - 1] = 0.;
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

  // every root satisfies |y| <= ||B||_inf; for a real spectrum with zero mean also
  // |y| <= sqrt((n-1)/n sum y_i^2) = sqrt(-2 c_{n-2} (n-1)/n) (Laguerre-Samuelson),
  // usually much tighter.  The tighter start is used only under the Budan-Fourier
  // certificate inside rightmost(); x0 is the unconditional one.
  const double x0 = R * (1. + 1e-12);
  double gp = -1., gm = -1.;
  if (guess && guess->valid) {
    gp = guess->yp;
    gm = -guess->ym;
  } else if (c[n - 2] < 0.) {
    gp = gm = sqrt(-2. * c[n - 2] * ((n - 1.) / n));
  }
  // outermost real roots of the undeflated polynomial; the leftmost root of p is
  // minus the rightmost root of (-1)^n p(-y)
  double cm[n];
#pragma unroll
  for (int k = 0; k < n; k++)
    cm[k] = ((n - k) & 1) ? -c[k] : c[k];
  double yp, ym;
#if PDE_EIG_PAIR
  if (!PolyRoots<n>::outer_pair(c, cm, x0, gp, gm, yp, ym))
    return false;
#else
  if (!PolyRoots<n>::rightmost(c, x0, gp, yp))
    return false;
  if (!PolyRoots<n>::rightmost(cm, x0, gm, ym))
    return false;
#endif
  ym = -ym;
  // both searches ending on the same root means a single real root (n odd) or a
  // multiple one
  const bool single = !(yp - ym > 1e-7 * R);
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
      if (!PolyRoots<3>::rightmost(ce, fmin(fb, x0) * (1. + 1e-12), -1., yr))
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
    guess->yp = yp;
    guess->ym = ym;
    guess->valid = 1;
  }
  return true;
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

template <int n>
EIG_FN double spectral_radius(double *a, int *path = nullptr, EigGuess *guess = nullptr) {
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
  if (guess)
    guess->valid = 0;
  balance<n>(a);
  return spectral_radius_qr<n>(a);
}

#endif // PYPDE_B200_EIG_CUH
