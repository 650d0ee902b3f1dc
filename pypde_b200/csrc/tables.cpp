#include "tables.h"

#include <cmath>
#include <cstdio>
#include <stdexcept>

namespace pypde {
namespace {

typedef long double ld;
typedef std::vector<ld> Poly; // ascending powers

ld peval(const Poly &p, ld x) {
  ld r = 0;
  for (int i = (int)p.size() - 1; i >= 0; i--)
    r = r * x + p[i];
  return r;
}

Poly pmul(const Poly &a, const Poly &b) {
  Poly r(a.size() + b.size() - 1, 0.0L);
  for (size_t i = 0; i < a.size(); i++)
    for (size_t j = 0; j < b.size(); j++)
      r[i + j] += a[i] * b[j];
  return r;
}

Poly pdiff(const Poly &a, int times = 1) {
  Poly r = a;
  for (int t = 0; t < times; t++) {
    if (r.size() <= 1)
      return Poly(1, 0.0L);
    Poly s(r.size() - 1);
    for (size_t i = 1; i < r.size(); i++)
      s[i - 1] = r[i] * (ld)i;
    r = s;
  }
  return r;
}

Poly pint(const Poly &a) {
  Poly r(a.size() + 1, 0.0L);
  for (size_t i = 0; i < a.size(); i++)
    r[i + 1] = a[i] / (ld)(i + 1);
  return r;
}

// Legendre P_n and P_n' at x by the three-term recurrence
void legendre(int n, ld x, ld &p, ld &dp) {
  ld p0 = 1, p1 = x;
  if (n == 0) {
    p = 1;
    dp = 0;
    return;
  }
  for (int k = 2; k <= n; k++) {
    ld pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1;
    p1 = pk;
  }
  p = p1;
  dp = n * (x * p1 - p0) / (x * x - 1);
}

// inverse of a small dense matrix by Gauss-Jordan with partial pivoting
std::vector<ld> invert(std::vector<ld> A, int n) {
  std::vector<ld> I(n * n, 0.0L);
  for (int i = 0; i < n; i++)
    I[i * n + i] = 1;
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++)
      if (fabsl(A[r * n + c]) > fabsl(A[piv * n + c]))
        piv = r;
    if (A[piv * n + c] == 0)
      throw std::runtime_error("singular table matrix");
    if (piv != c)
      for (int k = 0; k < n; k++) {
        std::swap(A[c * n + k], A[piv * n + k]);
        std::swap(I[c * n + k], I[piv * n + k]);
      }
    ld d = A[c * n + c];
    for (int k = 0; k < n; k++) {
      A[c * n + k] /= d;
      I[c * n + k] /= d;
    }
    for (int r = 0; r < n; r++) {
      if (r == c)
        continue;
      ld f = A[r * n + c];
      if (f == 0)
        continue;
      for (int k = 0; k < n; k++) {
        A[r * n + k] -= f * A[c * n + k];
        I[r * n + k] -= f * I[c * n + k];
      }
    }
  }
  return I;
}

std::vector<double> to_double(const std::vector<ld> &v) {
  std::vector<double> r(v.size());
  for (size_t i = 0; i < v.size(); i++)
    r[i] = (double)v[i];
  return r;
}

} // namespace

BasisTables make_tables(int N) {
  if (N < 1 || N > 8)
    throw std::runtime_error("order N must be in [1, 8]");
  BasisTables T;
  T.N = N;

  // Gauss-Legendre nodes/weights on [-1,1] (Newton on P_N), mapped to [0,1]
  std::vector<ld> x(N), w(N);
  const ld PI = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < N; i++) {
    ld z = -cosl(PI * (i + 0.75L) / (N + 0.5L)); // ascending order
    for (int it = 0; it < 100; it++) {
      ld p, dp;
      legendre(N, z, p, dp);
      ld dz = p / dp;
      z -= dz;
      if (fabsl(dz) < 1e-19L)
        break;
    }
    ld p, dp;
    legendre(N, z, p, dp);
    x[i] = (z + 1) / 2;
    w[i] = 1 / ((1 - z * z) * dp * dp); // (2/((1-z^2)P'^2)) / 2
  }
  if (N % 2 == 1)
    x[N / 2] = 0.5L;
  // symmetrise exactly as a Gauss rule is
  for (int i = 0; i < N / 2; i++) {
    ld xs = (x[i] + (1 - x[N - 1 - i])) / 2;
    x[i] = xs;
    x[N - 1 - i] = 1 - xs;
    ld ws = (w[i] + w[N - 1 - i]) / 2;
    w[i] = w[N - 1 - i] = ws;
  }

  // Lagrange basis on the nodes
  std::vector<Poly> psi(N);
  for (int i = 0; i < N; i++) {
    Poly p(1, 1.0L);
    for (int j = 0; j < N; j++) {
      if (j == i)
        continue;
      Poly f(2);
      f[0] = -x[j] / (x[i] - x[j]);
      f[1] = 1 / (x[i] - x[j]);
      p = pmul(p, f);
    }
    psi[i] = p;
  }

  std::vector<ld> derv(N * N), endv(2 * N);
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      derv[i * N + j] = peval(pdiff(psi[j]), x[i]);
  for (int j = 0; j < N; j++) {
    endv[j] = peval(psi[j], 0);
    endv[N + j] = peval(psi[j], 1);
  }

  // DG_MAT = DG_END - DG_DER^T
  std::vector<ld> dgend(N * N), dgder(N * N), dgmat(N * N);
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      dgend[i * N + j] = peval(psi[i], 1) * peval(psi[j], 1);
      if (i == j) {
        ld e0 = peval(psi[i], 0), e1 = peval(psi[i], 1);
        dgder[i * N + j] = (e1 * e1 - e0 * e0) / 2;
      } else {
        dgder[i * N + j] = w[i] * peval(pdiff(psi[j]), x[i]);
      }
    }
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      dgmat[i * N + j] = dgend[i * N + j] - dgder[j * N + i];

  // WENO stencil matrices: row i, column j = average of psi_j over the cell
  // at integer offset (first_cell + i) relative to the reconstructed cell
  const int FN2 = (N - 1) / 2;      // floor((N-1)/2)
  const int CN2 = (N - 1 + 1) / 2;  // ceil((N-1)/2)
  const int first[4] = {-(N - 1), 0, -CN2, -FN2};
  std::vector<ld> wm[4];
  for (int s = 0; s < 4; s++) {
    wm[s].assign(N * N, 0.0L);
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) {
        Poly P = pint(psi[j]);
        ld a = (ld)(first[s] + i);
        wm[s][i * N + j] = peval(P, a + 1) - peval(P, a);
      }
  }
  // oscillation indicator
  std::vector<ld> sig(N * N, 0.0L);
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      for (int a = 1; a < N; a++) {
        Poly P = pint(pmul(pdiff(psi[i], a), pdiff(psi[j], a)));
        sig[i * N + j] += peval(P, 1) - peval(P, 0);
      }

  T.nodes = to_double(x);
  T.wghts = to_double(w);
  T.derv = to_double(derv);
  T.endv = to_double(endv);
  T.dgmat = to_double(dgmat);
  T.dginv = to_double(invert(dgmat, N));
  T.sig = to_double(sig);
  for (int s = 0; s < 4; s++) {
    T.wm[s] = to_double(wm[s]);
    T.wminv[s] = to_double(invert(wm[s], N));
  }

  // stencil windows inside the 2N-1 line (weno.cpp:41-59)
  const double LAMS = 1., LAMC = 1e5;
  T.nstencils = 2;
  T.stencil_off[0] = 0;
  T.stencil_lam[0] = LAMS;
  T.stencil_off[1] = N - 1;
  T.stencil_lam[1] = LAMS;
  if (N > 2) {
    T.stencil_off[2] = FN2;
    T.stencil_lam[2] = LAMC;
    T.nstencils = 3;
    if (N % 2 == 0) {
      T.stencil_off[3] = CN2;
      T.stencil_lam[3] = LAMC;
      T.nstencils = 4;
    }
  }
  return T;
}

namespace {
void emit(std::string &s, const char *name, const std::vector<double> &v) {
  char buf[64];
  s += "__constant__ double ";
  s += name;
  snprintf(buf, sizeof buf, "[%zu] = {", v.size());
  s += buf;
  for (size_t i = 0; i < v.size(); i++) {
    snprintf(buf, sizeof buf, "%s%a", i ? ", " : "", v[i]);
    s += buf;
  }
  s += "};\n";
}
} // namespace

std::string tables_cuda_source(const BasisTables &T) {
  std::string s = "// generated by pypde_b200/csrc/tables.cpp\n";
  char buf[160];
  emit(s, "T_NODES", T.nodes);
  emit(s, "T_WGHTS", T.wghts);
  emit(s, "T_DERV", T.derv);
  emit(s, "T_ENDV", T.endv);
  emit(s, "T_DGMAT", T.dgmat);
  emit(s, "T_DGINV", T.dginv);
  emit(s, "T_SIG", T.sig);
  std::vector<double> minv, lam;
  for (int k = 0; k < T.nstencils; k++) {
    minv.insert(minv.end(), T.wminv[k].begin(), T.wminv[k].end());
    lam.push_back(T.stencil_lam[k]);
  }
  emit(s, "T_WMINV", minv);
  emit(s, "T_WLAM", lam);
  snprintf(buf, sizeof buf, "#define PDE_NSTENCILS %d\n", T.nstencils);
  s += buf;
  // compile-time so the stencil windows index registers, not local memory
  s += "#define T_WOFF_INIT {";
  for (int k = 0; k < T.nstencils; k++) {
    snprintf(buf, sizeof buf, "%s%d", k ? ", " : "", T.stencil_off[k]);
    s += buf;
  }
  s += "}\n";
  return s;
}

} // namespace pypde
