// Host build of the device eigen-solver text (eig.cuh) for CPU unit tests.
#include "../../include/pypde_b200.h"
#include "eig.cuh"
#include <vector>

extern "C" int pypde_b200_host_spectral_radius(const double *A, int n, int qr_only, double *rho,
                                               int *path) {
  if (n < 1 || n > 17 || !A || !rho)
    return 1;
  std::vector<double> a(A, A + (size_t)n * n);
  int pth = 0;
  double r = 0.;
#define CASE(N)                                                                                   \
  case N:                                                                                         \
    if (qr_only == 2) {                                                                           \
      r = spectral_radius_qr<N>(a.data());                                                        \
    } else if (qr_only == 1) {                                                                       \
      if (N > 2)                                                                                  \
        balance<N>(a.data());                                                                     \
      r = spectral_radius_qr<N>(a.data());                                                        \
    } else if (qr_only == 3 || qr_only == 4) {                                                    \
      /* warm start: solve perturbed copies first (mode 3: one, 1e-5 away; mode 4: a second     \
         two 1e-9 away, so that the last solve starts a relative 1e-7 from its roots) */        \
      EigGuess g{0., 0., 0., 0};                                                                  \
      for (int rep = 0; rep < (qr_only == 3 ? 1 : 3); rep++) {                                    \
        std::vector<double> w(a);                                                                 \
        const double eps = rep == 0 ? 1e-5 : rep * 1e-9;                                          \
        for (size_t i = 0; i < w.size(); i++)                                                     \
          w[i] *= 1. + eps * ((int)(i % 7) - 3);                                                  \
        spectral_radius<N>(w.data(), nullptr, &g);                                                \
      }                                                                                           \
      r = spectral_radius<N>(a.data(), &pth, &g);                                                 \
    } else {                                                                                      \
      r = spectral_radius<N>(a.data(), &pth);                                                     \
    }           \
    break;
  switch (n) {
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11)
    CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17)
  }
#undef CASE
  *rho = r;
  if (path)
    *path = (qr_only == 1 || qr_only == 2) ? 0 : pth;
  return 0;
}

extern "C" int pypde_b200_host_abs_matrix_apply(const double *A, int n, const double *x,
                                                double *y) {
  if (n < 1 || n > 17 || !A || !x || !y)
    return 1;
  std::vector<double> a(A, A + (size_t)n * n);
  bool ok = false;
#define CASE(N)                                                                                   \
  case N:                                                                                         \
    ok = abs_matrix_apply<N>(a.data(), x, y);                                                     \
    break;
  switch (n) {
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11)
    CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17)
  }
#undef CASE
  return ok ? 0 : 2;
}

// The projector form of |A| x for n = 3..5 (abs_matrix_apply_poly): 0 certified (y written),
// 2 not certified (the device then takes abs_matrix_apply).
extern "C" int pypde_b200_host_abs_matrix_apply_poly(const double *A, int n, const double *x,
                                                     double *y) {
  if (n < 3 || n > 5 || !A || !x || !y)
    return 1;
  bool ok = false;
  switch (n) {
  case 3:
    ok = abs_matrix_apply_poly<3>(A, x, y);
    break;
  case 4:
    ok = abs_matrix_apply_poly<4>(A, x, y);
    break;
  case 5:
    ok = abs_matrix_apply_poly<5>(A, x, y);
    break;
  }
  return ok ? 0 : 2;
}

// Two matrices through the two-sided polynomial path (what k_faces_fused runs per face
// point); ok[s] = 0 where that path defers to the QR iteration (rho[s] then comes from it).
extern "C" int pypde_b200_host_spectral_radius_pair(const double *A0, const double *A1, int n,
                                                    double *rho, int *ok) {
  if (n < 3 || n > 5 || !A0 || !A1 || !rho || !ok)
    return 1;
  bool k[2] = {false, false};
  EigGuess g0{0., 0., 0., 0}, g1{0., 0., 0., 0};
  switch (n) {
  case 3:
    spectral_radius_poly_pair<3>(A0, A1, rho, k, &g0, &g1);
    spectral_radius_poly_pair<3>(A0, A1, rho, k, &g0, &g1); // second pass: warm
    break;
  case 4:
    spectral_radius_poly_pair<4>(A0, A1, rho, k, &g0, &g1);
    spectral_radius_poly_pair<4>(A0, A1, rho, k, &g0, &g1);
    break;
  case 5:
    spectral_radius_poly_pair<5>(A0, A1, rho, k, &g0, &g1);
    spectral_radius_poly_pair<5>(A0, A1, rho, k, &g0, &g1);
    break;
  }
  const double *A[2] = {A0, A1};
  for (int s = 0; s < 2; s++) {
    ok[s] = k[s];
    if (!k[s]) {
      std::vector<double> a(A[s], A[s] + (size_t)n * n);
      int pth;
      pypde_b200_host_spectral_radius(a.data(), n, 1, &rho[s], &pth);
    }
  }
  return 0;
}
