// linalg.cuh — small fixed-size float64 linear algebra for the engine (device + host).  Written from scratch:
// Eigen is what the upstream CPU packages use, it does not exist here, and nothing of it is needed on a GPU.
#pragma once
#include "common.cuh"

namespace b2r {

// Symmetric 3x3 stored as 6 doubles: xx, xy, xz, yy, yz, zz
struct Sym3 {
  double xx, xy, xz, yy, yz, zz;
};

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (A row-major full).  w ascending, V columns = eigenvectors.
B2R_HD void sym_eigen3(const double* Ain, double* w, double* V) {
  double A[9];
#pragma unroll
  for (int i = 0; i < 9; i++) A[i] = Ain[i];
  V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 0; V[4] = 1; V[5] = 0; V[6] = 0; V[7] = 0; V[8] = 1;
  for (int sweep = 0; sweep < 32; sweep++) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; pq++) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      double apq = A[p * 3 + q];
      if (apq == 0.0) continue;
      double app = A[p * 3 + p], aqq = A[q * 3 + q];
      double theta = (aqq - app) / (2.0 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double akp = A[k * 3 + p], akq = A[k * 3 + q];
        A[k * 3 + p] = c * akp - s * akq;
        A[k * 3 + q] = s * akp + c * akq;
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double apk = A[p * 3 + k], aqk = A[q * 3 + k];
        A[p * 3 + k] = c * apk - s * aqk;
        A[q * 3 + k] = s * apk + c * aqk;
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
        V[k * 3 + p] = c * vkp - s * vkq;
        V[k * 3 + q] = s * vkp + c * vkq;
      }
    }
  }
  w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
  // sort ascending (3-element network), permuting columns
#define B2R_SWAPCOL(a, b)                                  \
  if (w[b] < w[a]) {                                       \
    double tw = w[a]; w[a] = w[b]; w[b] = tw;              \
    for (int k = 0; k < 3; k++) { double tv = V[k * 3 + a]; V[k * 3 + a] = V[k * 3 + b]; V[k * 3 + b] = tv; } \
  }
  B2R_SWAPCOL(0, 1)
  B2R_SWAPCOL(1, 2)
  B2R_SWAPCOL(0, 1)
#undef B2R_SWAPCOL
}

// general 3x3 inverse by cofactors; returns determinant
B2R_HD double inv3(const double* m, double* o) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c01 = m[5] * m[6] - m[3] * m[8];
  double c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id;
  o[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id;
  o[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return det;
}

B2R_HD void mul3(const double* a, const double* b, double* o) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) o[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
}

// ---- 6x6 host/device helpers used by the LM / Newton drivers -----------------------------------------------------
// LDL^T solve (fast_gicp: Eigen::LDLT(H + lambda I).solve(-b)).  Returns false on a zero/NaN pivot.
B2R_HD bool ldlt6_solve(const double* A, const double* b, double* x) {
  double L[36], D[6];
  for (int i = 0; i < 36; i++) L[i] = 0.0;
  for (int j = 0; j < 6; j++) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    D[j] = d;
    if (d == 0.0 || d != d) return false;
    L[j * 6 + j] = 1.0;
    for (int i = j + 1; i < 6; i++) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k] * D[k];
      L[i * 6 + j] = s / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * 6 + k] * y[k];
    y[i] = s;
  }
  for (int i = 0; i < 6; i++) y[i] /= D[i];
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * x[k];
    x[i] = s;
  }
  return true;
}

// se3_exp of fast_gicp so3.hpp: a = (omega, v) -> 4x4 row-major isometry
B2R_HD void se3_exp(const double* a, double* D) {
  const double ox = a[0], oy = a[1], oz = a[2];
  double theta_sq = ox * ox + oy * oy + oz * oz;
  double theta = sqrt(theta_sq);
  double imag, real;
  if (theta_sq < 1e-10) {
    double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    double th = sqrt(theta_sq);
    double half = 0.5 * th;
    imag = sin(half) / th;
    real = cos(half);
  }
  double qw = real, qx = imag * ox, qy = imag * oy, qz = imag * oz;
  double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
  double Om[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
  double Om2[9];
  mul3(Om, Om, Om2);
  double V[9];
  if (theta < 1e-10) {
    for (int i = 0; i < 9; i++) V[i] = R[i];
  } else {
    double c1 = (1.0 - cos(theta)) / theta_sq;
    double c2 = (theta - sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * Om[i] + c2 * Om2[i];
  }
  for (int i = 0; i < 16; i++) D[i] = 0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) D[r * 4 + c] = R[r * 3 + c];
    D[r * 4 + 3] = V[r * 3 + 0] * a[3] + V[r * 3 + 1] * a[4] + V[r * 3 + 2] * a[5];
  }
  D[15] = 1.0;
}

// Isometry product O = A * B (affine 3x4 part; last row 0 0 0 1)
B2R_HD void mul_iso(const double* A, const double* B, double* O) {
  double T[16];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) {
      double s = A[r * 4 + 0] * B[0 * 4 + c] + A[r * 4 + 1] * B[1 * 4 + c] + A[r * 4 + 2] * B[2 * 4 + c];
      if (c == 3) s += A[r * 4 + 3];
      T[r * 4 + c] = s;
    }
  T[12] = T[13] = T[14] = 0;
  T[15] = 1;
  for (int i = 0; i < 16; i++) O[i] = T[i];
}

B2R_HD bool gicp_is_converged(const double* D, double rot_eps, double trans_eps) {
  double mr = 0, mt = 0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) {
      double v = 1.0 / rot_eps * fabs(D[r * 4 + c] - (r == c ? 1.0 : 0.0));
      mr = v > mr ? v : mr;
    }
    double v = 1.0 / trans_eps * fabs(D[r * 4 + 3]);
    mt = v > mt ? v : mt;
  }
  return (mr > mt ? mr : mt) < 1;
}

}  // namespace b2r
