// solve_pnp_cv.cuh - what cv2.solvePnPRansac computes for Camera.solve_pnp
// (baseline/camera.py:92-103, reached from src/models/hrnet/prediction.py:369-372, 422-424,
// 509-519, 635-636), restated so that the pose the reference gets is reproduced - not just "a"
// least-squares pose:
//   * points are rounded to float32 first (solvePnPRansac converts CV_64F input);
//   * 4 points  -> P3P on the first three, the solution with the smallest squared reprojection
//     error over all four wins (solveP3P's sort);
//   * 5 points  -> EPnP directly (no RANSAC, no refit);
//   * > 5 points -> RANSAC (RNG seed 2^64-1, 5-point samples without replacement, EPnP on each,
//     8 px inlier threshold, at most 100 iterations shrunk by RANSACUpdateNumIters at 0.99
//     confidence), then the iterative solver on the consensus set; when no sample reaches 5
//     inliers the call fails - the reference then consumes uninitialised memory (DESIGN.md,
//     "unpinned").
// OpenCV's EPnP leans on two properties of cv::SVD (one-sided Jacobi, modules/core/src/lapack.cpp)
// that decide its answer on the coplanar pitch points and are therefore restated here too: the
// left singular vectors of zero singular values are drawn from a fixed pseudo-random sequence
// (RNG 0x12345678) and orthogonalised against the others, and a reflection R = U V^T is "repaired"
// by negating its third ROW.  The algorithms are published (Lepetit/Moreno-Noguer/Fua 2009; Gao et
// al. 2003 for P3P - here any exact P3P root finder gives the same candidates); this file follows
// OpenCV 4.x's calling conventions, not its source text.  Everything here runs per thread on
// thread-local arrays: RANSAC hypotheses are independent, so a team evaluates `nt` of them at once.
#pragma once
#include <float.h>

#include "solve_core.cuh"

namespace cal {
namespace solve {
namespace cvx {

// cv::RNG (multiply-with-carry)
struct CvRng {
  uint64_t s;
  CAL_HD explicit CvRng(uint64_t seed) : s(seed) {}
  CAL_HD uint32_t next() {
    s = (uint64_t)(uint32_t)s * 4164903690ull + (s >> 32);
    return (uint32_t)s;
  }
  CAL_HD int uniform(int a, int b) { return a == b ? a : (int)(next() % (uint32_t)(b - a) + (uint32_t)a); }
};

// One-sided Jacobi SVD of an m x n matrix (m >= n) handed over TRANSPOSED: At holds n rows of m
// values.  On return row i of At is the i-th left singular vector (unit length), W the singular
// values in decreasing order, Vt (n x n, may be null) the right singular vectors as rows.  Rows of
// zero singular values (i < n1) are filled with the +-1/m pseudo-random pattern, orthogonalised
// against the rows before them - only when the singular vectors are asked for (Vt != null).
// want_v = false: the caller needs U only; Vt is then any non-null pointer and is left untouched
// (the rotations of V have no influence on U or W).
CAL_HD_NOINLINE inline void jacobi_svd(double* At, int astep, double* W, double* Vt, int vstep, int m, int n, int n1,
                                       bool want_v = true) {
  const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) { const double t = At[i * astep + k]; sd += t * t; }
    W[i] = sd;
    if (Vt && want_v) {
      for (int k = 0; k < n; ++k) Vt[i * vstep + k] = 0;
      Vt[i * vstep + i] = 1;
    }
  }
  const int max_iter = m > 30 ? m : 30;
  for (int iter = 0; iter < max_iter; ++iter) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double* Ai = At + i * astep;
        double* Aj = At + j * astep;
        double a = W[i], p = 0, b = W[j];
        if (a == 0 || b == 0) continue;      // an all-zero row: p = 0 <= eps * sqrt(a * b) = 0, the pair is never rotated
        for (int k = 0; k < m; ++k) p += Ai[k] * Aj[k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        a = b = 0;
        for (int k = 0; k < m; ++k) {
          const double t0 = c * Ai[k] + s * Aj[k];
          const double t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0; Aj[k] = t1;
          a += t0 * t0; b += t1 * t1;
        }
        W[i] = a; W[j] = b;
        changed = true;
        if (Vt && want_v) {
          double* Vi = Vt + i * vstep;
          double* Vj = Vt + j * vstep;
          for (int k = 0; k < n; ++k) {
            const double t0 = c * Vi[k] + s * Vj[k];
            const double t1 = -s * Vi[k] + c * Vj[k];
            Vi[k] = t0; Vj[k] = t1;
          }
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) { const double t = At[i * astep + k]; sd += t * t; }
    W[i] = sqrt(sd);
  }
  for (int i = 0; i < n - 1; ++i) {
    int j = i;
    for (int k = i + 1; k < n; ++k)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      { const double t = W[i]; W[i] = W[j]; W[j] = t; }
      if (Vt) {
        for (int k = 0; k < m; ++k) { const double t = At[i * astep + k]; At[i * astep + k] = At[j * astep + k]; At[j * astep + k] = t; }
        if (want_v)
          for (int k = 0; k < n; ++k) { const double t = Vt[i * vstep + k]; Vt[i * vstep + k] = Vt[j * vstep + k]; Vt[j * vstep + k] = t; }
      }
    }
  }
  if (!Vt) return;
  CvRng rng(0x12345678ull);
  for (int i = 0; i < n1; ++i) {
    double sd = i < n ? W[i] : 0;
    for (int ii = 0; ii < 100 && sd <= minval; ++ii) {
      const double val0 = 1.0 / m;
      for (int k = 0; k < m; ++k) At[i * astep + k] = (rng.next() & 256) != 0 ? val0 : -val0;
      for (int iter = 0; iter < 2; ++iter)
        for (int j = 0; j < i; ++j) {
          sd = 0;
          for (int k = 0; k < m; ++k) sd += At[i * astep + k] * At[j * astep + k];
          double asum = 0;
          for (int k = 0; k < m; ++k) {
            const double t = At[i * astep + k] - sd * At[j * astep + k];
            At[i * astep + k] = t;
            asum += fabs(t);
          }
          asum = asum > eps * 100 ? 1 / asum : 0;
          for (int k = 0; k < m; ++k) At[i * astep + k] *= asum;
        }
      sd = 0;
      for (int k = 0; k < m; ++k) { const double t = At[i * astep + k]; sd += t * t; }
      sd = sqrt(sd);
    }
    const double s = sd > minval ? 1 / sd : 0.0;
    for (int k = 0; k < m; ++k) At[i * astep + k] *= s;
  }
}

// x = pinv(A) b through the Jacobi SVD (cv::solve(..., DECOMP_SVD)): A is m x n (m >= n <= 5),
// singular values at or below 2 eps sum(w) are dropped.
CAL_HD inline void solve_svd(const double* A, int m, int n, const double* b, double* x) {
  double At[5 * 6], W[5], Vt[5 * 5];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < m; ++k) At[i * m + k] = A[k * n + i];
  jacobi_svd(At, m, W, Vt, n, m, n, n);
  double thr = 0;
  for (int i = 0; i < n; ++i) thr += W[i];
  thr *= DBL_EPSILON * 2;
  for (int j = 0; j < n; ++j) x[j] = 0;
  for (int i = 0; i < n; ++i) {
    double wi = W[i];
    if (fabs(wi) <= thr) continue;
    wi = 1 / wi;
    double s = 0;
    for (int j = 0; j < m; ++j) s += At[i * m + j] * b[j];
    s *= wi;
    for (int j = 0; j < n; ++j) x[j] = x[j] + s * Vt[i * n + j];
  }
}

// SVD of a 3 x 3 matrix A (row-major): U (columns = left vectors), w, V (columns = right vectors)
CAL_HD inline void svd3(const double* A, double* U, double* w, double* V) {
  double At[9], Vt[9];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) At[i * 3 + k] = A[k * 3 + i];
  jacobi_svd(At, 3, w, Vt, 3, 3, 3, 3);
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) { U[k * 3 + i] = At[i * 3 + k]; V[k * 3 + i] = Vt[i * 3 + k]; }
}

struct Epnp5 {
  static constexpr int N = 5;
  double fu, fv, uc, vc;
  double pws[N * 3], us[N * 2], alphas[N * 4], pcs[N * 3];
  double cws[4][3], ccs[4][3];
  double ut[12 * 12];
  double qr_x[4];          // qr_solve leaves X untouched when it meets a zero column

  CAL_HD void choose_control_points() {
    for (int j = 0; j < 3; ++j) cws[0][j] = 0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < 3; ++j) cws[0][j] += pws[3 * i + j];
    for (int j = 0; j < 3; ++j) cws[0][j] /= N;
    double pw0[N * 3], C[9], U[9], dc[3], V[9];
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < 3; ++j) pw0[3 * i + j] = pws[3 * i + j] - cws[0][j];
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < N; ++k) s += pw0[3 * k + i] * pw0[3 * k + j];
        C[i * 3 + j] = s; C[j * 3 + i] = s;
      }
    svd3(C, U, dc, V);
    for (int i = 1; i < 4; ++i) {
      const double k = sqrt(dc[i - 1] / N);
      for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + k * U[j * 3 + (i - 1)];
    }
  }

  CAL_HD void compute_barycentric_coordinates() {
    double cc[9], U[9], w[3], V[9], ci[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
    svd3(cc, U, w, V);
    double thr = (w[0] + w[1] + w[2]) * (DBL_EPSILON * 2);
    for (int k = 0; k < 9; ++k) ci[k] = 0;
    for (int i = 0; i < 3; ++i) {
      if (fabs(w[i]) <= thr) continue;
      const double wi = 1 / w[i];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) ci[r * 3 + c] += (U[c * 3 + i] * wi) * V[r * 3 + i];
    }
    for (int i = 0; i < N; ++i) {
      const double* pi = pws + 3 * i;
      double* a = alphas + 4 * i;
      for (int j = 0; j < 3; ++j)
        a[1 + j] = ci[3 * j] * (pi[0] - cws[0][0]) + ci[3 * j + 1] * (pi[1] - cws[0][1]) + ci[3 * j + 2] * (pi[2] - cws[0][2]);
      a[0] = 1.0 - a[1] - a[2] - a[3];
    }
  }

  CAL_HD void qr_solve(double* A /* 6x4 */, double* b /* 6 */, double* X /* 4 */) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
    for (int k = 0; k < nc; ++k) {
      double eta = fabs(A[k * nc + k]);
      for (int i = k + 1; i < nr; ++i) {               // (scans rows k .. nr-2)
        const double elt = fabs(A[(i - 1) * nc + k]);
        if (eta < elt) eta = elt;
      }
      if (eta == 0) return;
      const double inv_eta = 1.0 / eta;
      double sum2 = 0;
      for (int i = k; i < nr; ++i) { A[i * nc + k] *= inv_eta; sum2 += A[i * nc + k] * A[i * nc + k]; }
      double sigma = sqrt(sum2);
      if (A[k * nc + k] < 0) sigma = -sigma;
      A[k * nc + k] += sigma;
      A1[k] = sigma * A[k * nc + k];
      A2[k] = -eta * sigma;
      for (int j = k + 1; j < nc; ++j) {
        double sum = 0;
        for (int i = k; i < nr; ++i) sum += A[i * nc + k] * A[i * nc + j];
        const double tau = sum / A1[k];
        for (int i = k; i < nr; ++i) A[i * nc + j] -= tau * A[i * nc + k];
      }
    }
    for (int j = 0; j < nc; ++j) {
      double tau = 0;
      for (int i = j; i < nr; ++i) tau += A[i * nc + j] * b[i];
      tau /= A1[j];
      for (int i = j; i < nr; ++i) b[i] -= tau * A[i * nc + j];
    }
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; --i) {
      double sum = 0;
      for (int j = i + 1; j < nc; ++j) sum += A[i * nc + j] * X[j];
      X[i] = (b[i] - sum) / A2[i];
    }
  }

  CAL_HD void gauss_newton(const double* L, const double* rho, double* betas) {
    for (int k = 0; k < 4; ++k) qr_x[k] = 0;
    for (int it = 0; it < 5; ++it) {
      double A[24], b[6];
      for (int i = 0; i < 6; ++i) {
        const double* r = L + i * 10;
        double* ra = A + i * 4;
        ra[0] = 2 * r[0] * betas[0] + r[1] * betas[1] + r[3] * betas[2] + r[6] * betas[3];
        ra[1] = r[1] * betas[0] + 2 * r[2] * betas[1] + r[4] * betas[2] + r[7] * betas[3];
        ra[2] = r[3] * betas[0] + r[4] * betas[1] + 2 * r[5] * betas[2] + r[8] * betas[3];
        ra[3] = r[6] * betas[0] + r[7] * betas[1] + r[8] * betas[2] + 2 * r[9] * betas[3];
        b[i] = rho[i] - (r[0] * betas[0] * betas[0] + r[1] * betas[0] * betas[1] + r[2] * betas[1] * betas[1] +
                         r[3] * betas[0] * betas[2] + r[4] * betas[1] * betas[2] + r[5] * betas[2] * betas[2] +
                         r[6] * betas[0] * betas[3] + r[7] * betas[1] * betas[3] + r[8] * betas[2] * betas[3] +
                         r[9] * betas[3] * betas[3]);
      }
      qr_solve(A, b, qr_x);
      for (int i = 0; i < 4; ++i) betas[i] += qr_x[i];
    }
  }

  CAL_HD double compute_R_and_t(const double* betas, double* R, double* t) {
    for (int i = 0; i < 4; ++i) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double* v = ut + 12 * (11 - i);
      for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 3; ++k) ccs[j][k] += betas[i] * v[3 * j + k];
    }
    for (int i = 0; i < N; ++i) {
      const double* a = alphas + 4 * i;
      for (int j = 0; j < 3; ++j) pcs[3 * i + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
    }
    if (pcs[2] < 0.0) {
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) ccs[i][j] = -ccs[i][j];
      for (int i = 0; i < N * 3; ++i) pcs[i] = -pcs[i];
    }
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < 3; ++j) { pc0[j] += pcs[3 * i + j]; pw0[j] += pws[3 * i + j]; }
    for (int j = 0; j < 3; ++j) { pc0[j] /= N; pw0[j] /= N; }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, U[9], d[3], V[9];
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < 3; ++j) {
        abt[3 * j] += (pcs[3 * i + j] - pc0[j]) * (pws[3 * i] - pw0[0]);
        abt[3 * j + 1] += (pcs[3 * i + j] - pc0[j]) * (pws[3 * i + 1] - pw0[1]);
        abt[3 * j + 2] += (pcs[3 * i + j] - pc0[j]) * (pws[3 * i + 2] - pw0[2]);
      }
    svd3(abt, U, d, V);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i * 3 + j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
    const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] -
                       R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
    if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
    for (int i = 0; i < 3; ++i) t[i] = pc0[i] - (R[i * 3] * pw0[0] + R[i * 3 + 1] * pw0[1] + R[i * 3 + 2] * pw0[2]);
    double sum2 = 0.0;
    for (int i = 0; i < N; ++i) {
      const double* pw = pws + 3 * i;
      const double Xc = R[0] * pw[0] + R[1] * pw[1] + R[2] * pw[2] + t[0];
      const double Yc = R[3] * pw[0] + R[4] * pw[1] + R[5] * pw[2] + t[1];
      const double inv_Zc = 1.0 / (R[6] * pw[0] + R[7] * pw[1] + R[8] * pw[2] + t[2]);
      const double ue = uc + fu * Xc * inv_Zc, ve = vc + fv * Yc * inv_Zc;
      const double u = us[2 * i], v = us[2 * i + 1];
      sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / N;
  }

  // obj: N x 3 (already float32-rounded), xn: N x 2 normalised image points (float32-rounded)
  CAL_HD void compute_pose(const double* obj, const double* xn, const double* K, double* R, double* t) {
    fu = K[0]; fv = K[4]; uc = K[2]; vc = K[5];
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < 3; ++j) pws[3 * i + j] = obj[3 * i + j];
      us[2 * i] = xn[2 * i] * fu + uc;
      us[2 * i + 1] = xn[2 * i + 1] * fv + vc;
    }
    choose_control_points();
    compute_barycentric_coordinates();
    double M[2 * N * 12];
    for (int i = 0; i < N; ++i) {
      const double* as = alphas + 4 * i;
      double* M1 = M + (2 * i) * 12;
      double* M2 = M1 + 12;
      for (int j = 0; j < 4; ++j) {
        M1[3 * j] = as[j] * fu; M1[3 * j + 1] = 0.0; M1[3 * j + 2] = as[j] * (uc - us[2 * i]);
        M2[3 * j] = 0.0; M2[3 * j + 1] = as[j] * fv; M2[3 * j + 2] = as[j] * (vc - us[2 * i + 1]);
      }
    }
    // M^T M (symmetric => its transpose, which the Jacobi routine takes, is itself)
    for (int i = 0; i < 12; ++i)
      for (int j = i; j < 12; ++j) {
        double s = 0;
        for (int k = 0; k < 2 * N; ++k) s += M[k * 12 + i] * M[k * 12 + j];
        ut[i * 12 + j] = s; ut[j * 12 + i] = s;
      }
    {
      double D[12];
      jacobi_svd(ut, 12, D, D /* unused */, 12, 12, 12, 12, false);
    }
    double L[60], rho[6];
    {
      double dv[4][6][3];
      for (int i = 0; i < 4; ++i) {
        const double* v = ut + 12 * (11 - i);
        int a = 0, b = 1;
        for (int j = 0; j < 6; ++j) {
          dv[i][j][0] = v[3 * a] - v[3 * b];
          dv[i][j][1] = v[3 * a + 1] - v[3 * b + 1];
          dv[i][j][2] = v[3 * a + 2] - v[3 * b + 2];
          ++b;
          if (b > 3) { ++a; b = a + 1; }
        }
      }
      auto dot3 = [](const double* p, const double* q) { return p[0] * q[0] + p[1] * q[1] + p[2] * q[2]; };
      for (int i = 0; i < 6; ++i) {
        double* row = L + 10 * i;
        row[0] = dot3(dv[0][i], dv[0][i]);
        row[1] = 2.0 * dot3(dv[0][i], dv[1][i]);
        row[2] = dot3(dv[1][i], dv[1][i]);
        row[3] = 2.0 * dot3(dv[0][i], dv[2][i]);
        row[4] = 2.0 * dot3(dv[1][i], dv[2][i]);
        row[5] = dot3(dv[2][i], dv[2][i]);
        row[6] = 2.0 * dot3(dv[0][i], dv[3][i]);
        row[7] = 2.0 * dot3(dv[1][i], dv[3][i]);
        row[8] = 2.0 * dot3(dv[2][i], dv[3][i]);
        row[9] = dot3(dv[3][i], dv[3][i]);
      }
      auto dist2 = [](const double* p, const double* q) {
        return (p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2]);
      };
      rho[0] = dist2(cws[0], cws[1]); rho[1] = dist2(cws[0], cws[2]); rho[2] = dist2(cws[0], cws[3]);
      rho[3] = dist2(cws[1], cws[2]); rho[4] = dist2(cws[1], cws[3]); rho[5] = dist2(cws[2], cws[3]);
    }
    double Betas[3][4], errs[3], Rs[3][9], ts[3][3];
    {  // betas = [B11 B12 B13 B14]
      double A[24], b4[4];
      for (int i = 0; i < 6; ++i) { A[i * 4] = L[i * 10]; A[i * 4 + 1] = L[i * 10 + 1]; A[i * 4 + 2] = L[i * 10 + 3]; A[i * 4 + 3] = L[i * 10 + 6]; }
      solve_svd(A, 6, 4, rho, b4);
      double* be = Betas[0];
      if (b4[0] < 0) { be[0] = sqrt(-b4[0]); be[1] = -b4[1] / be[0]; be[2] = -b4[2] / be[0]; be[3] = -b4[3] / be[0]; }
      else { be[0] = sqrt(b4[0]); be[1] = b4[1] / be[0]; be[2] = b4[2] / be[0]; be[3] = b4[3] / be[0]; }
    }
    {  // betas = [B11 B12 B22]
      double A[18], b3[3];
      for (int i = 0; i < 6; ++i) { A[i * 3] = L[i * 10]; A[i * 3 + 1] = L[i * 10 + 1]; A[i * 3 + 2] = L[i * 10 + 2]; }
      solve_svd(A, 6, 3, rho, b3);
      double* be = Betas[1];
      if (b3[0] < 0) { be[0] = sqrt(-b3[0]); be[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
      else { be[0] = sqrt(b3[0]); be[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
      if (b3[1] < 0) be[0] = -be[0];
      be[2] = 0.0; be[3] = 0.0;
    }
    {  // betas = [B11 B12 B22 B13 B23]
      double A[30], b5[5];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 5; ++j) A[i * 5 + j] = L[i * 10 + j];
      solve_svd(A, 6, 5, rho, b5);
      double* be = Betas[2];
      if (b5[0] < 0) { be[0] = sqrt(-b5[0]); be[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
      else { be[0] = sqrt(b5[0]); be[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
      if (b5[1] < 0) be[0] = -be[0];
      be[2] = b5[3] / be[0];
      be[3] = 0.0;
    }
    for (int q = 0; q < 3; ++q) {
      gauss_newton(L, rho, Betas[q]);
      errs[q] = compute_R_and_t(Betas[q], Rs[q], ts[q]);
    }
    int Nn = 0;
    if (errs[1] < errs[0]) Nn = 1;
    if (errs[2] < errs[Nn]) Nn = 2;
    for (int k = 0; k < 9; ++k) R[k] = Rs[Nn][k];
    for (int k = 0; k < 3; ++k) t[k] = ts[Nn][k];
  }
};

// undistortPoints on float32 pixels with no distortion: (x - c) * (1/f), stored as float32
CAL_HD inline double normalize_px(double px32, double c, double f) { return (double)(float)((px32 - c) * (1.0 / f)); }

// ---------------------------------------------------------------------------------- P3P
// roots of c[0] + c[1] x + ... + c[deg] x^deg (deg <= 4) by simultaneous (Durand-Kerner)
// iteration in complex arithmetic; real roots are polished by Newton steps
CAL_HD inline int real_roots(const double* c_in, int deg, double* roots) {
  double c[5];
  double mx = 0;
  for (int i = 0; i <= deg; ++i) mx = fmax(mx, fabs(c_in[i]));
  while (deg > 0 && fabs(c_in[deg]) <= 1e-14 * mx) --deg;
  if (deg == 0) return 0;
  for (int i = 0; i <= deg; ++i) c[i] = c_in[i] / c_in[deg];
  double zr[4], zi[4];
  double rad = 0;
  for (int i = 0; i < deg; ++i) rad = fmax(rad, fabs(c[i]));
  rad = 1.0 + rad;
  for (int i = 0; i < deg; ++i) {          // starting points on a circle, not symmetric about the real axis
    const double ang = 0.4 + 6.283185307179586 * i / deg;
    zr[i] = 0.5 * rad * cos(ang); zi[i] = 0.5 * rad * sin(ang);
  }
  for (int it = 0; it < 200; ++it) {
    double change = 0;
    for (int i = 0; i < deg; ++i) {
      double pr = 1.0, pi = 0.0;             // p(z_i), Horner on the monic polynomial
      for (int k = deg - 1; k >= 0; --k) {
        const double nr = pr * zr[i] - pi * zi[i] + c[k], ni = pr * zi[i] + pi * zr[i];
        pr = nr; pi = ni;
      }
      double dr = 1.0, di = 0.0;             // prod (z_i - z_j)
      for (int j = 0; j < deg; ++j) {
        if (j == i) continue;
        const double ar = zr[i] - zr[j], ai = zi[i] - zi[j];
        const double nr = dr * ar - di * ai, ni = dr * ai + di * ar;
        dr = nr; di = ni;
      }
      const double den = dr * dr + di * di;
      if (!(den > 0)) continue;
      const double qr = (pr * dr + pi * di) / den, qi = (pi * dr - pr * di) / den;
      zr[i] -= qr; zi[i] -= qi;
      change = fmax(change, fabs(qr) + fabs(qi));
    }
    if (change < 1e-15 * rad) break;
  }
  int n = 0;
  for (int i = 0; i < deg; ++i) {
    if (!(fabs(zi[i]) < 1e-7 * fmax(1.0, fabs(zr[i])))) continue;
    double x = zr[i];
    for (int it = 0; it < 4; ++it) {
      double p = 1.0, dp = 0.0;
      for (int k = deg - 1; k >= 0; --k) { dp = dp * x + p; p = p * x + c[k]; }
      if (dp != 0 && isfinite(p / dp)) x -= p / dp;
    }
    roots[n++] = x;
  }
  return n;
}

// rigid transform taking three world points onto three camera-frame points (exact for a
// consistent triangle)
CAL_HD inline bool rigid_from_3(const double* Xw, const double* Xc, double* R, double* t) {
  double Fw[9], Fc[9];
  for (int s = 0; s < 2; ++s) {
    const double* Pp = s == 0 ? Xw : Xc;
    double* F = s == 0 ? Fw : Fc;
    double e1[3] = {Pp[3] - Pp[0], Pp[4] - Pp[1], Pp[5] - Pp[2]};
    double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    if (!(n1 > 0)) return false;
    for (int k = 0; k < 3; ++k) e1[k] /= n1;
    const double d[3] = {Pp[6] - Pp[0], Pp[7] - Pp[1], Pp[8] - Pp[2]};
    double e3[3] = {e1[1] * d[2] - e1[2] * d[1], e1[2] * d[0] - e1[0] * d[2], e1[0] * d[1] - e1[1] * d[0]};
    double n3 = sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
    if (!(n3 > 0)) return false;
    for (int k = 0; k < 3; ++k) e3[k] /= n3;
    const double e2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
    for (int k = 0; k < 3; ++k) { F[k * 3] = e1[k]; F[k * 3 + 1] = e2[k]; F[k * 3 + 2] = e3[k]; }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = Fc[i * 3] * Fw[j * 3] + Fc[i * 3 + 1] * Fw[j * 3 + 1] + Fc[i * 3 + 2] * Fw[j * 3 + 2];
  for (int i = 0; i < 3; ++i) t[i] = Xc[i] - (R[i * 3] * Xw[0] + R[i * 3 + 1] * Xw[1] + R[i * 3 + 2] * Xw[2]);
  return true;
}

// solvePnP(P3P) on four points: obj 4 x 3, px 4 x 2 (pixels, float32-rounded).  Returns false
// when the first three points admit no pose.
CAL_HD_NOINLINE inline bool p3p_4points(const double* obj, const double* px, const double* K, double* R, double* t) {
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  double f[3][3];
  for (int i = 0; i < 3; ++i) {
    // solveP3P undistorts with P = K (pixels again, float32), the solver normalises with 1/f and c/f
    const double ux = (double)(float)(normalize_px(px[2 * i], cx, fx) * fx + cx);
    const double uy = (double)(float)(normalize_px(px[2 * i + 1], cy, fy) * fy + cy);
    const double mu = (1.0 / fx) * ux - cx / fx, mv = (1.0 / fy) * uy - cy / fy;
    const double nrm = sqrt(mu * mu + mv * mv + 1);
    f[i][0] = mu / nrm; f[i][1] = mv / nrm; f[i][2] = 1.0 / nrm;
  }
  auto dist = [&](int a, int b) {
    const double dx = obj[3 * a] - obj[3 * b], dy = obj[3 * a + 1] - obj[3 * b + 1], dz = obj[3 * a + 2] - obj[3 * b + 2];
    return sqrt(dx * dx + dy * dy + dz * dz);
  };
  const double a = dist(1, 2), b = dist(0, 2), c = dist(0, 1);
  if (!(a > 0 && b > 0 && c > 0)) return false;
  const double ca = f[1][0] * f[2][0] + f[1][1] * f[2][1] + f[1][2] * f[2][2];
  const double cb = f[0][0] * f[2][0] + f[0][1] * f[2][1] + f[0][2] * f[2][2];
  const double cg = f[0][0] * f[1][0] + f[0][1] * f[1][1] + f[0][2] * f[1][2];
  // depths s2 = u s1, s3 = v s1: u = N(v) / D(v), quartic in v from the third side
  const double k = (a * a - c * c) / (b * b), q = c * c / (b * b);
  const double Np[3] = {1 + k, -2 * k * cb, k - 1}, Dp[2] = {2 * cg, -2 * ca}, Qp[3] = {1, -2 * cb, 1};
  double D2[3] = {Dp[0] * Dp[0], 2 * Dp[0] * Dp[1], Dp[1] * Dp[1]};
  double poly[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 3; ++i) poly[i] += D2[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) poly[i + j] += Np[i] * Np[j] - q * Qp[i] * D2[j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 2; ++j) poly[i + j] += -2 * cg * Np[i] * Dp[j];
  double roots[4];
  const int nr = real_roots(poly, 4, roots);
  bool found = false;
  double best = 0;
  for (int r = 0; r < nr; ++r) {
    const double v = roots[r];
    const double den = 2 * (cg - v * ca);
    if (fabs(den) < 1e-12) continue;
    const double u = ((k - 1) * v * v - 2 * k * cb * v + 1 + k) / den;
    const double w = 1 + v * v - 2 * v * cb;
    if (!(w > 0)) continue;
    const double s1 = b / sqrt(w), s2 = u * s1, s3 = v * s1;
    if (!(s1 > 0 && s2 > 0 && s3 > 0)) continue;
    double Xc[9], Rc[9], tc[3];
    for (int j = 0; j < 3; ++j) { Xc[j] = s1 * f[0][j]; Xc[3 + j] = s2 * f[1][j]; Xc[6 + j] = s3 * f[2][j]; }
    if (!rigid_from_3(obj, Xc, Rc, tc)) continue;
    double e = 0;
    for (int i = 0; i < 4; ++i) {
      const double X = Rc[0] * obj[3 * i] + Rc[1] * obj[3 * i + 1] + Rc[2] * obj[3 * i + 2] + tc[0];
      const double Y = Rc[3] * obj[3 * i] + Rc[4] * obj[3 * i + 1] + Rc[5] * obj[3 * i + 2] + tc[1];
      const double Z = Rc[6] * obj[3 * i] + Rc[7] * obj[3 * i + 1] + Rc[8] * obj[3 * i + 2] + tc[2];
      const double du = X / Z * fx + cx - px[2 * i], dv = Y / Z * fy + cy - px[2 * i + 1];
      e += du * du + dv * dv;
    }
    if (!isfinite(e)) continue;
    if (!found || e < best) {
      found = true; best = e;
      for (int j = 0; j < 9; ++j) R[j] = Rc[j];
      for (int j = 0; j < 3; ++j) t[j] = tc[j];
    }
  }
  return found;
}

// RANSACUpdateNumIters
CAL_HD inline int update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.); p = fmin(p, 1.);
  ep = fmax(ep, 0.); ep = fmin(ep, 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

// squared reprojection errors as PnPRansacCallback::computeError forms them: projections and
// differences in float32; returns the inlier mask (bit i = point i) and count for threshold 8 px
CAL_HD inline int pnp_inliers(const double* obj, const double* px, int n, const double* K, const double* R,
                              const double* t, unsigned long long* mask_out) {
  unsigned long long mask = 0;
  int cnt = 0;
  const float thr = (float)(8.0 * 8.0);
  for (int i = 0; i < n; ++i) {
    const double X = R[0] * obj[3 * i] + R[1] * obj[3 * i + 1] + R[2] * obj[3 * i + 2] + t[0];
    const double Y = R[3] * obj[3 * i] + R[4] * obj[3 * i + 1] + R[5] * obj[3 * i + 2] + t[1];
    double z = R[6] * obj[3 * i] + R[7] * obj[3 * i + 1] + R[8] * obj[3 * i + 2] + t[2];
    z = z ? 1. / z : 1;
    const float pu = (float)(X * z * K[0] + K[2]), pv = (float)(Y * z * K[4] + K[5]);
    const float du = (float)px[2 * i] - pu, dv = (float)px[2 * i + 1] - pv;
    const float e = du * du + dv * dv;
    if (e <= thr) { mask |= 1ull << i; ++cnt; }
  }
  *mask_out = mask;
  return cnt;
}

}  // namespace cvx

// cv2.findHomography(world_xy, img, RANSAC, thr) on ws.hx/hy/hu/hv (ws.hn > 4 float32-rounded
// points; with exactly 4 OpenCV skips the RANSAC) -> ws.H, ws.inl; ws.flag = 1 on success.
// OpenCV's RANSACPointSetRegistrator: RNG seed 2^64-1, 4-point samples without replacement that
// pass HomographyEstimatorCallback::checkSubset (last point not collinear with any earlier pair in
// either image; the four triangles keep or flip their orientation together), squared transfer
// error in float32 against thr^2, the first sample to beat the best count wins, iteration budget
// 2000 shrunk by RANSACUpdateNumIters at 0.995.  Thread 0 draws `nt` samples per round, the team
// fits and scores them, thread 0 replays the sequential bookkeeping.  Then least squares on the
// consensus set (OpenCV: DLT on the inliers + 10 LM iterations).
CAL_HD_NOINLINE inline void homography_ransac(const Team& T, Workspace& ws, double thr) {
  CAL_COUNT(g_ransac);
  const int n = ws.hn;
  unsigned long long* hyp_mask = reinterpret_cast<unsigned long long*>(ws.hyp_err);
  unsigned char* samples = reinterpret_cast<unsigned char*>(&ws.res[0][0]);     // nt x 4 (nt <= 320)
  const int round = T.nt < 320 ? T.nt : 320;
  if (T.tid == 0) {
    ws.flag = 0;
    for (int i = 0; i < n; ++i) ws.inl[i] = 1;
    ws.pnp_niters = 2000; ws.pnp_best = -1; ws.pnp_maxgood = 0; ws.pnp_status = 0;
    ws.pnp_mask = 0;
  }
  T.sync();
  if (n < 4) return;
  if (n > 4) {
    uint64_t rng_state = 0xFFFFFFFFFFFFFFFFull;       // carried by thread 0 across rounds
    const float thr2 = (float)(thr * thr);
    for (int base = 0; base < 2000; base += round) {
      const int budget = ws.pnp_niters;
      T.sync();
      if (base >= budget || ws.pnp_status < 0) break;
      if (T.tid == 0) {
        cvx::CvRng rng(rng_state);
        int drawn = 0;
        for (int h = 0; h < round && base + h < budget; ++h) {
          bool found = false;
          for (int attempt = 0; attempt < 10000 && !found; ++attempt) {
            unsigned char* idx = samples + h * 4;
            for (int i = 0; i < 4; ++i) {
              int v;
              for (;;) {
                v = rng.uniform(0, n);
                bool dup = false;
                for (int j = 0; j < i; ++j) dup = dup || idx[j] == v;
                if (!dup) break;
              }
              idx[i] = (unsigned char)v;
            }
            // checkSubset (points are float32; differences and products in double)
            bool bad = false;
            for (int side = 0; side < 2 && !bad; ++side) {
              const double* px = side == 0 ? ws.hx : ws.hu;
              const double* py = side == 0 ? ws.hy : ws.hv;
              const int i3 = idx[3];
              for (int j = 0; j < 3 && !bad; ++j) {
                const double dx1 = px[idx[j]] - px[i3], dy1 = py[idx[j]] - py[i3];
                for (int k = 0; k < j; ++k) {
                  const double dx2 = px[idx[k]] - px[i3], dy2 = py[idx[k]] - py[i3];
                  if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) { bad = true; break; }
                }
              }
            }
            if (!bad) {
              const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
              int negative = 0;
              for (int q = 0; q < 4; ++q) {
                double det[2];
                for (int side = 0; side < 2; ++side) {
                  const double* px = side == 0 ? ws.hx : ws.hu;
                  const double* py = side == 0 ? ws.hy : ws.hv;
                  const double x0 = px[idx[tt[q][0]]], y0 = py[idx[tt[q][0]]], x1 = px[idx[tt[q][1]]], y1 = py[idx[tt[q][1]]];
                  const double x2 = px[idx[tt[q][2]]], y2 = py[idx[tt[q][2]]];
                  det[side] = x0 * (y1 - y2) - y0 * (x1 - x2) + (x1 * y2 - y1 * x2);
                }
                negative += det[0] * det[1] < 0 ? 1 : 0;
              }
              if (negative != 0 && negative != 4) bad = true;
            }
            found = !bad;
          }
          if (!found) { if (base + h == 0) ws.pnp_status = -1; else ws.pnp_niters = base + h; break; }
          ++drawn;
        }
        rng_state = rng.s;
        ws.iters = drawn;
      }
      T.sync();
      const int drawn = ws.iters;
      if (T.tid < drawn) {
        const unsigned char* idx = samples + T.tid * 4;
        double x[4], y[4], u[4], v[4], H[9];
        for (int k = 0; k < 4; ++k) { x[k] = ws.hx[idx[k]]; y[k] = ws.hy[idx[k]]; u[k] = ws.hu[idx[k]]; v[k] = ws.hv[idx[k]]; }
        int cnt = -1;
        unsigned long long m = 0;
        if (homography_4pt(x, y, u, v, H)) {
          cnt = 0;
          const float Hf[8] = {(float)H[0], (float)H[1], (float)H[2], (float)H[3], (float)H[4], (float)H[5], (float)H[6], (float)H[7]};
          for (int i = 0; i < n; ++i) {
            const float Mx = (float)ws.hx[i], My = (float)ws.hy[i];
            const float ww = 1.f / (Hf[6] * Mx + Hf[7] * My + 1.f);
            const float dx = (Hf[0] * Mx + Hf[1] * My + Hf[2]) * ww - (float)ws.hu[i];
            const float dy = (Hf[3] * Mx + Hf[4] * My + Hf[5]) * ww - (float)ws.hv[i];
            const float e = dx * dx + dy * dy;
            if (e <= thr2) { ++cnt; m |= 1ull << i; }
          }
        }
        ws.hyp_cnt[T.tid] = cnt;
        hyp_mask[T.tid] = m;
      }
      T.sync();
      if (T.tid == 0) {
        for (int q = 0; q < drawn && base + q < ws.pnp_niters; ++q) {
          const int good = ws.hyp_cnt[q];
          if (good > (ws.pnp_maxgood > 3 ? ws.pnp_maxgood : 3)) {
            ws.pnp_best = base + q; ws.pnp_maxgood = good; ws.pnp_mask = hyp_mask[q];
            ws.pnp_niters = cvx::update_num_iters(0.995, (double)(n - good) / n, 4, ws.pnp_niters);
          }
        }
      }
      T.sync();
    }
    const bool have = ws.pnp_best >= 0 && ws.pnp_status >= 0;
    T.sync();
    if (!have) return;
    if (T.tid == 0)
      for (int i = 0; i < n; ++i) ws.inl[i] = (ws.pnp_mask >> i) & 1ull ? 1 : 0;
    T.sync();
  }
  const bool ok = homography_fit(T, ws, ws.inl, ws.H);     // refit on the consensus set
  if (T.tid == 0) ws.flag = ok ? 1 : 0;
  T.sync();
}

}  // namespace solve
}  // namespace cal
