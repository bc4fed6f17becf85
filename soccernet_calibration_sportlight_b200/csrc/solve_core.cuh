// solve_core.cuh - the per-frame camera solve as block-cooperative code.
//
// Restates, for one frame per thread block, what CameraCreator.__call__ does on the host in
// the reference (src/models/hrnet/prediction.py:130-437, 464-640) on top of Camera
// (baseline/camera.py:92-119, 249-277, 366-426) and of the OpenCV routines those call:
//   cv2.calibrateCamera  (planar views, principal point / aspect / distortion fixed)
//        -> per-view DLT homography, Zhang closed-form focal length, per-view pose from the
//           homography, then Levenberg-Marquardt over [f | pose_1 .. pose_V] to the minimiser;
//   cv2.solvePnPRefineLM -> 6-DoF Levenberg-Marquardt to the minimiser;
//   cv2.solvePnPRansac   -> pose minimising the reprojection error over the matched points
//        (what the reference gets whenever its RANSAC succeeds and refines; when OpenCV's
//        RANSAC fails the reference consumes uninitialised memory - not reproducible, see
//        DESIGN.md);
//   cv2.findHomography(RANSAC) -> OpenCV's seeded 4-point samples and sequential consensus
//        bookkeeping (solve_pnp_cv.cuh), least squares on the consensus set;
//   numpy SVD/Cholesky for K-from-homography (camera.py:366-426) -> one-sided Jacobi.
// All arithmetic is fp64.  `Team` abstracts the thread block: on the device the parallel-for
// loops stride over threadIdx.x and sync() is __syncthreads(); compiled for the host
// (tests/host_solver, one "thread") the same source runs serially so the numerics can be
// validated against OpenCV without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/calib_b200.h"

#if defined(__CUDACC__)
#define CAL_HD __host__ __device__
#define CAL_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CAL_HD
#define CAL_HD_NOINLINE
#endif

#ifndef CAL_COUNT
#define CAL_COUNT(x) ((void)0)     // host profiling hook (tests only)
#endif

namespace cal {
namespace solve {

constexpr int NKP = CAL_NUM_KEYPOINTS;   // 57
constexpr int MAXV = 3;                  // planar views: ground plane, left goal, right goal
constexpr int MAXOBS = 80;               // 53 + 10 + 10 observations at most
constexpr int MAXP = 1 + 6 * MAXV;       // f + 3 poses
#ifndef LM_STEP_TOL
#define LM_STEP_TOL 1e-11
#define LM_COST_TOL 1e-14
#endif
constexpr int NHYP = 512;                // RANSAC hypotheses evaluated per round (scratch size)

struct Team {
  int tid, nt;
  CAL_HD void sync() const {
#if defined(__CUDA_ARCH__)
    if (nt <= 32) __syncwarp();      // a one-warp team: barrier latency is what the solver is made of
    else __syncthreads();
#endif
  }
};

// Team-parallel solve of the damped normal equations M x = rhs (M symmetric positive definite,
// n <= MAXP): Gaussian elimination without pivoting, each elimination step spread over the team
// (the serial version costs ~n^3/3 dependent shared-memory round trips on one thread).
// The solution overwrites rhs; returns false (uniformly) on a non-positive pivot.
CAL_HD inline bool solve_spd_team(const Team& T, double* M, double* rhs, int n, int lda) {
  bool ok = true;
  for (int c = 0; c < n; ++c) {
    const double piv = M[c * lda + c];
    if (!(piv > 1e-300) || !isfinite(piv)) ok = false;          // same value on every thread
    const int w = n - c;                                        // columns c+1..n-1 and the rhs
    const int cnt = (n - c - 1) * w;
    if (ok) {
      const double ip = 1.0 / piv;
      for (int idx = T.tid; idx < cnt; idx += T.nt) {
        const int r = c + 1 + idx / w, kk = idx - (idx / w) * w;
        const double f = M[r * lda + c] * ip;
        if (kk < w - 1) M[r * lda + c + 1 + kk] -= f * M[c * lda + c + 1 + kk];
        else rhs[r] -= f * rhs[c];
      }
    }
    T.sync();
    if (!ok) break;
  }
  if (!ok) return false;
  for (int c = n - 1; c >= 0; --c) {
    if (T.tid == 0) rhs[c] /= M[c * lda + c];
    T.sync();
    const double xc = rhs[c];
    for (int r = T.tid; r < c; r += T.nt) rhs[r] -= M[r * lda + c] * xc;
    T.sync();
  }
  return true;
}

// ---------------------------------------------------------------- static tables
// prediction.py:15-26 (plane sets, keep_points)
CAL_HD inline bool is_top(int i) { return i == 0 || i == 1 || i == 24 || i == 25; }
CAL_HD inline bool in_goal_left(int i) {
  return i == 0 || i == 1 || i == 2 || i == 3 || i == 6 || i == 7 || (i >= 10 && i <= 13);
}
CAL_HD inline bool in_goal_right(int i) {
  return i == 18 || i == 19 || i == 22 || i == 23 || (i >= 24 && i <= 29);
}
CAL_HD inline bool in_keep(int i) {
  return i < 29 || i == 40 || i == 41 || i == 42 || i == 44 || i == 45 || i == 48 || i == 51 || i == 52 || i == 55;
}
CAL_HD inline bool in_plane(int plane, int i) {
  return plane == 0 ? !is_top(i) : (plane == 1 ? in_goal_left(i) : in_goal_right(i));
}

// ---------------------------------------------------------------- small linear algebra
CAL_HD inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
CAL_HD inline void mat3_vec(const double* A, const double* x, double* y) {
  for (int i = 0; i < 3; ++i) y[i] = A[i * 3] * x[0] + A[i * 3 + 1] * x[1] + A[i * 3 + 2] * x[2];
}
CAL_HD inline bool mat3_inv(const double* A, double* B) {
  const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
  if (!(fabs(det) > 0.0) || !isfinite(det)) return false;
  const double id = 1.0 / det;
  B[0] = c0 * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  B[3] = c1 * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  B[6] = c2 * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}

// exp([w]x) (cv2.Rodrigues vector -> matrix)
CAL_HD inline void rodrigues_to_R(const double* w, double* R) {
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (th < 1e-14) {
    R[0] = 1; R[1] = -w[2]; R[2] = w[1]; R[3] = w[2]; R[4] = 1; R[5] = -w[0]; R[6] = -w[1]; R[7] = w[0]; R[8] = 1;
    return;
  }
  const double c = cos(th), s = sin(th), c1 = 1.0 - c, x = w[0] / th, y = w[1] / th, z = w[2] / th;
  R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
  R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
  R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}
// rotation matrix -> Rodrigues vector (cv2.Rodrigues matrix -> vector, R orthonormal)
CAL_HD inline void R_to_rodrigues(const double* R, double* w) {
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double th = acos(c);
  if (s < 1e-5) {
    if (c > 0) { w[0] = w[1] = w[2] = 0.0; return; }
    double t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0);
    t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1.0 : 1.0);
    t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1.0 : 1.0);
    if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
    const double n = th / sqrt(rx * rx + ry * ry + rz * rz);
    w[0] = rx * n; w[1] = ry * n; w[2] = rz * n;
    return;
  }
  const double k = 0.5 * th / s;
  w[0] = rx * k; w[1] = ry * k; w[2] = rz * k;
}
// nearest rotation (polar factor U V^T of a near-orthonormal matrix) by Newton iteration
CAL_HD_NOINLINE inline bool orthonormalize(double* R) {
  for (int it = 0; it < 40; ++it) {
    double Ri[9];
    if (!mat3_inv(R, Ri)) return false;
    double d = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double v = 0.5 * (R[i * 3 + j] + Ri[j * 3 + i]);
        d = fmax(d, fabs(v - R[i * 3 + j]));
        R[i * 3 + j] = v;
      }
    if (d < 1e-15) break;
  }
  return true;
}

// right singular vector of the smallest singular value of a 6 x 6 matrix (one-sided Jacobi:
// orthogonalises the columns of A, accurate to working precision relative to each column)
CAL_HD_NOINLINE inline void null_vector6(double* A /* 6x6 row-major, destroyed */, double* v /* 6 */) {
  const int n = 6;
  double V[36];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double a = 0, b = 0, g = 0;
        for (int k = 0; k < n; ++k) {
          a += A[k * n + p] * A[k * n + p];
          b += A[k * n + q] * A[k * n + q];
          g += A[k * n + p] * A[k * n + q];
        }
        if (g == 0.0 || fabs(g) <= 1e-16 * sqrt(a * b)) continue;
        rotated = true;
        const double zeta = (b - a) / (2.0 * g);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < n; ++k) {
          const double x = A[k * n + p], y = A[k * n + q];
          A[k * n + p] = c * x - s * y;
          A[k * n + q] = s * x + c * y;
          const double vx = V[k * n + p], vy = V[k * n + q];
          V[k * n + p] = c * vx - s * vy;
          V[k * n + q] = s * vx + c * vy;
        }
      }
    if (!rotated) break;
  }
  int best = 0;
  double bn = 1e300;
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int k = 0; k < n; ++k) s += A[k * n + j] * A[k * n + j];
    if (s < bn) { bn = s; best = j; }
  }
  for (int k = 0; k < n; ++k) v[k] = V[k * n + best];
}

// solves the symmetric positive (semi-)definite system A x = b (n <= MAXP) by Gaussian
// elimination with partial pivoting; returns false on a singular pivot
CAL_HD_NOINLINE inline bool solve_linear(double* A, double* b, int n, int lda) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(A[c * lda + c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(A[r * lda + c]) > best) { best = fabs(A[r * lda + c]); piv = r; }
    if (!(best > 1e-300) || !isfinite(best)) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) { const double t = A[c * lda + k]; A[c * lda + k] = A[piv * lda + k]; A[piv * lda + k] = t; }
      const double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    const double inv = 1.0 / A[c * lda + c];
    for (int r = c + 1; r < n; ++r) {
      const double f = A[r * lda + c] * inv;
      if (f == 0.0) continue;
      for (int k = c; k < n; ++k) A[r * lda + k] -= f * A[c * lda + k];
      b[r] -= f * b[c];
    }
  }
  for (int c = n - 1; c >= 0; --c) {
    double s = b[c];
    for (int k = c + 1; k < n; ++k) s -= A[c * lda + k] * b[k];
    b[c] = s / A[c * lda + c];
  }
  return true;
}

// ---------------------------------------------------------------- data
struct Obs {           // one observation of a planar view or of the 3-D point set
  double X, Y, Z;      // object coordinates (plane coordinates have Z = 0)
  double u, v;         // image point
  double w;            // weight (duplicated views of get_camera_all_points)
  int view;
};

struct Pose { double R[9]; double t[3]; };

struct Points {        // the reference's `camera_points` dict, in insertion order
  int n;
  int id[NKP];
  double x[NKP], y[NKP];
  CAL_HD int find(int i) const {
    for (int k = 0; k < n; ++k) if (id[k] == i) return k;
    return -1;
  }
};

struct CamState {      // Camera (baseline/camera.py:79-90) numeric state
  double R[9], pos[3];
  double K[9];         // calibration matrix handed to solve_pnp / refine_camera
  double fx, fy;       // xfocal_length / yfocal_length (project_point, JSON)
  double ppx, ppy;     // principal_point attribute (project_point, JSON)
  int ok;
};

// Workspace of one frame: shared memory on the device.
struct Workspace {
  Obs obs[MAXOBS];
  int nobs, nviews, use_f, guard;
  double f;                    // shared focal length (use_f) ...
  double fx, fy, cx, cy;       // ... or fixed intrinsics
  Pose pose[MAXV];
  Pose cand_pose[MAXV];
  Pose save_pose[MAXV];
  double cand_f, save_f, save_cost;
  double jac[MAXOBS][2][8];    // d residual / d [f, w(3), t(3)]  (homography polish: 8 columns)
  double res[MAXOBS][2];
  double part[MAXOBS];         // per-observation partial costs
  double JtJ[MAXP * MAXP], Jtr[MAXP], M[MAXP * MAXP], delta[MAXP];
  double cost, cand_cost, lambda;
  int flag, iters;
  // homography scratch
  double H[9], Hview[MAXV][9], Hcand[9];
  double hx[NKP], hy[NKP], hu[NKP], hv[NKP];
  int hn, hid[NKP];
  unsigned char inl[NKP];
  int hyp_cnt[NHYP];
  double hyp_err[NHYP];
  // cascade state
  Points pts, sub;
  CamState cam, hom, best;
  double hom_rmse, best_rmse;
  int best_tag, best_set;
  // memo of candidate cameras by point subset (the voter's four subsets often coincide, and the
  // homography camera of voter@0.5 is the one original_voter@0.5 already computed)
  CamState memo_cam[4];
  double memo_rmse[4];
  unsigned long long memo_mask[4];
  int memo_n;
  unsigned long long hom_mask;
  int hom_memo_valid;
  // solvePnPRansac restatement (solve_pnp_cv.cuh): float32-rounded matches, RANSAC state.  The
  // hypotheses' samples, poses, inlier counts and masks live in res / jac / hyp_cnt / hyp_err,
  // which are scratch between two least-squares solves.
  double pnp_obj[NKP * 3], pnp_px[NKP * 2], pnp_pose[12];
  unsigned long long pnp_mask;
  int pnp_niters, pnp_best, pnp_maxgood, pnp_status;
};

// ---------------------------------------------------------------- projection
CAL_HD inline void project_obs(const Obs& o, const Pose& p, double fx, double fy, double cx, double cy,
                               double* Xc, double* uv) {
  Xc[0] = p.R[0] * o.X + p.R[1] * o.Y + p.R[2] * o.Z + p.t[0];
  Xc[1] = p.R[3] * o.X + p.R[4] * o.Y + p.R[5] * o.Z + p.t[1];
  Xc[2] = p.R[6] * o.X + p.R[7] * o.Y + p.R[8] * o.Z + p.t[2];
  const double iz = 1.0 / Xc[2];
  uv[0] = fx * Xc[0] * iz + cx;
  uv[1] = fy * Xc[1] * iz + cy;
}

// weighted sum of squared reprojection errors for (f | poses); result in *out (thread 0 sums)
CAL_HD_NOINLINE inline void eval_cost(const Team& T, Workspace& ws, const Pose* poses, double f, double* out) {
  for (int i = T.tid; i < ws.nobs; i += T.nt) {
    const Obs& o = ws.obs[i];
    double Xc[3], uv[2];
    const double fx = ws.use_f ? f : ws.fx, fy = ws.use_f ? f : ws.fy;
    project_obs(o, poses[o.view], fx, fy, ws.cx, ws.cy, Xc, uv);
    const double du = uv[0] - o.u, dv = uv[1] - o.v;
    // the projective residual is blind to the sign of the depth.  For 3-D point sets
    // (ws.guard) a point at or behind the camera plane is rejected; planar views may cross
    // z = 0 and are mapped back by unmirror_planar_views()
    ws.part[i] = (!ws.guard || Xc[2] > 1e-9) ? o.w * (du * du + dv * dv) : 1e30;
  }
  T.sync();
  if (T.tid == 0) {
    double s = 0.0;
    for (int i = 0; i < ws.nobs; ++i) s += ws.part[i];
    *out = s;
  }
  T.sync();
}

// Levenberg-Marquardt over [f (if use_f) | pose_0 .. pose_{V-1}] on ws.obs, to convergence.
// Rotations are updated on the manifold (R <- exp([dw]x) R); the minimiser does not depend on
// the parameterisation.  Normal equations J^T W J are assembled in parallel (one thread per
// entry), the damped system is solved by thread 0.
CAL_HD_NOINLINE inline void lm_solve(const Team& T, Workspace& ws, int max_iter) {
  const int nf = ws.use_f ? 1 : 0;
  const int P = nf + 6 * ws.nviews;
  CAL_COUNT(g_lm_calls);
  eval_cost(T, ws, ws.pose, ws.f, &ws.cost);
  if (T.tid == 0) { ws.lambda = 1e-3; ws.flag = 0; ws.iters = 0; }
  T.sync();
  for (int it = 0; it < max_iter; ++it) {
    if (!isfinite(ws.cost)) break;
    CAL_COUNT(g_lm_iters);
    // residuals and Jacobian rows
    for (int i = T.tid; i < ws.nobs; i += T.nt) {
      const Obs& o = ws.obs[i];
      const Pose& p = ws.pose[o.view];
      const double fx = ws.use_f ? ws.f : ws.fx, fy = ws.use_f ? ws.f : ws.fy;
      double Xc[3], uv[2];
      project_obs(o, p, fx, fy, ws.cx, ws.cy, Xc, uv);
      const double iz = 1.0 / Xc[2], xn = Xc[0] * iz, yn = Xc[1] * iz;
      const double sw = sqrt(o.w);
      ws.res[i][0] = sw * (uv[0] - o.u);
      ws.res[i][1] = sw * (uv[1] - o.v);
      // d(u,v)/dXc
      const double a00 = fx * iz, a02 = -fx * xn * iz, a11 = fy * iz, a12 = -fy * yn * iz;
      // dXc/dw = -[Xc - t]x = -[q]x with q = R X
      const double q0 = Xc[0] - p.t[0], q1 = Xc[1] - p.t[1], q2 = Xc[2] - p.t[2];
      // -[q]x = [[0, q2, -q1], [-q2, 0, q0], [q1, -q0, 0]]
      double* ju = ws.jac[i][0];
      double* jv = ws.jac[i][1];
      ju[0] = sw * xn; jv[0] = sw * yn;                    // d/df (shared focal)
      ju[1] = sw * (a02 * q1);              jv[1] = sw * (a11 * (-q2) + a12 * q1);
      ju[2] = sw * (a00 * q2 + a02 * (-q0)); jv[2] = sw * (a12 * (-q0));
      ju[3] = sw * (a00 * (-q1));           jv[3] = sw * (a11 * q0);
      ju[4] = sw * a00; jv[4] = 0.0;
      ju[5] = 0.0;      jv[5] = sw * a11;
      ju[6] = sw * a02; jv[6] = sw * a12;
    }
    T.sync();
    // normal equations: param a -> (view, local column): a < nf: f; else view (a-nf)/6, col 1+(a-nf)%6
    for (int e = T.tid; e < P * P + P; e += T.nt) {
      if (e < P * P) {
        const int a = e / P, b = e % P;
        if (b < a) continue;
        const int va = a < nf ? -1 : (a - nf) / 6, ca = a < nf ? 0 : 1 + (a - nf) % 6;
        const int vb = b < nf ? -1 : (b - nf) / 6, cb = b < nf ? 0 : 1 + (b - nf) % 6;
        double s = 0.0;
        if (!(va >= 0 && vb >= 0 && va != vb)) {
          for (int i = 0; i < ws.nobs; ++i) {
            const int vi = ws.obs[i].view;
            if ((va >= 0 && vi != va) || (vb >= 0 && vi != vb)) continue;
            s += ws.jac[i][0][ca] * ws.jac[i][0][cb] + ws.jac[i][1][ca] * ws.jac[i][1][cb];
          }
        }
        ws.JtJ[a * P + b] = s;
        ws.JtJ[b * P + a] = s;
      } else {
        const int a = e - P * P;
        const int va = a < nf ? -1 : (a - nf) / 6, ca = a < nf ? 0 : 1 + (a - nf) % 6;
        double s = 0.0;
        for (int i = 0; i < ws.nobs; ++i) {
          if (va >= 0 && ws.obs[i].view != va) continue;
          s += ws.jac[i][0][ca] * ws.res[i][0] + ws.jac[i][1][ca] * ws.res[i][1];
        }
        ws.Jtr[a] = s;
      }
    }
    T.sync();
    // damped steps until the cost decreases
    bool accepted = false;
    for (int trial = 0; trial < 10; ++trial) {
      CAL_COUNT(g_lm_trials);
      for (int e = T.tid; e < P * P + P; e += T.nt) {
        if (e < P * P) {
          const int a = e / P, b = e - a * P;
          double v = ws.JtJ[e];
          if (a == b) v += ws.lambda * (v > 0 ? v : 1.0);
          ws.M[e] = v;
        } else {
          ws.delta[e - P * P] = -ws.Jtr[e - P * P];
        }
      }
      T.sync();
      const bool solved = solve_spd_team(T, ws.M, ws.delta, P, P);
      if (solved) {
        if (T.tid == 0) {
          ws.cand_f = ws.f + (nf ? ws.delta[0] : 0.0);
          for (int v = 0; v < ws.nviews; ++v) {
            const double* d = ws.delta + nf + 6 * v;
            double dR[9];
            rodrigues_to_R(d, dR);
            mat3_mul(dR, ws.pose[v].R, ws.cand_pose[v].R);
            for (int k = 0; k < 3; ++k) ws.cand_pose[v].t[k] = ws.pose[v].t[k] + d[3 + k];
          }
        }
        T.sync();
        eval_cost(T, ws, ws.cand_pose, ws.cand_f, &ws.cand_cost);
      }
      const bool better = solved && isfinite(ws.cand_cost) && ws.cand_cost <= ws.cost;
      T.sync();
      if (T.tid == 0) {
        if (better) {
          double md = 0.0;   // convergence: relative step
          for (int a = 0; a < P; ++a) {
            double scale = 1.0;
            if (a < nf) scale = fabs(ws.f) + 1.0;
            else if ((a - nf) % 6 >= 3) scale = fabs(ws.pose[(a - nf) / 6].t[(a - nf) % 6 - 3]) + 1.0;
            md = fmax(md, fabs(ws.delta[a]) / scale);
          }
          const double dc = ws.cost - ws.cand_cost;
          ws.f = ws.cand_f;
          for (int v = 0; v < ws.nviews; ++v) ws.pose[v] = ws.cand_pose[v];
          ws.flag = (md < LM_STEP_TOL || dc <= LM_COST_TOL * ws.cost) ? 2 : 1;
          ws.cost = ws.cand_cost;
          ws.lambda = fmax(ws.lambda * 0.1, 1e-15);
        } else {
          ws.lambda = ws.lambda * 10.0;
          ws.flag = 0;
        }
      }
      T.sync();
      if (better) { accepted = true; break; }
      if (ws.lambda > 1e10) break;
    }
    if (T.tid == 0) ws.iters = it + 1;
    if (!accepted || ws.flag == 2) break;
    T.sync();
  }
  T.sync();
  // keep rotations orthonormal after many multiplicative updates
  if (T.tid < ws.nviews) orthonormalize(ws.pose[T.tid].R);
  T.sync();
}

// A planar view (all Z = 0) whose points ended up behind the camera is the exact twin of a
// pose in front of it: R' = -R diag(1,1,-1), t' = -t gives R'X + t' = -(RX + t) on the plane,
// i.e. identical projections with positive depths.
CAL_HD_NOINLINE inline void unmirror_planar_views(const Team& T, Workspace& ws) {
  if (T.tid < ws.nviews) {
    const int v = T.tid;
    Pose& p = ws.pose[v];
    int neg = 0, tot = 0;
    for (int i = 0; i < ws.nobs; ++i) {
      if (ws.obs[i].view != v) continue;
      const Obs& o = ws.obs[i];
      const double z = p.R[6] * o.X + p.R[7] * o.Y + p.R[8] * o.Z + p.t[2];
      neg += (z < 0) ? 1 : 0;
      ++tot;
    }
    if (2 * neg > tot) {
      for (int r = 0; r < 3; ++r) { p.R[r * 3] = -p.R[r * 3]; p.R[r * 3 + 1] = -p.R[r * 3 + 1]; p.t[r] = -p.t[r]; }
    }
  }
  T.sync();
}

// ---------------------------------------------------------------- homography
// Least-squares homography of the points flagged in `use` (all if null), executed by the team:
// Hartley-normalised linear start (h33 = 1 in normalised coordinates, 8x8 normal equations)
// followed by Levenberg-Marquardt on the reprojection error to the minimiser - what
// cv2.findHomography(method 0) converges to.  Result in H (h33 = 1); returns false if degenerate.
CAL_HD_NOINLINE inline bool homography_fit(const Team& T, Workspace& ws, const unsigned char* use, double* H) {
  const int n = ws.hn;
  // ---- normalisation (thread 0; O(n))
  if (T.tid == 0) {
    double cx = 0, cy = 0, cu = 0, cv = 0;
    int m = 0;
    for (int i = 0; i < n; ++i) {
      if (use && !use[i]) continue;
      cx += ws.hx[i]; cy += ws.hy[i]; cu += ws.hu[i]; cv += ws.hv[i];
      ++m;
    }
    bool ok = m >= 4;
    double sx = 0, sy = 0, su = 0, sv = 0;
    if (ok) {
      cx /= m; cy /= m; cu /= m; cv /= m;
      for (int i = 0; i < n; ++i) {
        if (use && !use[i]) continue;
        sx += fabs(ws.hx[i] - cx); sy += fabs(ws.hy[i] - cy);
        su += fabs(ws.hu[i] - cu); sv += fabs(ws.hv[i] - cv);
      }
      ok = (sx > 1e-12) && (sy > 1e-12) && (su > 1e-12) && (sv > 1e-12);
    }
    if (ok) { sx = m / sx; sy = m / sy; su = m / su; sv = m / sv; }
    ws.Hcand[0] = cx; ws.Hcand[1] = cy; ws.Hcand[2] = cu; ws.Hcand[3] = cv;
    ws.Hcand[4] = sx; ws.Hcand[5] = sy; ws.Hcand[6] = su; ws.Hcand[7] = sv;
    ws.flag = ok ? 1 : 0;
  }
  T.sync();
  if (ws.flag == 0) { T.sync(); return false; }
  const double cx = ws.Hcand[0], cy = ws.Hcand[1], cu = ws.Hcand[2], cv = ws.Hcand[3];
  const double sx = ws.Hcand[4], sy = ws.Hcand[5], su = ws.Hcand[6], sv = ws.Hcand[7];
  T.sync();
  // ---- linear start: rows (x y 1 0 0 0 -ux -uy | u), (0 0 0 x y 1 -vx -vy | v) per point
  for (int i = T.tid; i < n; i += T.nt) {
    const bool on = !(use && !use[i]);
    const double x = (ws.hx[i] - cx) * sx, y = (ws.hy[i] - cy) * sy;
    const double u = (ws.hu[i] - cu) * su, v = (ws.hv[i] - cv) * sv;
    double* ju = ws.jac[i][0];
    double* jv = ws.jac[i][1];
    const double z = on ? 1.0 : 0.0;
    ju[0] = z * x; ju[1] = z * y; ju[2] = z; ju[3] = 0; ju[4] = 0; ju[5] = 0; ju[6] = -z * u * x; ju[7] = -z * u * y;
    jv[0] = 0; jv[1] = 0; jv[2] = 0; jv[3] = z * x; jv[4] = z * y; jv[5] = z; jv[6] = -z * v * x; jv[7] = -z * v * y;
    ws.res[i][0] = z * u;
    ws.res[i][1] = z * v;
  }
  T.sync();
  for (int e = T.tid; e < 72; e += T.nt) {
    double s = 0.0;
    if (e < 64) {
      const int a = e >> 3, b = e & 7;
      for (int i = 0; i < n; ++i) s += ws.jac[i][0][a] * ws.jac[i][0][b] + ws.jac[i][1][a] * ws.jac[i][1][b];
      ws.M[e] = s;
    } else {
      const int a = e - 64;
      for (int i = 0; i < n; ++i) s += ws.jac[i][0][a] * ws.res[i][0] + ws.jac[i][1][a] * ws.res[i][1];
      ws.delta[a] = s;
    }
  }
  T.sync();
  CAL_COUNT(g_dlt);
  const bool lin_ok = solve_spd_team(T, ws.M, ws.delta, 8, 8);
  if (T.tid == 0) {
    bool ok = lin_ok;
    if (ok) {
      const double h[9] = {ws.delta[0], ws.delta[1], ws.delta[2], ws.delta[3], ws.delta[4], ws.delta[5],
                           ws.delta[6], ws.delta[7], 1.0};
      // denormalise: H = T_dst^-1 * Hn * T_src
      const double Ts[9] = {sx, 0, -cx * sx, 0, sy, -cy * sy, 0, 0, 1};
      const double Tdi[9] = {1.0 / su, 0, cu, 0, 1.0 / sv, cv, 0, 0, 1};
      double tmp[9], Hd[9];
      mat3_mul(h, Ts, tmp);
      mat3_mul(Tdi, tmp, Hd);
      ok = fabs(Hd[8]) > 1e-300 && isfinite(Hd[8]);
      if (ok) {
        const double s = 1.0 / Hd[8];
        for (int k = 0; k < 9; ++k) H[k] = Hd[k] * s;
        for (int k = 0; k < 9; ++k) ok = ok && isfinite(H[k]);
      }
    }
    ws.flag = ok ? 1 : 0;
    ws.lambda = 1e-3;
  }
  T.sync();
  if (ws.flag == 0) { T.sync(); return false; }
  T.sync();
  // ---- LM polish on the reprojection error (8 parameters, h33 = 1)
  CAL_COUNT(g_polish_calls);
  auto eval = [&](const double* h, double* out) {
    for (int i = T.tid; i < n; i += T.nt) {
      double c = 0.0;
      if (!(use && !use[i])) {
        const double w = h[6] * ws.hx[i] + h[7] * ws.hy[i] + 1.0;
        const double du = (h[0] * ws.hx[i] + h[1] * ws.hy[i] + h[2]) / w - ws.hu[i];
        const double dv = (h[3] * ws.hx[i] + h[4] * ws.hy[i] + h[5]) / w - ws.hv[i];
        c = du * du + dv * dv;
      }
      ws.part[i] = c;
    }
    T.sync();
    if (T.tid == 0) {
      double s = 0.0;
      for (int i = 0; i < n; ++i) s += ws.part[i];
      *out = s;
    }
    T.sync();
  };
  eval(H, &ws.cost);
  for (int it = 0; it < 30; ++it) {
    CAL_COUNT(g_polish_iters);
    for (int i = T.tid; i < n; i += T.nt) {
      const double z = (use && !use[i]) ? 0.0 : 1.0;
      const double x = ws.hx[i], y = ws.hy[i];
      const double w = H[6] * x + H[7] * y + 1.0, iw = z / w;
      const double pu = (H[0] * x + H[1] * y + H[2]) / w, pv = (H[3] * x + H[4] * y + H[5]) / w;
      double* ju = ws.jac[i][0];
      double* jv = ws.jac[i][1];
      ju[0] = x * iw; ju[1] = y * iw; ju[2] = iw; ju[3] = 0; ju[4] = 0; ju[5] = 0; ju[6] = -x * pu * iw; ju[7] = -y * pu * iw;
      jv[0] = 0; jv[1] = 0; jv[2] = 0; jv[3] = x * iw; jv[4] = y * iw; jv[5] = iw; jv[6] = -x * pv * iw; jv[7] = -y * pv * iw;
      ws.res[i][0] = z * (pu - ws.hu[i]);
      ws.res[i][1] = z * (pv - ws.hv[i]);
    }
    T.sync();
    for (int e = T.tid; e < 72; e += T.nt) {
      double s = 0.0;
      if (e < 64) {
        const int a = e >> 3, b = e & 7;
        for (int i = 0; i < n; ++i) s += ws.jac[i][0][a] * ws.jac[i][0][b] + ws.jac[i][1][a] * ws.jac[i][1][b];
        ws.JtJ[e] = s;
      } else {
        const int a = e - 64;
        for (int i = 0; i < n; ++i) s += ws.jac[i][0][a] * ws.res[i][0] + ws.jac[i][1][a] * ws.res[i][1];
        ws.Jtr[a] = s;
      }
    }
    T.sync();
    bool accepted = false;
    for (int trial = 0; trial < 16; ++trial) {
      for (int e = T.tid; e < 72; e += T.nt) {
        if (e < 64) {
          double v = ws.JtJ[e];
          if ((e >> 3) == (e & 7)) v += ws.lambda * (v > 0 ? v : 1.0);
          ws.M[e] = v;
        } else {
          ws.delta[e - 64] = -ws.Jtr[e - 64];
        }
      }
      T.sync();
      const bool solved = solve_spd_team(T, ws.M, ws.delta, 8, 8);
      if (solved) {
        if (T.tid == 0) {
          for (int k = 0; k < 8; ++k) ws.Hcand[k] = H[k] + ws.delta[k];
          ws.Hcand[8] = 1.0;
        }
        T.sync();
        eval(ws.Hcand, &ws.cand_cost);
      }
      const bool better = solved && isfinite(ws.cand_cost) && ws.cand_cost <= ws.cost;
      T.sync();
      if (T.tid == 0) {
        if (better) {
          double md = 0;
          for (int k = 0; k < 8; ++k) md = fmax(md, fabs(ws.delta[k]) / (fabs(H[k]) + 1e-12));
          ws.flag = (md < 1e-11 || ws.cost - ws.cand_cost <= 1e-14 * ws.cost) ? 2 : 1;
          for (int k = 0; k < 9; ++k) H[k] = ws.Hcand[k];
          ws.cost = ws.cand_cost;
          ws.lambda = fmax(ws.lambda * 0.1, 1e-15);
        } else {
          ws.lambda *= 10.0;
          ws.flag = 0;
        }
      }
      T.sync();
      if (better) { accepted = true; break; }
      if (ws.lambda > 1e10) break;
    }
    if (!accepted || ws.flag == 2) break;
    T.sync();
  }
  T.sync();
  return true;
}

// exact homography through 4 correspondences (8x8 linear system, h33 = 1)
CAL_HD_NOINLINE inline bool homography_4pt(const double* x, const double* y, const double* u, const double* v, double* H) {
  double A[64], b[8];
  for (int k = 0; k < 4; ++k) {
    double* r1 = A + (2 * k) * 8;
    double* r2 = A + (2 * k + 1) * 8;
    r1[0] = x[k]; r1[1] = y[k]; r1[2] = 1; r1[3] = 0; r1[4] = 0; r1[5] = 0; r1[6] = -u[k] * x[k]; r1[7] = -u[k] * y[k];
    r2[0] = 0; r2[1] = 0; r2[2] = 0; r2[3] = x[k]; r2[4] = y[k]; r2[5] = 1; r2[6] = -v[k] * x[k]; r2[7] = -v[k] * y[k];
    b[2 * k] = u[k];
    b[2 * k + 1] = v[k];
  }
  // scale-aware singularity guard
  double nrm = 0;
  for (int k = 0; k < 64; ++k) nrm = fmax(nrm, fabs(A[k]));
  if (!solve_linear(A, b, 8, 8)) return false;
  for (int k = 0; k < 8; ++k) {
    if (!isfinite(b[k])) return false;
    H[k] = b[k];
  }
  H[8] = 1.0;
  (void)nrm;
  return true;
}

// ---------------------------------------------------------------- pose from a plane homography
// H maps plane coordinates (X, Y, 1) to pixels; K = diag(fx, fy) with principal point (cx, cy).
// R, t such that x ~ K [r1 r2 t] (X, Y, 1)^T with the plane in front of the camera
// (OpenCV's planar branch of solvePnP / calibrateCamera's extrinsic initialisation).
CAL_HD_NOINLINE inline bool pose_from_homography(const double* H, double fx, double fy, double cx, double cy, Pose* p) {
  double M[9];
  for (int j = 0; j < 3; ++j) {
    M[j] = (H[j] - cx * H[6 + j]) / fx;
    M[3 + j] = (H[3 + j] - cy * H[6 + j]) / fy;
    M[6 + j] = H[6 + j];
  }
  double n1 = sqrt(M[0] * M[0] + M[3] * M[3] + M[6] * M[6]);
  double n2 = sqrt(M[1] * M[1] + M[4] * M[4] + M[7] * M[7]);
  if (!(n1 > 1e-300) || !(n2 > 1e-300)) return false;
  double s = 2.0 / (n1 + n2);
  if (M[8] < 0) { s = -s; n1 = -n1; n2 = -n2; }          // t_z > 0
  const double r1[3] = {M[0] / n1, M[3] / n1, M[6] / n1}, r2[3] = {M[1] / n2, M[4] / n2, M[7] / n2};
  const double r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  p->R[0] = r1[0]; p->R[1] = r2[0]; p->R[2] = r3[0];
  p->R[3] = r1[1]; p->R[4] = r2[1]; p->R[5] = r3[1];
  p->R[6] = r1[2]; p->R[7] = r2[2]; p->R[8] = r3[2];
  if (!orthonormalize(p->R)) return false;
  p->t[0] = M[2] * s; p->t[1] = M[5] * s; p->t[2] = M[8] * s;
  return isfinite(p->t[0]) && isfinite(p->t[1]) && isfinite(p->t[2]);
}

// cvInitIntrinsicParams2D with the principal point fixed at (cx, cy) and aspect ratio 1:
// two vanishing-point constraints per view, least squares for (1/fx^2, 1/fy^2), f = mean.
CAL_HD_NOINLINE inline bool zhang_focal(const double (*Hs)[9], int nviews, double cx, double cy, double* f_out) {
  double AtA[4] = {0, 0, 0, 0}, Atb[2] = {0, 0};
  for (int v = 0; v < nviews; ++v) {
    double H[9];
    for (int k = 0; k < 9; ++k) H[k] = Hs[v][k];
    for (int j = 0; j < 3; ++j) { H[j] -= H[6 + j] * cx; H[3 + j] -= H[6 + j] * cy; }
    double h[3], vv[3], d1[3], d2[3], n[4] = {0, 0, 0, 0};
    for (int j = 0; j < 3; ++j) {
      const double t0 = H[j * 3], t1 = H[j * 3 + 1];
      h[j] = t0; vv[j] = t1; d1[j] = (t0 + t1) * 0.5; d2[j] = (t0 - t1) * 0.5;
      n[0] += t0 * t0; n[1] += t1 * t1; n[2] += d1[j] * d1[j]; n[3] += d2[j] * d2[j];
    }
    for (int j = 0; j < 4; ++j) n[j] = 1.0 / sqrt(n[j]);
    for (int j = 0; j < 3; ++j) { h[j] *= n[0]; vv[j] *= n[1]; d1[j] *= n[2]; d2[j] *= n[3]; }
    const double rows[2][2] = {{h[0] * vv[0], h[1] * vv[1]}, {d1[0] * d2[0], d1[1] * d2[1]}};
    const double rhs[2] = {-h[2] * vv[2], -d1[2] * d2[2]};
    for (int r = 0; r < 2; ++r) {
      AtA[0] += rows[r][0] * rows[r][0]; AtA[1] += rows[r][0] * rows[r][1];
      AtA[3] += rows[r][1] * rows[r][1];
      Atb[0] += rows[r][0] * rhs[r]; Atb[1] += rows[r][1] * rhs[r];
    }
  }
  AtA[2] = AtA[1];
  const double det = AtA[0] * AtA[3] - AtA[1] * AtA[2];
  if (!(fabs(det) > 1e-300)) return false;
  const double x0 = (Atb[0] * AtA[3] - Atb[1] * AtA[1]) / det, x1 = (AtA[0] * Atb[1] - AtA[2] * Atb[0]) / det;
  const double fx = sqrt(fabs(1.0 / x0)), fy = sqrt(fabs(1.0 / x1));
  const double f = (fx + fy) * 0.5;
  *f_out = f;
  return isfinite(f) && f > 0;
}

}  // namespace solve
}  // namespace cal
