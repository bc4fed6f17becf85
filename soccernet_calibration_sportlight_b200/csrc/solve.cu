// solve.cu - batched camera solve: one thread block per frame (CameraCreator.__call__,
// src/models/hrnet/prediction.py:130-136 and everything it dispatches to), the two
// single-camera helpers of the Camera mirror (baseline/camera.py:92-119) and the
// line-intersection keypoints (prediction.py:110-124).  The arithmetic lives in
// solve_core.cuh / solve_cascade.cuh; here are the kernels and the C ABI.
#include <math_constants.h>

#define CAL_TU "solve.cu"
#include "common.cuh"
#include "solve_cascade.cuh"

namespace cal {
namespace {

constexpr int SOLVE_THREADS = 64;    // two warps: the solver is a chain of short team-parallel loops, barrier latency dominates
// Register budget: 6 blocks per SM = at most 168 registers per thread, so that a solve block fits next to
// a persistent 320-thread x 168-register conv CTA (53,760 + 10,752 of the SM's 65,536 registers) and the
// solve of one batch can run under the networks of the next (common.cu: smem_headroom).
constexpr int SOLVE_MIN_BLOCKS = 6;

__global__ void __launch_bounds__(SOLVE_THREADS, SOLVE_MIN_BLOCKS) camera_solve_kernel(const float* __restrict__ preds,
                                                                     const double* __restrict__ line_pts,
                                                                     const __grid_constant__ CalSolveParams P,
                                                                     CalCameraRecord* __restrict__ out) {
  __shared__ solve::Workspace ws;
  const solve::Team T{static_cast<int>(threadIdx.x), static_cast<int>(blockDim.x)};
  solve::Frame fr;
  fr.pred = preds + static_cast<size_t>(blockIdx.x) * solve::NKP * 3;
  fr.line_pts = line_pts ? line_pts + static_cast<size_t>(blockIdx.x) * solve::NKP * 2 : nullptr;
  solve::solve_frame(T, ws, P, fr, out + blockIdx.x);
}

// mode 0: Camera.refine_camera - least squares from (rvec, tvec);
// mode 1: Camera.solve_pnp - the cv2.solvePnPRansac restatement (solve_pnp_cv.cuh); *ok_out = 1 when
// the reference's pose is reproduced, 0 when OpenCV's RANSAC fails on these matches (the reference
// then holds uninitialised memory; the least-squares pose is returned)
__global__ void __launch_bounds__(SOLVE_THREADS) pnp_kernel(const double* __restrict__ obj,
                                                            const double* __restrict__ img, int n,
                                                            const double* __restrict__ K, double* rvec,
                                                            double* tvec, int32_t* ok_out, int mode) {
  __shared__ solve::Workspace ws;
  __shared__ solve::CamState cam;
  const solve::Team T{static_cast<int>(threadIdx.x), static_cast<int>(blockDim.x)};
  if (mode == 0) {
    if (T.tid == 0) {
      ws.nobs = n; ws.nviews = 1; ws.use_f = 0; ws.guard = 1;
      ws.fx = K[0]; ws.fy = K[4]; ws.cx = K[2]; ws.cy = K[5]; ws.f = K[0];
      for (int k = 0; k < n; ++k) {
        solve::Obs& o = ws.obs[k];
        o.X = obj[3 * k]; o.Y = obj[3 * k + 1]; o.Z = obj[3 * k + 2];
        o.u = img[2 * k]; o.v = img[2 * k + 1]; o.w = 1.0; o.view = 0;
      }
      solve::rodrigues_to_R(rvec, ws.pose[0].R);
      for (int k = 0; k < 3; ++k) ws.pose[0].t[k] = tvec[k];
    }
    T.sync();
    solve::lm_solve(T, ws, 100);
    if (T.tid == 0) {
      const bool fin = isfinite(ws.cost);
      if (fin) {
        solve::R_to_rodrigues(ws.pose[0].R, rvec);
        for (int k = 0; k < 3; ++k) tvec[k] = ws.pose[0].t[k];
      }
      if (ok_out) *ok_out = fin ? 1 : 0;
    }
    return;
  }
  if (T.tid == 0) {
    for (int k = 0; k < n; ++k) {
      for (int j = 0; j < 3; ++j) ws.pnp_obj[3 * k + j] = static_cast<double>(static_cast<float>(obj[3 * k + j]));
      for (int j = 0; j < 2; ++j) ws.pnp_px[2 * k + j] = static_cast<double>(static_cast<float>(img[2 * k + j]));
    }
    for (int k = 0; k < 9; ++k) { cam.K[k] = K[k]; cam.R[k] = (k % 4 == 0) ? 1.0 : 0.0; }
    cam.pos[0] = cam.pos[1] = cam.pos[2] = 0.0;
    cam.ok = 1;
  }
  T.sync();
  solve::solve_pnp_core(T, ws, n, &cam, false);
  if (T.tid == 0) {
    if (cam.ok) {
      double t[3];
      solve::mat3_vec(cam.R, cam.pos, t);
      solve::R_to_rodrigues(cam.R, rvec);
      for (int k = 0; k < 3; ++k) tvec[k] = -t[k];
    }
    if (ok_out) *ok_out = cam.ok ? (ws.pnp_status > 0 ? 1 : 0) : -1;
  }
}

// one thread per (frame, keypoint): float64 arithmetic on the float32 peaks, in the reference's
// operation order (under its pinned numpy 1.24 the float32 scalars promote to float64 at the first
// ``x * scale``, export_line_result.py:112-121)
__global__ void line_points_kernel(const float* __restrict__ peaks, int B, int n_lines,
                                   const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b,
                                   float prob_thre, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * solve::NKP) return;
  const int b = idx / solve::NKP, i = idx - b * solve::NKP;
  double ox = CUDART_NAN, oy = CUDART_NAN;
  const int la = pair_a[i], lb = pair_b[i];
  if (la >= 0 && lb >= 0 && la < n_lines && lb < n_lines) {
    double k[2], c[2];
    bool ok = true;
    for (int q = 0; q < 2; ++q) {
      const float* p = peaks + (static_cast<size_t>(b) * n_lines + (q == 0 ? la : lb)) * 6;
      // get_line_data: both peaks must pass the threshold (export_line_result.py:112-126)
      if (!(p[2] >= prob_thre) || !(p[5] >= prob_thre)) { ok = false; break; }
      // calculate_slope_intercept (:51-82): identical points -> (None, None); the reference's
      // CameraCreator.__init__ would then raise on that line pair - here the pair yields no keypoint
      if (p[0] == p[3] && p[1] == p[4]) { ok = false; break; }
      const double x1 = p[0], y1 = p[1], x2 = p[3], y2 = p[4];
      const double slope = (y2 - y1) / (x2 - x1 + 0.00001);
      k[q] = slope;
      c[q] = y1 - slope * x1;
    }
    // line_eq_intersection (prediction.py:643-653)
    if (ok && fabs(k[0] - k[1]) > 1e-4) {
      const double x = (c[1] - c[0]) / (k[0] - k[1]);
      ox = x;
      oy = k[0] * x + c[0];
    }
  }
  out[2 * idx] = ox;
  out[2 * idx + 1] = oy;
}

}  // namespace
}  // namespace cal

extern "C" int cal_camera_solve(const float* preds, const double* line_pts, const CalSolveParams* h_params,
                                int B, CalCameraRecord* out, void* stream) {
  using namespace cal;
  CAL_REQUIRE(B >= 0, CAL_E_INVALID, "cal_camera_solve: B %d", B);
  if (B == 0) return CAL_OK;
  CAL_REQUIRE(preds && h_params && out, CAL_E_INVALID, "cal_camera_solve: null pointer");
  CAL_REQUIRE(h_params->algorithm >= 0 && h_params->algorithm <= 4, CAL_E_INVALID, "cal_camera_solve: algorithm %d",
              h_params->algorithm);
  CAL_REQUIRE(h_params->n_conf_threshs >= 0 && h_params->n_conf_threshs <= 8, CAL_E_INVALID,
              "cal_camera_solve: n_conf_threshs %d", h_params->n_conf_threshs);
  CAL_REQUIRE(h_params->img_w > 0 && h_params->img_h > 0, CAL_E_INVALID, "cal_camera_solve: image size");
  camera_solve_kernel<<<B, SOLVE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(preds, line_pts, *h_params, out);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_pnp_refine(const double* obj, const double* img, int n, const double* K, double* rvec,
                              double* tvec, void* stream) {
  using namespace cal;
  CAL_REQUIRE(obj && img && K && rvec && tvec, CAL_E_INVALID, "cal_pnp_refine: null pointer");
  CAL_REQUIRE(n >= 3 && n <= solve::MAXOBS, CAL_E_INVALID, "cal_pnp_refine: n %d (3..%d)", n, solve::MAXOBS);
  pnp_kernel<<<1, SOLVE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(obj, img, n, K, rvec, tvec, nullptr, 0);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_pnp_solve(const double* obj, const double* img, int n, const double* K, double* rvec,
                             double* tvec, int32_t* ok, void* stream) {
  using namespace cal;
  CAL_REQUIRE(obj && img && K && rvec && tvec, CAL_E_INVALID, "cal_pnp_solve: null pointer");
  CAL_REQUIRE(n >= 4 && n <= solve::NKP, CAL_E_INVALID, "cal_pnp_solve: n %d (4..%d)", n, solve::NKP);
  pnp_kernel<<<1, SOLVE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(obj, img, n, K, rvec, tvec, ok, 1);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_line_points(const float* peaks, int B, int n_lines, const int32_t* pair_a, const int32_t* pair_b,
                               float prob_thre, double* out, void* stream) {
  using namespace cal;
  CAL_REQUIRE(B >= 0 && n_lines >= 1, CAL_E_INVALID, "cal_line_points: bad shape");
  if (B == 0) return CAL_OK;
  CAL_REQUIRE(peaks && pair_a && pair_b && out, CAL_E_INVALID, "cal_line_points: null pointer");
  const int total = B * solve::NKP;
  line_points_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(peaks, B, n_lines, pair_a,
                                                                                       pair_b, prob_thre, out);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}
