// TEST INFRASTRUCTURE: compiles the block-cooperative solver source (csrc/solve_core.cuh,
// solve_cascade.cuh) for the HOST with a one-thread team, so its numerics can be checked
// against OpenCV in the CPU test-suite (no GPU needed).  Never loaded by the product path.
#include "../../soccernet_calibration_sportlight_b200/csrc/solve_cascade.cuh"

extern "C" int host_camera_solve(const float* preds, const double* line_pts, const CalSolveParams* P, int B,
                                 CalCameraRecord* out) {
  using namespace cal::solve;
  Workspace* ws = new Workspace();
  const Team T{0, 1};
  for (int b = 0; b < B; ++b) {
    Frame fr;
    fr.pred = preds + (size_t)b * NKP * 3;
    fr.line_pts = line_pts ? line_pts + (size_t)b * NKP * 2 : nullptr;
    solve_frame(T, *ws, *P, fr, out + b);
  }
  delete ws;
  return 0;
}

extern "C" int host_workspace_bytes(void) { return (int)sizeof(cal::solve::Workspace); }

// Camera.solve_pnp alone (cv2.solvePnPRansac restatement): n matches, K; returns status
// (1 = the reference's pose reproduced, 0 = OpenCV's RANSAC fails here), pose and inlier mask.
extern "C" int host_pnp_ransac(const double* obj, const double* img, int n, const double* K, double* R, double* t,
                               unsigned long long* mask) {
  using namespace cal::solve;
  Workspace* ws = new Workspace();
  const Team T{0, 1};
  for (int k = 0; k < n; ++k) {
    for (int j = 0; j < 3; ++j) ws->pnp_obj[3 * k + j] = (double)(float)obj[3 * k + j];
    for (int j = 0; j < 2; ++j) ws->pnp_px[2 * k + j] = (double)(float)img[2 * k + j];
  }
  CamState cam{};
  for (int k = 0; k < 9; ++k) cam.K[k] = K[k];
  cam.ok = 1;
  solve_pnp_core(T, *ws, n, &cam, false);
  for (int k = 0; k < 9; ++k) R[k] = cam.R[k];
  mat3_vec(cam.R, cam.pos, t);
  for (int k = 0; k < 3; ++k) t[k] = -t[k];
  *mask = ws->pnp_mask;
  const int st = cam.ok ? ws->pnp_status : -1;
  delete ws;
  return st;
}

// Camera.refine_camera alone: least squares over n matches from the pose (R, t); returns the cost
extern "C" double host_pnp_refine(const double* obj, const double* img, int n, const double* K, double* R, double* t,
                                  int max_iter) {
  using namespace cal::solve;
  Workspace* ws = new Workspace();
  const Team T{0, 1};
  ws->nobs = n; ws->nviews = 1; ws->use_f = 0; ws->guard = 1;
  ws->fx = K[0]; ws->fy = K[4]; ws->cx = K[2]; ws->cy = K[5]; ws->f = K[0];
  for (int k = 0; k < n; ++k) {
    Obs& o = ws->obs[k];
    o.X = obj[3 * k]; o.Y = obj[3 * k + 1]; o.Z = obj[3 * k + 2];
    o.u = img[2 * k]; o.v = img[2 * k + 1]; o.w = 1.0; o.view = 0;
  }
  for (int k = 0; k < 9; ++k) ws->pose[0].R[k] = R[k];
  for (int k = 0; k < 3; ++k) ws->pose[0].t[k] = t[k];
  lm_solve(T, *ws, max_iter);
  for (int k = 0; k < 9; ++k) R[k] = ws->pose[0].R[k];
  for (int k = 0; k < 3; ++k) t[k] = ws->pose[0].t[k];
  const double c = ws->cost;
  delete ws;
  return c;
}

// cv2.findHomography(src, dst, RANSAC, thr) alone: n float32 correspondences -> H (h33 = 1), inlier mask
extern "C" int host_homography_ransac(const float* src_xy, const float* dst_xy, int n, double thr, double* H,
                                      unsigned char* inliers) {
  using namespace cal::solve;
  Workspace* ws = new Workspace();
  const Team T{0, 1};
  ws->hn = n;
  for (int k = 0; k < n; ++k) {
    ws->hx[k] = src_xy[2 * k]; ws->hy[k] = src_xy[2 * k + 1];
    ws->hu[k] = dst_xy[2 * k]; ws->hv[k] = dst_xy[2 * k + 1];
  }
  homography_ransac(T, *ws, thr);
  const int ok = ws->flag;
  for (int k = 0; k < 9; ++k) H[k] = ws->H[k];
  for (int k = 0; k < n; ++k) inliers[k] = ws->inl[k];
  delete ws;
  return ok;
}
