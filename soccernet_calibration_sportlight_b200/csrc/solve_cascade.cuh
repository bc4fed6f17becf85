// solve_cascade.cuh - the reference's camera heuristics (CameraCreator's five algorithms and
// the candidate-camera helpers, src/models/hrnet/prediction.py:130-640) for one frame, on top
// of the numerical blocks of solve_core.cuh.  Every function is executed by the whole team;
// decisions are taken on values in the shared workspace after a sync, so all threads follow
// the same path.
#pragma once
#include <string.h>
#include "solve_core.cuh"
#include "solve_pnp_cv.cuh"

namespace cal {
namespace solve {

// branch codes written to CalCameraRecord.branch
enum Branch {
  BR_NONE = 0,
  BR_CALIBRATION = 1,        // opencv_calibration
  BR_MULTIPLANE = 2,         // opencv_calibration_multiplane (+16 when refined)
  BR_OV_CALIBRATION = 3,     // original_voter main branch (+16 when refined)
  BR_OV_CALIBRATION_PNP = 4, // original_voter main branch through solve_pnp (+16 when refined)
  BR_OV_HOMOGRAPHY = 5,      // original_voter fallback to the homography camera
  BR_VOTER_REL = 6, BR_VOTER_ACC = 7, BR_VOTER_ALL = 8, BR_VOTER_GROUND = 9,
  BR_VOTER_HOMOGRAPHY = 10,
  BR_REFINED = 16,
};

struct Frame {
  const float* pred;        // (57,3) x, y, conf
  const double* line_pts;   // (57,2) or null; NaN = absent
};

CAL_HD inline double f32r(double v) { return (double)(float)v; }

CAL_HD inline void plane_xy(const CalSolveParams& P, int plane, int id, double* X, double* Y) {
  // sets_transforms (prediction.py:29-41); inputs to calibrateCamera / findHomography are float32
  const double* w = P.pitch_xyz + 3 * id;
  if (plane == 0) { *X = f32r(w[0]); *Y = f32r(w[1]); }
  else { *X = f32r(w[1]); *Y = f32r(w[2]); }
}

// ---- point selection (prediction.py:178-186, 263-269, 345-354) ------------------------------
CAL_HD inline void select_points(const Team& T, Workspace& ws, const CalSolveParams& P, const Frame& fr,
                                 float thr, bool reliable_gate) {
  if (T.tid == 0) {
    int n_det = 0;
    for (int i = 0; i < NKP; ++i) n_det += (fr.pred[3 * i + 2] > thr) ? 1 : 0;
    Points& p = ws.pts;
    p.n = 0;
    for (int i = 0; i < NKP; ++i) {
      if (fr.pred[3 * i + 2] > thr && (!reliable_gate || n_det < P.reliable_thresh || in_keep(i))) {
        p.id[p.n] = i; p.x[p.n] = (double)fr.pred[3 * i]; p.y[p.n] = (double)fr.pred[3 * i + 1];
        ++p.n;
      }
    }
  }
  T.sync();
}

// line-intersection keypoints merged into the selection; `rule`: 0 multiplane (:187-192),
// 1 voter (:271-278), 2 original_voter (:356-364)
CAL_HD inline void merge_line_points(const Team& T, Workspace& ws, const CalSolveParams& P, const Frame& fr, int rule) {
  if (T.tid == 0 && fr.line_pts) {
    Points& p = ws.pts;
    int n_ground0 = 0;
    for (int k = 0; k < p.n; ++k) n_ground0 += is_top(p.id[k]) ? 0 : 1;
    for (int i = 0; i < NKP; ++i) {
      const double lx = fr.line_pts[2 * i], ly = fr.line_pts[2 * i + 1];
      if (isnan(lx) || isnan(ly) || p.find(i) >= 0 || p.n >= NKP) continue;
      bool add;
      if (rule == 0) add = p.n <= P.min_points;
      else if (rule == 1) {
        int g = 0;
        for (int k = 0; k < p.n; ++k) g += is_top(p.id[k]) ? 0 : 1;
        add = g < P.min_points_per_plane;
      } else {
        add = n_ground0 < P.min_points_per_plane || (0 <= lx && lx <= P.img_w && 0 <= ly && ly <= P.img_h);
      }
      if (add) { p.id[p.n] = i; p.x[p.n] = lx; p.y[p.n] = ly; ++p.n; }
    }
  }
  T.sync();
}

// ---- calibrateCamera on planar views ---------------------------------------------------------
// Builds the views of `pts` into ws.obs.  dup = false: one view per plane with >= min_pts points
// (prediction.py:194-212, 374-394); dup = true: get_camera_all_points' duplicated views
// (:523-546) expressed as weights.  Returns the number of views; first_plane = plane of view 0;
// total_len = sum of the lengths of the (possibly duplicated) lists (get_camera_gen's gate).
CAL_HD_NOINLINE inline int build_views(const Team& T, Workspace& ws, const CalSolveParams& P, const Points& pts, int min_pts,
                              bool dup, int* first_plane, int* total_len, int* n_ground, int n_planes = 3) {
  if (T.tid == 0) {
    int nv = 0, no = 0, tl = 0, fp = -1;
    for (int plane = 0; plane < n_planes; ++plane) {
      int cnt = 0, first_pos = -1, pos = 0, len = 0;
      for (int i = 0; i < 58; ++i) {
        if (!in_plane(plane, i)) continue;
        if (i < NKP && pts.find(i) >= 0) { ++cnt; if (first_pos < 0) first_pos = pos; }
        ++pos; len = pos;
      }
      if (plane == 0) ws.best_set = cnt;   // scratch: ground-plane count
      if (cnt == 0 || cnt < min_pts) continue;
      const double weight = dup ? (double)(len - first_pos) : 1.0;
      for (int i = 0; i < NKP; ++i) {
        if (!in_plane(plane, i)) continue;
        const int k = pts.find(i);
        if (k < 0) continue;
        Obs& o = ws.obs[no++];
        plane_xy(P, plane, i, &o.X, &o.Y);
        o.Z = 0.0;
        o.u = f32r(pts.x[k]); o.v = f32r(pts.y[k]);
        o.w = weight; o.view = nv;
      }
      tl += cnt * (int)weight;
      if (fp < 0) fp = plane;
      ++nv;
    }
    ws.nobs = no; ws.nviews = nv;
    ws.hn = fp; ws.iters = tl;             // scratch hand-over to all threads
  }
  T.sync();
  *first_plane = ws.hn;
  *total_len = ws.iters;
  *n_ground = ws.best_set;
  const int nv = ws.nviews;
  T.sync();
  return nv;
}

// cv2.calibrateCamera(views, size, None, None, FIX_PRINCIPAL_POINT | FIX_ASPECT_RATIO | all
// distortion fixed) on ws.obs: returns false if the initialisation is degenerate (OpenCV would
// raise).  On success ws.f and ws.pose[v] hold the minimiser.
CAL_HD_NOINLINE inline bool calibrate_views(const Team& T, Workspace& ws, const CalSolveParams& P) {
  const double cx = (P.img_w - 1) * 0.5, cy = (P.img_h - 1) * 0.5;    // OpenCV's fixed principal point
  bool hom_ok = true;
  for (int v = 0; v < ws.nviews && hom_ok; ++v) {
    if (T.tid == 0) {
      ws.hn = 0;
      for (int i = 0; i < ws.nobs; ++i) {
        if (ws.obs[i].view != v) continue;
        ws.hx[ws.hn] = ws.obs[i].X; ws.hy[ws.hn] = ws.obs[i].Y;
        ws.hu[ws.hn] = ws.obs[i].u; ws.hv[ws.hn] = ws.obs[i].v;
        ++ws.hn;
      }
    }
    T.sync();
    hom_ok = homography_fit(T, ws, nullptr, ws.Hview[v]);
  }
  if (T.tid == 0) { ws.cx = cx; ws.cy = cy; ws.save_cost = INFINITY; }
  T.sync();
  if (!hom_ok) return false;
  // Closed-form start (cvInitIntrinsicParams2D + per-view extrinsics from the homographies), then
  // the extrinsics alone with that f (cvFindExtrinsicCameraParams2).  With several views a
  // degenerate goal-plane view (collinear goal-line points) can wreck the joint closed-form f, so
  // the start from the first view's own f is tried as well and the lower final cost wins.
  const int n_starts = ws.nviews > 1 ? 2 : 1;
  bool any = false;
  for (int start = 0; start < n_starts; ++start) {
    if (T.tid == 0) {
      double f0 = 0.0;
      bool ok = zhang_focal(ws.Hview, start == 0 ? ws.nviews : 1, cx, cy, &f0);
      for (int v = 0; v < ws.nviews && ok; ++v) ok = pose_from_homography(ws.Hview[v], f0, f0, cx, cy, &ws.pose[v]);
      ws.f = f0; ws.fx = f0; ws.fy = f0;
      ws.use_f = 0; ws.guard = 0;
      ws.flag = ok ? 1 : 0;
#ifdef CAL_SOLVE_DEBUG
      printf("calib start %d ok=%d f0=%g nviews=%d t0=(%g %g %g)\n", start, (int)ok, f0, ws.nviews, ws.pose[0].t[0],
             ws.pose[0].t[1], ws.pose[0].t[2]);
#endif
    }
    T.sync();
    const bool ok = ws.flag != 0;
    T.sync();
    if (!ok) continue;
    lm_solve(T, ws, 12);
    unmirror_planar_views(T, ws);
#ifdef CAL_SOLVE_DEBUG
    if (T.tid == 0) printf("  stage1 cost=%g iters=%d t0=(%g %g %g)\n", ws.cost, ws.iters, ws.pose[0].t[0], ws.pose[0].t[1], ws.pose[0].t[2]);
#endif
    if (T.tid == 0) ws.use_f = 1;          // joint refinement of f and all poses
    T.sync();
    lm_solve(T, ws, 30);
    unmirror_planar_views(T, ws);
#ifdef CAL_SOLVE_DEBUG
    if (T.tid == 0) printf("  stage2 cost=%g iters=%d f=%g t0=(%g %g %g)\n", ws.cost, ws.iters, ws.f, ws.pose[0].t[0], ws.pose[0].t[1], ws.pose[0].t[2]);
#endif
    any = true;
    if (T.tid == 0 && isfinite(ws.cost) && ws.cost < ws.save_cost) {
      ws.save_cost = ws.cost; ws.save_f = ws.f;
      for (int v = 0; v < ws.nviews; ++v) ws.save_pose[v] = ws.pose[v];
    }
    T.sync();
  }
  if (!any) return false;
  const bool have = isfinite(ws.save_cost);
  T.sync();
  if (!have) return false;
  if (T.tid == 0) {
    ws.f = ws.save_f; ws.cost = ws.save_cost;
    for (int v = 0; v < ws.nviews; ++v) ws.pose[v] = ws.save_pose[v];
  }
  T.sync();
  const bool fin = isfinite(ws.cost) && isfinite(ws.f);
  T.sync();
  return fin;
}

// the Camera the reference builds from calibrateCamera's mtx, rvecs[0], tvecs[0]
// (prediction.py:160-168, 228-237, 411-420, 625-633)
CAL_HD inline void camera_from_view0(const Team& T, Workspace& ws, const CalSolveParams& P, CamState* cam) {
  if (T.tid == 0) {
    const Pose& p = ws.pose[0];
    for (int k = 0; k < 9; ++k) cam->R[k] = p.R[k];
    for (int i = 0; i < 3; ++i) cam->pos[i] = -(p.R[i] * p.t[0] + p.R[3 + i] * p.t[1] + p.R[6 + i] * p.t[2]);
    const double K[9] = {ws.f, 0, (P.img_w - 1) * 0.5, 0, ws.f, (P.img_h - 1) * 0.5, 0, 0, 1};
    for (int k = 0; k < 9; ++k) cam->K[k] = K[k];
    cam->fx = ws.f; cam->fy = ws.f;
    cam->ppx = P.img_w / 2.0; cam->ppy = P.img_h / 2.0;
    cam->ok = 1;
  }
  T.sync();
}

// ---- 6-DoF refinement over the matched 3-D points (Camera.refine_camera, camera.py:105-119)
CAL_HD inline void load_matched(const Team& T, Workspace& ws, const CalSolveParams& P, const Points& pts, const CamState& cam) {
  if (T.tid == 0) {
    ws.nobs = pts.n; ws.nviews = 1; ws.use_f = 0; ws.guard = 1;
    ws.fx = cam.K[0]; ws.fy = cam.K[4]; ws.cx = cam.K[2]; ws.cy = cam.K[5]; ws.f = cam.K[0];
    for (int k = 0; k < pts.n; ++k) {
      Obs& o = ws.obs[k];
      const double* w = P.pitch_xyz + 3 * pts.id[k];
      o.X = w[0]; o.Y = w[1]; o.Z = w[2];
      o.u = pts.x[k]; o.v = pts.y[k]; o.w = 1.0; o.view = 0;
    }
    for (int k = 0; k < 9; ++k) ws.pose[0].R[k] = cam.R[k];
    mat3_vec(cam.R, cam.pos, ws.pose[0].t);
    for (int k = 0; k < 3; ++k) ws.pose[0].t[k] = -ws.pose[0].t[k];
  }
  T.sync();
}
CAL_HD inline void store_pose(const Team& T, Workspace& ws, CamState* cam) {
  if (T.tid == 0) {
    const Pose& p = ws.pose[0];
    for (int k = 0; k < 9; ++k) cam->R[k] = p.R[k];
    for (int i = 0; i < 3; ++i) cam->pos[i] = -(p.R[i] * p.t[0] + p.R[3 + i] * p.t[1] + p.R[6 + i] * p.t[2]);
    if (!(isfinite(ws.cost) && isfinite(cam->pos[0]) && isfinite(cam->pos[1]) && isfinite(cam->pos[2]))) cam->ok = 0;
  }
  T.sync();
}
CAL_HD_NOINLINE inline void refine_camera(const Team& T, Workspace& ws, const CalSolveParams& P, const Points& pts, CamState* cam) {
  load_matched(T, ws, P, pts, *cam);
  lm_solve(T, ws, 30);
  store_pose(T, ws, cam);
}

// 6-DoF least squares over the matches flagged in `mask` (bit k = pts entry k), from the pose in *cam
CAL_HD_NOINLINE inline void refine_camera_masked(const Team& T, Workspace& ws, int n, unsigned long long mask, CamState* cam) {
  if (T.tid == 0) {
    ws.nviews = 1; ws.use_f = 0; ws.guard = 1;
    ws.fx = cam->K[0]; ws.fy = cam->K[4]; ws.cx = cam->K[2]; ws.cy = cam->K[5]; ws.f = cam->K[0];
    int no = 0;
    for (int k = 0; k < n; ++k) {
      if (!((mask >> k) & 1ull)) continue;
      Obs& o = ws.obs[no++];
      o.X = ws.pnp_obj[3 * k]; o.Y = ws.pnp_obj[3 * k + 1]; o.Z = ws.pnp_obj[3 * k + 2];
      o.u = ws.pnp_px[2 * k]; o.v = ws.pnp_px[2 * k + 1]; o.w = 1.0; o.view = 0;
    }
    ws.nobs = no;
    for (int k = 0; k < 9; ++k) ws.pose[0].R[k] = cam->R[k];
    mat3_vec(cam->R, cam->pos, ws.pose[0].t);
    for (int k = 0; k < 3; ++k) ws.pose[0].t[k] = -ws.pose[0].t[k];
  }
  T.sync();
  lm_solve(T, ws, 30);
  store_pose(T, ws, cam);
}

CAL_HD inline void set_pose(CamState* cam, const double* R, const double* t) {
  for (int k = 0; k < 9; ++k) cam->R[k] = R[k];
  for (int i = 0; i < 3; ++i) cam->pos[i] = -(R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2]);
  bool fin = true;
  for (int k = 0; k < 9; ++k) fin = fin && isfinite(R[k]);
  for (int k = 0; k < 3; ++k) fin = fin && isfinite(cam->pos[k]);
  if (!fin) cam->ok = 0;
}

// Camera.solve_pnp (camera.py:92-103) = cv2.solvePnPRansac(obj, img, K, None) with the result flag
// ignored.  Restated in solve_pnp_cv.cuh; here the team-level driver:
//   4 points -> P3P, 5 points -> EPnP (both by thread 0);
//   more -> RANSAC: thread 0 draws every 5-point sample of the (at most 100) iterations from
//   OpenCV's seeded generator, the team evaluates `nt` hypotheses at a time (EPnP + inlier count,
//   one per thread), thread 0 replays OpenCV's sequential bookkeeping (best count, shrinking
//   iteration budget) over them; then the iterative solver on the consensus set = least squares
//   over the inliers from the winning sample's pose.
// When the RANSAC fails (no sample reaches 5 inliers) the reference goes on with uninitialised
// memory; nothing to reproduce - the least-squares pose over all matches is returned instead
// (from the pose in *cam when `keep_init`, else from the ground-plane homography of the matches).
// ws.pnp_status: 1 reproduced the reference's pose, 0 the RANSAC failed (fallback pose).
// (the n matches are in ws.pnp_obj / ws.pnp_px, float32-rounded; K and the start pose in *cam)
CAL_HD_NOINLINE inline void solve_pnp_core(const Team& T, Workspace& ws, int n, CamState* cam, bool keep_init,
                                           bool refine_follows = false) {
  if (T.tid == 0) {
    ws.pnp_status = 0; ws.pnp_best = -1; ws.pnp_maxgood = 0; ws.pnp_niters = 100;
    ws.pnp_mask = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
  }
  T.sync();
  double* hyp_pose = &ws.jac[0][0][0];                                   // 100 x 12
  unsigned long long* hyp_mask = reinterpret_cast<unsigned long long*>(ws.hyp_err);
  unsigned char* samples = reinterpret_cast<unsigned char*>(&ws.res[0][0]);   // 100 x 5
  if (n == 4 || n == 5) {
    if (T.tid == 0) {
      double R[9], t[3];
      bool ok;
      if (n == 4) {
        ok = cvx::p3p_4points(ws.pnp_obj, ws.pnp_px, cam->K, R, t);
      } else {
        double xn[10];
        for (int k = 0; k < 5; ++k) {
          xn[2 * k] = cvx::normalize_px(ws.pnp_px[2 * k], cam->K[2], cam->K[0]);
          xn[2 * k + 1] = cvx::normalize_px(ws.pnp_px[2 * k + 1], cam->K[5], cam->K[4]);
        }
        cvx::Epnp5 e;
        e.compute_pose(ws.pnp_obj, xn, cam->K, R, t);
        ok = true;
      }
      if (ok) { set_pose(cam, R, t); ws.pnp_status = 1; }
    }
    T.sync();
  } else if (n >= 6 && n <= NKP) {
    // Samples that draw the same five points (in any order) give the same pose up to rounding: each
    // distinct point set is evaluated once (with 6..8 matches there are only 6..56 of them, one round of the
    // team instead of two), every iteration of OpenCV's loop then looks its result up.
    unsigned char* rep = samples + 500;            // iteration -> first iteration with the same point set
    unsigned char* todo = samples + 600;           // the representatives evaluated in this round
    if (T.tid == 0) {
      cvx::CvRng rng(0xFFFFFFFFFFFFFFFFull);
      unsigned long long* smask = reinterpret_cast<unsigned long long*>(ws.part);     // scratch: 100 point-set masks
      unsigned long long* smask2 = reinterpret_cast<unsigned long long*>(ws.JtJ);
      for (int h = 0; h < 100; ++h) {
        unsigned long long m = 0;
        for (int i = 0; i < 5; ++i) {
          int v;
          for (;;) {
            v = rng.uniform(0, n);
            bool dup = false;
            for (int j = 0; j < i; ++j) dup = dup || samples[h * 5 + j] == v;
            if (!dup) break;
          }
          samples[h * 5 + i] = (unsigned char)v;
          m |= 1ull << v;
        }
        (h < MAXOBS ? smask[h] : smask2[h - MAXOBS]) = m;
        int r = h;
        for (int q = 0; q < h; ++q)
          if ((q < MAXOBS ? smask[q] : smask2[q - MAXOBS]) == m) { r = q; break; }
        rep[h] = (unsigned char)r;
        ws.hyp_cnt[h] = -2;                        // not evaluated yet
      }
      ws.hn = 0;                                   // scratch: next iteration of OpenCV's loop to account for
    }
    T.sync();
    for (;;) {
      if (T.tid == 0) {
        const int pos = ws.hn;
        int ntodo = 0, end = pos;
        if (pos >= ws.pnp_niters || pos >= 100) {
          ntodo = -1;
        } else {
          for (; end < 100 && end < ws.pnp_niters; ++end) {
            const int r = rep[end];
            if (ws.hyp_cnt[r] != -2) continue;     // evaluated in an earlier round (or queued in this one: -3)
            if (ntodo == T.nt || ntodo == 64) break;
            todo[ntodo++] = (unsigned char)r;
            ws.hyp_cnt[r] = -3;
          }
        }
        ws.iters = ntodo; ws.flag = end;
      }
      T.sync();
      const int ntodo = ws.iters, end = ws.flag;
      T.sync();
      if (ntodo < 0) break;
      if (T.tid < ntodo) {
        const int h = todo[T.tid];
        double obj[15], xn[10], R[9], t[3];
        for (int i = 0; i < 5; ++i) {
          const int k = samples[h * 5 + i];
          obj[3 * i] = ws.pnp_obj[3 * k]; obj[3 * i + 1] = ws.pnp_obj[3 * k + 1]; obj[3 * i + 2] = ws.pnp_obj[3 * k + 2];
          xn[2 * i] = cvx::normalize_px(ws.pnp_px[2 * k], cam->K[2], cam->K[0]);
          xn[2 * i + 1] = cvx::normalize_px(ws.pnp_px[2 * k + 1], cam->K[5], cam->K[4]);
        }
        cvx::Epnp5 e;
        e.compute_pose(obj, xn, cam->K, R, t);
        unsigned long long m = 0;
        ws.hyp_cnt[h] = cvx::pnp_inliers(ws.pnp_obj, ws.pnp_px, n, cam->K, R, t, &m);
        hyp_mask[h] = m;
        for (int k = 0; k < 9; ++k) hyp_pose[h * 12 + k] = R[k];
        for (int k = 0; k < 3; ++k) hyp_pose[h * 12 + 9 + k] = t[k];
      }
      T.sync();
      if (T.tid == 0) {
        int q = ws.hn;
        for (; q < end && q < ws.pnp_niters; ++q) {
          const int r = rep[q];
          const int good = ws.hyp_cnt[r];
          if (good > (ws.pnp_maxgood > 4 ? ws.pnp_maxgood : 4)) {
            ws.pnp_best = r; ws.pnp_maxgood = good;
            ws.pnp_niters = cvx::update_num_iters(0.99, (double)(n - good) / n, 5, ws.pnp_niters);
          }
        }
        ws.hn = end;
      }
      T.sync();
    }
    const int best = ws.pnp_best;
    T.sync();
    if (best >= 0) {
      // the iterative solver on the consensus set, started from the winning sample's pose
      // (OpenCV 4.x hands the RANSAC model to the final solvePnP as its extrinsic guess: on the
      // planar pitch the refit therefore stays in the basin - often the mirrored planar pose - that
      // EPnP put the sample in)
      if (T.tid == 0) {
        ws.pnp_mask = hyp_mask[best];
        ws.pnp_status = 1;
        for (int k = 0; k < 12; ++k) ws.pnp_pose[k] = hyp_pose[best * 12 + k];
        set_pose(cam, ws.pnp_pose, ws.pnp_pose + 9);
      }
      T.sync();
      const bool ok = cam->ok != 0;
      T.sync();
      if (ok) refine_camera_masked(T, ws, n, ws.pnp_mask, cam);
      return;
    }
  }
  const bool reproduced = ws.pnp_status != 0;
  T.sync();
  if (reproduced && n <= 5) return;
  // least squares over the consensus set (all matches when the RANSAC failed)
  if (!keep_init) {
    if (T.tid == 0) {
      ws.hn = 0;
      for (int k = 0; k < n; ++k) {
        if (ws.pnp_obj[3 * k + 2] != 0.0 || !((ws.pnp_mask >> k) & 1ull)) continue;     // ground-plane inliers
        ws.hx[ws.hn] = ws.pnp_obj[3 * k]; ws.hy[ws.hn] = ws.pnp_obj[3 * k + 1];
        ws.hu[ws.hn] = ws.pnp_px[2 * k]; ws.hv[ws.hn] = ws.pnp_px[2 * k + 1];
        ++ws.hn;
      }
    }
    T.sync();
    const bool enough = ws.hn >= 4;
    T.sync();
    const bool fit = enough && homography_fit(T, ws, nullptr, ws.H);
    if (T.tid == 0) {
      Pose p;
      const bool ok = fit && pose_from_homography(ws.H, cam->K[0], cam->K[4], cam->K[2], cam->K[5], &p);
      if (ok) {
        set_pose(cam, p.R, p.t);
      } else if (ws.pnp_best >= 0) {
        set_pose(cam, ws.pnp_pose, ws.pnp_pose + 9);
      } else {
        cam->ok = 0;
      }
    }
    T.sync();
  }
  const bool ok = cam->ok != 0;
  T.sync();
  // (the caller's refine_camera over the same matches follows: its least squares starts from this pose)
  if (ok && !refine_follows) refine_camera_masked(T, ws, n, ws.pnp_mask, cam);
}

CAL_HD inline void solve_pnp(const Team& T, Workspace& ws, const CalSolveParams& P, const Points& pts, CamState* cam,
                             bool keep_init, bool refine_follows = false) {
  if (T.tid == 0) {
    for (int k = 0; k < pts.n; ++k) {
      const double* w = P.pitch_xyz + 3 * pts.id[k];
      ws.pnp_obj[3 * k] = f32r(w[0]); ws.pnp_obj[3 * k + 1] = f32r(w[1]); ws.pnp_obj[3 * k + 2] = f32r(w[2]);
      ws.pnp_px[2 * k] = f32r(pts.x[k]); ws.pnp_px[2 * k + 1] = f32r(pts.y[k]);
    }
  }
  T.sync();
  solve_pnp_core(T, ws, pts.n, cam, keep_init, refine_follows);
}

// Camera.projection_rmse (camera.py:249-277): mean L2 distance; project_point rounds the
// normalised coordinates to float32 (distort(), camera.py:247) and returns the origin for
// points at depth <= 1e-3
CAL_HD_NOINLINE inline double projection_rmse(const Team& T, Workspace& ws, const CalSolveParams& P, const Points& pts, const CamState& cam) {
  for (int k = T.tid; k < pts.n; k += T.nt) {
    const double* w = P.pitch_xyz + 3 * pts.id[k];
    const double d[3] = {w[0] - cam.pos[0], w[1] - cam.pos[1], w[2] - cam.pos[2]};
    double r[3];
    mat3_vec(cam.R, d, r);
    double px = 0.0, py = 0.0;
    if (!(r[2] <= 1e-3)) {
      px = f32r(r[0] / r[2]) * cam.fx + cam.ppx;
      py = f32r(r[1] / r[2]) * cam.fy + cam.ppy;
    }
    const double du = pts.x[k] - px, dv = pts.y[k] - py;
    ws.part[k] = sqrt(du * du + dv * dv);
  }
  T.sync();
  if (T.tid == 0) {
    double s = 0.0;
    for (int k = 0; k < pts.n; ++k) s += ws.part[k];
    ws.cand_cost = pts.n > 0 ? s / pts.n : NAN;
  }
  T.sync();
  const double r = ws.cand_cost;
  T.sync();
  return r;
}

// good_camera / is_good_camera (prediction.py:469-484)
CAL_HD inline bool feasible(const CamState& c) {
  return c.ok && c.K[0] >= 10 && c.K[0] <= 20000 && -250 < c.pos[0] && c.pos[0] < 250 && -250 < c.pos[1] &&
         c.pos[1] < 250 && -100 < c.pos[2] && c.pos[2] < 0;
}

// ---- camera from the ground-plane homography (prediction.py:487-520) -------------------------
CAL_HD inline void load_ground_for_homography(const Team& T, Workspace& ws, const CalSolveParams& P, const Points& pts) {
  if (T.tid == 0) {
    ws.hn = 0;
    for (int i = 0; i < NKP; ++i) {
      if (is_top(i)) continue;
      const int k = pts.find(i);
      if (k < 0) continue;
      plane_xy(P, 0, i, &ws.hx[ws.hn], &ws.hy[ws.hn]);
      ws.hu[ws.hn] = f32r(pts.x[k]); ws.hv[ws.hn] = f32r(pts.y[k]);
      ws.hid[ws.hn] = i;
      ++ws.hn;
    }
  }
  T.sync();
}

// Camera.estimate_calibration_matrix_from_plane_homography (camera.py:366-426); thread 0 only
CAL_HD_NOINLINE inline bool k_from_homography(const double* H, double ppx, double ppy, double* fx, double* fy) {
  double A[36];
  for (int k = 0; k < 36; ++k) A[k] = 0.0;
  A[0 * 6 + 1] = 1.0;
  A[1 * 6 + 0] = 1.0; A[1 * 6 + 2] = -1.0;
  A[2 * 6 + 3] = ppy / ppx; A[2 * 6 + 4] = -1.0;
  A[3 * 6 + 0] = H[0] * H[1]; A[3 * 6 + 1] = H[0] * H[4] + H[1] * H[3]; A[3 * 6 + 2] = H[3] * H[4];
  A[3 * 6 + 3] = H[0] * H[7] + H[1] * H[6]; A[3 * 6 + 4] = H[3] * H[7] + H[4] * H[6]; A[3 * 6 + 5] = H[6] * H[7];
  A[4 * 6 + 0] = H[0] * H[0] - H[1] * H[1]; A[4 * 6 + 1] = 2 * H[0] * H[3] - 2 * H[1] * H[4];
  A[4 * 6 + 2] = H[3] * H[3] - H[4] * H[4]; A[4 * 6 + 3] = 2 * H[0] * H[6] - 2 * H[1] * H[7];
  A[4 * 6 + 4] = 2 * H[3] * H[6] - 2 * H[4] * H[7]; A[4 * 6 + 5] = H[6] * H[6] - H[7] * H[7];
  double w[6];
  null_vector6(A, w);
  if (!(fabs(w[5]) > 0)) return false;
  const double W[9] = {w[0] / w[5], w[1] / w[5], w[3] / w[5], w[1] / w[5], w[2] / w[5], w[4] / w[5],
                       w[3] / w[5], w[4] / w[5], 1.0};
  // Cholesky W = L L^T (numpy.linalg.cholesky, lower)
  double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = W[i * 3 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 3 + k] * L[j * 3 + k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[i * 3 + i] = sqrt(s);
      } else {
        L[i * 3 + j] = s / L[j * 3 + j];
      }
    }
  // K = inv(L^T), normalised by K[2][2]
  const double Lt[9] = {L[0], L[3], L[6], 0, L[4], L[7], 0, 0, L[8]};
  double K[9];
  if (!mat3_inv(Lt, K)) return false;
  *fx = K[0] / K[8];
  *fy = K[4] / K[8];
  return isfinite(*fx) && isfinite(*fy);
}

// memo key of a point selection: ids AND coordinates (a line-intersection point merged at one
// threshold can be replaced by the detected keypoint of the same id at the next)
CAL_HD inline unsigned long long points_mask(const Points& pts) {
  unsigned long long h = 1469598103934665603ull;
  for (int k = 0; k < pts.n; ++k) {
    unsigned long long w[3];
    w[0] = (unsigned long long)pts.id[k];
    memcpy(&w[1], &pts.x[k], 8);
    memcpy(&w[2], &pts.y[k], 8);
    for (int j = 0; j < 3; ++j) { h ^= w[j]; h *= 1099511628211ull; }
  }
  return h ^ ((unsigned long long)pts.n << 58);
}

// get_camera_from_homography: ws.hom / ws.hom_rmse; hom.ok = 0 when the reference returns None
CAL_HD_NOINLINE inline void homography_camera(const Team& T, Workspace& ws, const CalSolveParams& P, const Points& pts) {
  // same point set (ids and coordinates) as the last call: the result is still in ws.hom
  const unsigned long long mask = points_mask(pts);
  const bool memo = ws.hom_memo_valid && ws.hom_mask == mask;
  T.sync();
  if (memo) return;
  if (T.tid == 0) { ws.hom.ok = 0; ws.hom_rmse = 10000.0; ws.hom_mask = mask; ws.hom_memo_valid = 1; }
  load_ground_for_homography(T, ws, P, pts);
  const int n = ws.hn;
  T.sync();
  if (n < 4) return;
  homography_ransac(T, ws, 10.0);
  const bool found = ws.flag != 0;
  T.sync();
  if (!found) return;
  if (T.tid == 0) {
    CamState& c = ws.hom;
    c.ppx = 960 / 2.0; c.ppy = 540 / 2.0;            // Camera() defaults (prediction.py:500)
    double fx = 1.0, fy = 1.0;
    const bool kok = k_from_homography(ws.H, c.ppx, c.ppy, &fx, &fy);
    c.ok = kok ? 1 : 0;                              // (reference: K stays identity -> unusable camera)
    c.fx = fx; c.fy = fy;
    const double K[9] = {fx, 0, c.ppx, 0, fy, c.ppy, 0, 0, 1};
    for (int k = 0; k < 9; ++k) c.K[k] = K[k];
    Pose p;
    if (kok && pose_from_homography(ws.H, fx, fy, c.ppx, c.ppy, &p)) {
      for (int k = 0; k < 9; ++k) c.R[k] = p.R[k];
      for (int i = 0; i < 3; ++i) c.pos[i] = -(p.R[i] * p.t[0] + p.R[3 + i] * p.t[1] + p.R[6 + i] * p.t[2]);
    } else {
      c.ok = 0;
    }
  }
  T.sync();
  const bool ok = ws.hom.ok != 0;
  T.sync();
  if (!ok) return;
  solve_pnp(T, ws, P, pts, &ws.hom, true, true);
  const bool pnp_ok = ws.hom.ok != 0;
  T.sync();
  if (pnp_ok) refine_camera(T, ws, P, pts, &ws.hom);
  const double r = projection_rmse(T, ws, P, pts, ws.hom);
  if (T.tid == 0) ws.hom_rmse = r;
  T.sync();
}

// get_camera_all_points + get_camera_gen (prediction.py:523-555, 609-640) on `pts`:
// result in ws.cam (ok flag) and returned rmse
CAL_HD_NOINLINE inline double all_points_camera(const Team& T, Workspace& ws, const CalSolveParams& P, const Points& pts) {
  if (T.tid == 0) ws.cam.ok = 0;
  T.sync();
  int first_plane, total_len, n_ground;
  const int nv = build_views(T, ws, P, pts, 6, true, &first_plane, &total_len, &n_ground);
  if (!(nv > 0 && total_len > 6)) return NAN;
  if (!calibrate_views(T, ws, P)) return NAN;
  camera_from_view0(T, ws, P, &ws.cam);
  // n_groundplane is always 2 there (len of a dict): solve_pnp always runs.  The calibrated
  // pose of a ground-plane view 0 is the natural initial pose; a goal-plane view 0 lives in
  // swapped coordinates, so start from the ground homography instead.
  solve_pnp(T, ws, P, pts, &ws.cam, first_plane == 0, pts.n > 6);
  const bool pnp_ok = ws.cam.ok != 0;
  T.sync();
  if (pnp_ok && pts.n > 6) refine_camera(T, ws, P, pts, &ws.cam);
  const bool ok = ws.cam.ok != 0;
  T.sync();
  if (!ok) return NAN;
  return projection_rmse(T, ws, P, pts, ws.cam);
}

CAL_HD inline void subset_points(const Team& T, Workspace& ws, int kind) {
  // kind 0: all, 1: keep_points (:558-562), 2: ground plane (:565-569)
  if (T.tid == 0) {
    ws.sub.n = 0;
    for (int k = 0; k < ws.pts.n; ++k) {
      const int i = ws.pts.id[k];
      if ((kind == 1 && !in_keep(i)) || (kind == 2 && is_top(i))) continue;
      ws.sub.id[ws.sub.n] = i; ws.sub.x[ws.sub.n] = ws.pts.x[k]; ws.sub.y[ws.sub.n] = ws.pts.y[k];
      ++ws.sub.n;
    }
  }
  T.sync();
}

// get_camera_accurate_points' subset (prediction.py:572-606); returns false if it yields None
CAL_HD_NOINLINE inline bool accurate_subset(const Team& T, Workspace& ws, const CalSolveParams& P, double thr) {
  load_ground_for_homography(T, ws, P, ws.pts);
  const int n = ws.hn;
  T.sync();
  if (n < 4) return false;
  homography_ransac(T, ws, thr);
  const bool found = ws.flag != 0;
  T.sync();
  if (!found) return false;
  if (T.tid == 0) {
    // insertion order of the reference: ground points by id (reprojection < thr), then crossbars
    ws.sub.n = 0;
    for (int j = 0; j < ws.hn; ++j) {
      const double w = ws.H[6] * ws.hx[j] + ws.H[7] * ws.hy[j] + ws.H[8];
      const double du = (ws.H[0] * ws.hx[j] + ws.H[1] * ws.hy[j] + ws.H[2]) / w - ws.hu[j];
      const double dv = (ws.H[3] * ws.hx[j] + ws.H[4] * ws.hy[j] + ws.H[5]) / w - ws.hv[j];
      if (sqrt(du * du + dv * dv) < thr) {
        const int k = ws.pts.find(ws.hid[j]);
        ws.sub.id[ws.sub.n] = ws.hid[j]; ws.sub.x[ws.sub.n] = ws.pts.x[k]; ws.sub.y[ws.sub.n] = ws.pts.y[k];
        ++ws.sub.n;
      }
    }
    const int tops[4] = {0, 1, 24, 25};
    for (int q = 0; q < 4; ++q) {
      const int k = ws.pts.find(tops[q]);
      if (k < 0) continue;
      ws.sub.id[ws.sub.n] = tops[q]; ws.sub.x[ws.sub.n] = ws.pts.x[k]; ws.sub.y[ws.sub.n] = ws.pts.y[k];
      ++ws.sub.n;
    }
  }
  T.sync();
  return true;
}

// ---- the five algorithms ------------------------------------------------------------------
// Each leaves its result in ws.best (ok flag), ws.best_rmse, ws.best_tag.

CAL_HD inline void clear_best(const Team& T, Workspace& ws) {
  if (T.tid == 0) { ws.best.ok = 0; ws.best_rmse = NAN; ws.best_tag = BR_NONE; }
  T.sync();
}
CAL_HD inline void set_best(const Team& T, Workspace& ws, const CamState& c, double rmse, int tag) {
  T.sync();
  if (T.tid == 0) { ws.best = c; ws.best_rmse = rmse; ws.best_tag = tag; }
  T.sync();
}

// prediction.py:138-170
CAL_HD inline void algo_opencv_calibration(const Team& T, Workspace& ws, const CalSolveParams& P, const Frame& fr, float thr) {
  clear_best(T, ws);
  if (T.tid == 0) {
    ws.pts.n = 0;
    for (int i = 0; i < NKP; ++i)
      if (!is_top(i) && fr.pred[3 * i + 2] > thr) {
        ws.pts.id[ws.pts.n] = i; ws.pts.x[ws.pts.n] = (double)fr.pred[3 * i]; ws.pts.y[ws.pts.n] = (double)fr.pred[3 * i + 1];
        ++ws.pts.n;
      }
  }
  T.sync();
  const int n = ws.pts.n;
  T.sync();
  if (n <= 5) return;
  int fp, tl, ng;
  build_views(T, ws, P, ws.pts, 1, false, &fp, &tl, &ng, 1);
  if (!calibrate_views(T, ws, P)) return;
  camera_from_view0(T, ws, P, &ws.cam);
  set_best(T, ws, ws.cam, NAN, BR_CALIBRATION);
}

// prediction.py:172-243
CAL_HD inline void algo_multiplane(const Team& T, Workspace& ws, const CalSolveParams& P, const Frame& fr, float thr) {
  clear_best(T, ws);
  select_points(T, ws, P, fr, thr, true);
  merge_line_points(T, ws, P, fr, 0);
  const int n = ws.pts.n;
  int fp, tl, ng;
  const int nv = build_views(T, ws, P, ws.pts, P.min_points_per_plane, false, &fp, &tl, &ng);
  if (!(nv > 0 && n > P.min_points)) return;
  if (!calibrate_views(T, ws, P)) return;
  const bool f_ok = ws.f > P.min_focal_length;
  T.sync();
  if (!f_ok) return;
  camera_from_view0(T, ws, P, &ws.cam);
  int tag = BR_MULTIPLANE;
  if (n > P.min_points_for_refinement) { refine_camera(T, ws, P, ws.pts, &ws.cam); tag |= BR_REFINED; }
  const bool ok = ws.cam.ok != 0;
  T.sync();
  if (ok) set_best(T, ws, ws.cam, NAN, tag);
}

// prediction.py:339-437
CAL_HD_NOINLINE inline void algo_original_voter(const Team& T, Workspace& ws, const CalSolveParams& P, const Frame& fr, float thr) {
  clear_best(T, ws);
  select_points(T, ws, P, fr, thr, true);
  merge_line_points(T, ws, P, fr, 2);
  const int n = ws.pts.n;
  homography_camera(T, ws, P, ws.pts);
  int fp, tl, ng;
  const int nv = build_views(T, ws, P, ws.pts, P.min_points_per_plane, false, &fp, &tl, &ng);
  bool have = false;
  int tag = BR_OV_CALIBRATION;
  if (nv > 0 && n > P.min_points) {
    // a degenerate configuration makes cv2.calibrateCamera raise: the exception leaves
    // original_voter without a camera (and without the homography fallback)
    if (!calibrate_views(T, ws, P)) return;
    camera_from_view0(T, ws, P, &ws.cam);
    if (ng < P.min_points_per_plane) { solve_pnp(T, ws, P, ws.pts, &ws.cam, false); tag = BR_OV_CALIBRATION_PNP; }
    const bool good = feasible(ws.cam);
    T.sync();
    if (good) {
      if (n > P.min_points_for_refinement) { refine_camera(T, ws, P, ws.pts, &ws.cam); tag |= BR_REFINED; }
      have = ws.cam.ok != 0;
      T.sync();
    }
  }
  if (have) {
    const double r = projection_rmse(T, ws, P, ws.pts, ws.cam);
    set_best(T, ws, ws.cam, r, tag);
    return;
  }
  const bool hom_ok = ws.hom.ok != 0 && ws.hom_rmse < 26;
  T.sync();
  if (hom_ok) set_best(T, ws, ws.hom, ws.hom_rmse, BR_OV_HOMOGRAPHY);
}

// prediction.py:259-330
CAL_HD_NOINLINE inline void algo_voter(const Team& T, Workspace& ws, const CalSolveParams& P, const Frame& fr, float thr) {
  clear_best(T, ws);
  if (T.tid == 0) ws.memo_n = 0;
  select_points(T, ws, P, fr, thr, false);
  merge_line_points(T, ws, P, fr, 1);
  homography_camera(T, ws, P, ws.pts);
  // candidates in the reference's list order: reliable, accurate, all, ground
  bool any = false, best_flag = false;
  double best_rmse = 0.0;
  for (int c = 0; c < 4; ++c) {
    bool has_subset = true;
    if (c == 0) subset_points(T, ws, 1);
    else if (c == 1) has_subset = accurate_subset(T, ws, P, 5.0);
    else if (c == 2) subset_points(T, ws, 0);
    else subset_points(T, ws, 2);
    if (!has_subset) continue;
    const unsigned long long mask = points_mask(ws.sub);
    int hit = -1;
    for (int q = 0; q < ws.memo_n; ++q) if (ws.memo_mask[q] == mask) hit = q;
    T.sync();
    double r;
    if (hit >= 0) {
      r = ws.memo_rmse[hit];
      if (T.tid == 0) ws.cam = ws.memo_cam[hit];
      T.sync();
    } else {
      r = all_points_camera(T, ws, P, ws.sub);
      if (T.tid == 0) {
        ws.memo_cam[ws.memo_n] = ws.cam; ws.memo_rmse[ws.memo_n] = r; ws.memo_mask[ws.memo_n] = mask;
        ++ws.memo_n;
      }
      T.sync();
    }
    const bool ok = ws.cam.ok != 0 && feasible(ws.cam) && !isnan(r);
    T.sync();
    if (!ok) continue;
    // max over key (tag == 'camera_rel' and rmse < max_rmse_rel, 1 / rmse), first maximum wins
    const bool flag = (c == 0) && (r < P.max_rmse_rel);
    const bool better = !any || (flag && !best_flag) || (flag == best_flag && 1.0 / r > 1.0 / best_rmse);
    if (better) {
      any = true; best_flag = flag; best_rmse = r;
      set_best(T, ws, ws.cam, r, BR_VOTER_REL + c);
    }
  }
  if (any && best_rmse < P.max_rmse) return;
  clear_best(T, ws);
  const bool hom_ok = ws.hom.ok != 0 && ws.hom_rmse < P.max_rmse;
  T.sync();
  if (hom_ok) set_best(T, ws, ws.hom, ws.hom_rmse, BR_VOTER_HOMOGRAPHY);
}

// prediction.py:245-257
CAL_HD inline void algo_iterative_voter(const Team& T, Workspace& ws, const CalSolveParams& P, const Frame& fr) {
  algo_original_voter(T, ws, P, fr, 0.5f);
  bool ok = ws.best.ok != 0;
  T.sync();
  if (ok) return;
  for (int k = 0; k < P.n_conf_threshs; ++k) {
    algo_voter(T, ws, P, fr, P.conf_threshs[k]);
    ok = ws.best.ok != 0;
    T.sync();
    if (ok) return;
  }
}

CAL_HD inline void solve_frame(const Team& T, Workspace& ws, const CalSolveParams& P, const Frame& fr, CalCameraRecord* out) {
  if (T.tid == 0) { ws.hom_memo_valid = 0; ws.memo_n = 0; }
  T.sync();
  switch (P.algorithm) {
    case 0: algo_opencv_calibration(T, ws, P, fr, P.conf_thresh); break;
    case 1: algo_multiplane(T, ws, P, fr, P.conf_thresh); break;
    case 2: algo_original_voter(T, ws, P, fr, P.conf_thresh); break;
    case 3: algo_voter(T, ws, P, fr, P.conf_thresh); break;
    default: algo_iterative_voter(T, ws, P, fr); break;
  }
  T.sync();
  if (T.tid == 0) {
    const CamState& c = ws.best;
    const bool ok = c.ok != 0;
    for (int k = 0; k < 3; ++k) out->position[k] = ok ? c.pos[k] : 0.0;
    for (int k = 0; k < 9; ++k) out->rotation[k] = ok ? c.R[k] : 0.0;
    out->fx = ok ? c.fx : 0.0;
    out->fy = ok ? c.fy : 0.0;
    out->rmse = ok ? ws.best_rmse : 0.0;
    out->valid = ok ? 1 : 0;
    out->branch = ok ? ws.best_tag : BR_NONE;
  }
  T.sync();
}

}  // namespace solve
}  // namespace cal
