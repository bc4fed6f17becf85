// evaluate.cu - the official calibration metric, one thread block per frame (calib_b200.h, "metric").
// fp64 throughout, in the reference's operation order (numpy cross products, divisions and square roots
// of baseline/evaluate_camera.py), so that the comparisons against the pixel threshold and the image
// border fall on the same side.
#include <math_constants.h>

#define CAL_TU "evaluate.cu"
#include "common.cuh"

namespace cal {
namespace {

constexpr int EV_THREADS = 128;
constexpr int NC = CAL_EVAL_CLASSES;

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// Camera.project_point (baseline/camera.py:249-268) with all distortion coefficients zero: the
// normalised point is rounded to float32 (distort(), :247)
__device__ __forceinline__ void project_point(const CalCameraRecord& c, double ppx, double ppy, const double* p, double* ext) {
  const double d[3] = {p[0] - c.position[0], p[1] - c.position[1], p[2] - c.position[2]};
  double r[3];
  for (int i = 0; i < 3; ++i) r[i] = c.rotation[3 * i] * d[0] + c.rotation[3 * i + 1] * d[1] + c.rotation[3 * i + 2] * d[2];
  if (r[2] <= 1e-3) { ext[0] = ext[1] = ext[2] = 0.0; return; }
  const double x = static_cast<double>(static_cast<float>(r[0] / r[2])), y = static_cast<double>(static_cast<float>(r[1] / r[2]));
  ext[0] = x * c.fx + ppx;
  ext[1] = y * c.fy + ppy;
  ext[2] = 1.0;
}

// the border crossing of the segment (prev, ext) nearest to ext (evaluate_camera.py:52-74, 86-104)
__device__ bool border_crossing(const double* ext, const double* prev, int width, int height, double* out) {
  double line[3];
  cross3(ext, prev, line);
  const double sides[4][3] = {{1, 0, 0}, {1, 0, -static_cast<double>(width) + 1}, {0, 1, 0}, {0, 1, -static_cast<double>(height) + 1}};
  bool have = false;
  double best = 0.0;
  for (int s = 0; s < 4; ++s) {
    double it[3];
    cross3(line, sides[s], it);
    const double w = it[2];
    it[0] /= w; it[1] /= w; it[2] /= w;
    if (0 <= it[0] && it[0] < width && 0 <= it[1] && it[1] < height) {
      const double d0 = it[0] - ext[0], d1 = it[1] - ext[1], d2 = it[2] - ext[2];
      const double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
      if (!have || dist < best) { have = true; best = dist; out[0] = it[0]; out[1] = it[1]; }
    }
  }
  return have;
}

__device__ __forceinline__ double dist2d(double ax, double ay, double bx, double by) {
  const double dx = ax - bx, dy = ay - by;
  return sqrt(dx * dx + dy * dy);
}

// distance_to_polyline (evaluate_camera.py:110-160)
__device__ double distance_to_polyline(double px, double py, const double* poly, int n) {
  if (n < 2) return dist2d(px, py, poly[0], poly[1]);
  double best = CUDART_INF;
  const double pt[3] = {px, py, 1.0};
  for (int i = 0; i + 1 < n; ++i) {
    const double o[3] = {poly[2 * i], poly[2 * i + 1], 1.0}, e[3] = {poly[2 * i + 2], poly[2 * i + 3], 1.0};
    double line[3];
    cross3(o, e, line);
    const double nrm = sqrt(line[0] * line[0] + line[1] * line[1]);
    line[0] /= nrm; line[1] /= nrm; line[2] /= nrm;
    const double dir[3] = {line[0], line[1], 0.0};
    double t[3], pr[3];
    cross3(dir, pt, t);
    cross3(t, line, pr);
    const double w = pr[2];
    pr[0] /= w; pr[1] /= w; pr[2] /= w;
    const double v1[3] = {pr[0] - o[0], pr[1] - o[1], pr[2] - o[2]}, v2[3] = {e[0] - o[0], e[1] - o[1], e[2] - o[2]};
    const double k = (v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) / (v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
    double sd;
    if (0 < k && k < 1) {
      const double d0 = pr[0] - pt[0], d1 = pr[1] - pt[1], d2 = pr[2] - pt[2];
      sd = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    } else {
      const double d1 = dist2d(px, py, o[0], o[1]), d2 = dist2d(px, py, e[0], e[1]);
      sd = d1 < d2 ? d1 : d2;                    // np.min: a NaN would propagate; distances of finite points are finite
    }
    if (sd < best || sd != sd) best = sd;        // np.min semantics (NaN wins)
    if (best != best) break;
  }
  return best;
}

__global__ void __launch_bounds__(EV_THREADS) evaluate_kernel(
    const CalCameraRecord* __restrict__ cams, const double* __restrict__ field_pts, const int32_t* __restrict__ class_off,
    const int32_t* __restrict__ class_id, int n_proj, const int32_t* __restrict__ mirror, const uint8_t* __restrict__ is_circle,
    const double* __restrict__ gt_pts, const int32_t* __restrict__ gt_count, int max_gt, int width, int height,
    double threshold, double* __restrict__ poly, int32_t* __restrict__ poly_count, int max_poly, int from_polylines,
    double* __restrict__ dist, CalEvalRecord* __restrict__ out) {
  const int b = blockIdx.x;
  const CalCameraRecord cam = cams[b];
  CalEvalRecord* rec = out + b;
  __shared__ int s_proj_of[NC];          // dataset class -> projectable slot, -1 if none
  __shared__ int s_cnt[NC];              // polyline length per dataset class
  if (!cam.valid && !(from_polylines & 1)) {
    if (threadIdx.x == 0) {
      rec->accuracy = 0; rec->l2_sum = 0; rec->l2_count = 0; rec->labelling = 0; rec->valid = 0; rec->pad = 0;
      for (int k = 0; k < 4; ++k) rec->confusion[k] = 0;
      for (int c = 0; c < NC; ++c) { rec->touched[c] = 0; for (int k = 0; k < 4; ++k) rec->per_class[c][k] = 0; }
    }
    return;
  }
  for (int c = threadIdx.x; c < NC; c += EV_THREADS) { s_proj_of[c] = -1; s_cnt[c] = 0; }
  __syncthreads();
  // ---- 1. polylines: one thread per class walks its sampled points in order (get_polylines)
  double* my_poly = poly + static_cast<size_t>(b) * n_proj * max_poly * 2;
  int32_t* my_cnt = poly_count + static_cast<size_t>(b) * n_proj;
  for (int s = threadIdx.x; s < n_proj; s += EV_THREADS) {
    double* pl = my_poly + static_cast<size_t>(s) * max_poly * 2;
    int n = 0;
    if (from_polylines & 1) {
      n = my_cnt[s];
    } else {
      const double ppx = width / 2.0, ppy = height / 2.0;   // Camera(width, height).principal_point of the solved camera
      bool in_img = false;
      double prev[3] = {0, 0, 0};
      const int p0 = class_off[s], p1 = class_off[s + 1];
      for (int i = p0; i < p1; ++i) {
        double ext[3];
        project_point(cam, ppx, ppy, field_pts + 3 * i, ext);
        if (ext[2] < 1e-5) continue;                      // at infinity or behind the camera
        const bool inside = 0 <= ext[0] && ext[0] < width && 0 <= ext[1] && ext[1] < height;
        if (inside) {
          if (!in_img && i > p0) {
            double it[2];
            if (border_crossing(ext, prev, width, height, it) && n < max_poly) { pl[2 * n] = it[0]; pl[2 * n + 1] = it[1]; ++n; }
          }
          if (n < max_poly) { pl[2 * n] = ext[0]; pl[2 * n + 1] = ext[1]; ++n; }
          in_img = true;
        } else if (in_img) {
          double it[2];
          if (border_crossing(ext, prev, width, height, it) && n < max_poly) { pl[2 * n] = it[0]; pl[2 * n + 1] = it[1]; ++n; }
          in_img = false;
        }
        prev[0] = ext[0]; prev[1] = ext[1]; prev[2] = ext[2];
      }
      my_cnt[s] = n;
    }
    s_proj_of[class_id[s]] = s;
    s_cnt[class_id[s]] = n;
  }
  __syncthreads();
  // ---- 2. distances of every annotated point to the polyline of its class, for both labellings
  const double* my_gt = gt_pts + static_cast<size_t>(b) * NC * max_gt * 2;
  const int32_t* my_gc = gt_count + static_cast<size_t>(b) * NC;
  double* my_dist = dist + static_cast<size_t>(b) * 2 * NC * max_gt;
  for (int item = threadIdx.x; item < 2 * NC * max_gt; item += EV_THREADS) {
    const int l = item / (NC * max_gt), rem = item - l * NC * max_gt, c = rem / max_gt, j = rem - c * max_gt;
    if (j >= my_gc[c]) continue;
    const int k = l == 0 ? c : mirror[c];                  // the class these points are compared with
    if (s_cnt[k] <= 0) continue;
    const double* pl = my_poly + static_cast<size_t>(s_proj_of[k]) * max_poly * 2;
    my_dist[item] = distance_to_polyline(my_gt[(c * max_gt + j) * 2], my_gt[(c * max_gt + j) * 2 + 1], pl, s_cnt[k]);
  }
  __syncthreads();
  // ---- 3. confusion matrices (evaluate_camera_prediction) and the choice of the labelling (Evaluator)
  if (threadIdx.x < 2) {
    const int l = threadIdx.x;
    double conf[4] = {0, 0, 0, 0}, l2 = 0.0;
    int l2n = 0;
    __shared__ double s_conf[2][4], s_l2[2], s_pc[2][NC][4];
    __shared__ int s_l2n[2];
    __shared__ unsigned char s_touch[2][NC];
    for (int k = 0; k < NC; ++k) {
      double pc[4] = {0, 0, 0, 0};
      const int c = l == 0 ? k : mirror[k];                // annotated class that carries label k under this labelling
      const bool detected = s_cnt[k] > 0, annotated = my_gc[c] >= 0;
      unsigned char touched = 0;
      if (detected && !annotated) {
        pc[1] = is_circle[k] ? 9.0 : 2.0;
        conf[1] += 1; touched = 1;
      } else if (!detected && annotated) {
        pc[2] = my_gc[c];
        conf[2] += 1; touched = 1;
      } else if (detected && annotated) {
        bool all_below = true;
        double cls_sum = 0.0;
        for (int j = 0; j < my_gc[c]; ++j) {
          const double d = my_dist[(l * NC + c) * max_gt + j];
          if (d < threshold) pc[0] += 1; else { pc[1] += 1; all_below = false; }
          cls_sum += d;
          ++l2n;
        }
        l2 += cls_sum;
        if (all_below) conf[0] += 1; else conf[1] += 1;
        touched = 1;
      }
      for (int q = 0; q < 4; ++q) s_pc[l][k][q] = pc[q];
      s_touch[l][k] = touched;
    }
    for (int q = 0; q < 4; ++q) s_conf[l][q] = conf[q];
    s_l2[l] = l2; s_l2n[l] = l2n;
    __syncwarp(0x3);
    if (l == 0) {
      // float32 confusion matrices in the reference: the accuracy is a float32 quotient
      float acc[2];
      for (int q = 0; q < 2; ++q) {
        const float tot = static_cast<float>(s_conf[q][0]) + static_cast<float>(s_conf[q][1]) + static_cast<float>(s_conf[q][2]);
        acc[q] = tot > 0 ? static_cast<float>(s_conf[q][0]) / tot : 0.0f;
      }
      const int pick = (from_polylines & 2) ? 0 : (acc[0] > acc[1] ? 0 : 1);
      rec->accuracy = acc[pick];
      for (int q = 0; q < 4; ++q) rec->confusion[q] = s_conf[pick][q];
      rec->l2_sum = s_l2[pick]; rec->l2_count = s_l2n[pick];
      rec->labelling = pick; rec->valid = 1; rec->pad = 0;
      for (int k = 0; k < NC; ++k) {
        rec->touched[k] = s_touch[pick][k];
        for (int q = 0; q < 4; ++q) rec->per_class[k][q] = s_pc[pick][k][q];
      }
    }
  }
}

}  // namespace
}  // namespace cal

extern "C" int cal_evaluate_cameras(const CalCameraRecord* cams, int B, const double* field_pts, const int32_t* class_off,
                                    const int32_t* class_id, int n_proj, const int32_t* mirror, const uint8_t* is_circle,
                                    const double* gt_pts, const int32_t* gt_count, int max_gt, int img_w, int img_h,
                                    double threshold, double* poly, int32_t* poly_count, int max_poly, int from_polylines,
                                    double* dist, CalEvalRecord* out, void* stream) {
  using namespace cal;
  CAL_REQUIRE(B >= 0, CAL_E_INVALID, "cal_evaluate_cameras: B %d", B);
  if (B == 0) return CAL_OK;
  CAL_REQUIRE(cams && field_pts && class_off && class_id && mirror && is_circle && gt_pts && gt_count && poly && poly_count && dist && out,
              CAL_E_INVALID, "cal_evaluate_cameras: null pointer");
  CAL_REQUIRE(n_proj >= 1 && n_proj <= CAL_EVAL_CLASSES && max_gt >= 1 && max_poly >= 2 && img_w > 0 && img_h > 0, CAL_E_INVALID,
              "cal_evaluate_cameras: bad sizes");
  evaluate_kernel<<<B, EV_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(cams, field_pts, class_off, class_id, n_proj, mirror,
                                                                          is_circle, gt_pts, gt_count, max_gt, img_w, img_h, threshold,
                                                                          poly, poly_count, max_poly, from_polylines, dist, out);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}
