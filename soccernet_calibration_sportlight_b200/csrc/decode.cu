// decode.cu - heat-map decoders (HBM-bound streaming reductions).
//
//  cal_kp_decode   : src/models/hrnet/transforms.py:228-239
//  cal_line_decode : src/models/line/transforms.py:216-280
//
// Keypoints.  One CTA per (frame, channel) plane; the background channel C-1 is
// never read.  Each warp streams whole rows with 128-bit loads: a lane keeps the
// running column maxima of its own columns in registers and one shuffle tree per
// row yields the row maximum.  exp() is monotone, so maxima / arg-maxima are taken
// on the raw log-probabilities; the only place exp matters is tie-creation by
// rounding, reproduced exactly by comparing against the smallest float t with
// E(t) == E(max), E = correctly rounded fp32 exp (fp64 evaluation, one rounding).
// Algorithmic traffic: one read of (C-1)*h*w*4 bytes per frame, 12 bytes written
// per plane.
#include <math_constants.h>

#define CAL_TU "decode.cu"
#include "common.cuh"

namespace cal {
namespace {

constexpr int KP_THREADS = 256;
constexpr int KP_WARPS = KP_THREADS / 32;

__device__ __forceinline__ float exp_rn(float x) {
  return __double2float_rn(exp(static_cast<double>(x)));
}
// monotone map float -> uint32 (total order, -inf lowest)
__device__ __forceinline__ uint32_t fkey(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float funkey(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(b);
}
// smallest float t with exp_rn(t) == exp_rn(m)  (gallop down, then bisect)
__device__ float exp_tie_threshold(float m, float em) {
  const uint32_t km = fkey(m);
  const uint32_t kmin = fkey(-CUDART_INF_F);
  uint32_t hi = km, step = 1, lo;
  while (true) {
    if (hi - kmin <= step) { lo = kmin; break; }
    uint32_t cand = hi - step;
    if (exp_rn(funkey(cand)) == em) { hi = cand; step <<= 1; }
    else { lo = cand + 1; break; }
  }
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (exp_rn(funkey(mid)) == em) hi = mid; else lo = mid + 1;
  }
  return funkey(hi);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// Final stage shared by both keypoint kernels: s_col[w], s_row[h] hold raw maxima.
__device__ void kp_finish(const float* s_col, const float* s_row, int h, int w, int H_img,
                          int W_img, float* out3, float* s_red, int* s_redi) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -CUDART_INF_F;
  for (int i = tid; i < w; i += KP_THREADS) m = fmaxf(m, s_col[i]);
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = s_red[0];
    for (int i = 1; i < KP_WARPS; ++i) mm = fmaxf(mm, s_red[i]);
    float em = exp_rn(mm);
    s_red[KP_WARPS] = exp_tie_threshold(mm, em);
    s_red[KP_WARPS + 1] = em;
  }
  __syncthreads();
  const float thr = s_red[KP_WARPS];
  int xi = 0x7fffffff, yi = 0x7fffffff;
  for (int i = tid; i < w; i += KP_THREADS) if (s_col[i] >= thr) { xi = i; break; }
  for (int i = tid; i < h; i += KP_THREADS) if (s_row[i] >= thr) { yi = i; break; }
  xi = warp_min_i(xi);
  yi = warp_min_i(yi);
  if (lane == 0) { s_redi[warp] = xi; s_redi[KP_WARPS + warp] = yi; }
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < KP_WARPS; ++i) {
      xi = min(xi, s_redi[i]);
      yi = min(yi, s_redi[KP_WARPS + i]);
    }
    if (xi == 0x7fffffff) xi = 0;   // only for NaN-poisoned planes
    if (yi == 0x7fffffff) yi = 0;
    out3[0] = __fdiv_rn(static_cast<float>(xi * W_img), static_cast<float>(w));
    out3[1] = __fdiv_rn(static_cast<float>(yi * H_img), static_cast<float>(h));
    out3[2] = s_red[KP_WARPS + 1];
  }
}

// Fast path: w % 4 == 0, w <= NV*128, 16-byte aligned rows.
template <int NV>
__global__ void __launch_bounds__(KP_THREADS) kp_decode_vec_kernel(
    const float* __restrict__ logp, int C, int h, int w, int H_img, int W_img,
    float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* s_col = smem;            // w
  float* s_row = smem + w;        // h
  __shared__ float s_red[KP_WARPS + 2];
  __shared__ int s_redi[2 * KP_WARPS];

  const int plane = blockIdx.x;                  // b*(C-1)+c
  const int b = plane / (C - 1), c = plane - b * (C - 1);
  const float4* base = reinterpret_cast<const float4*>(logp + (static_cast<size_t>(b) * C + c) * h * w);
  const int w4 = w >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  float4 cm[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) cm[k] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  for (int i = tid; i < w; i += KP_THREADS) s_col[i] = -CUDART_INF_F;

  for (int r = warp; r < h; r += KP_WARPS) {
    const float4* row = base + static_cast<size_t>(r) * w4;
    float4 v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int g = lane + 32 * k;
      v[k] = (g < w4) ? ld_stream(row + g)
                      : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    }
    float rm = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      cm[k].x = fmaxf(cm[k].x, v[k].x); cm[k].y = fmaxf(cm[k].y, v[k].y);
      cm[k].z = fmaxf(cm[k].z, v[k].z); cm[k].w = fmaxf(cm[k].w, v[k].w);
      rm = fmaxf(rm, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
    }
    rm = warp_max(rm);
    if (lane == 0) s_row[r] = rm;
  }
  __syncthreads();
  // merge the 8 per-warp column maxima, one warp at a time (exact, no float atomics)
  for (int wi = 0; wi < KP_WARPS; ++wi) {
    if (warp == wi) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int g = lane + 32 * k;
        if (g < w4) {
          float4* p = reinterpret_cast<float4*>(s_col) + g;
          float4 o = *p;
          o.x = fmaxf(o.x, cm[k].x); o.y = fmaxf(o.y, cm[k].y);
          o.z = fmaxf(o.z, cm[k].z); o.w = fmaxf(o.w, cm[k].w);
          *p = o;
        }
      }
    }
    __syncthreads();
  }
  kp_finish(s_col, s_row, h, w, H_img, W_img, out + static_cast<size_t>(plane) * 3, s_red, s_redi);
}

// Generic path: any h, w, any alignment (two scalar passes; the second hits L2).
__global__ void __launch_bounds__(KP_THREADS) kp_decode_generic_kernel(
    const float* __restrict__ logp, int C, int h, int w, int H_img, int W_img,
    float* __restrict__ out) {
  extern __shared__ float smem[];
  float* s_col = smem;
  float* s_row = smem + w;
  __shared__ float s_red[KP_WARPS + 2];
  __shared__ int s_redi[2 * KP_WARPS];
  const int plane = blockIdx.x;
  const int b = plane / (C - 1), c = plane - b * (C - 1);
  const float* base = logp + (static_cast<size_t>(b) * C + c) * h * w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int r = warp; r < h; r += KP_WARPS) {
    float rm = -CUDART_INF_F;
    for (int x = lane; x < w; x += 32) rm = fmaxf(rm, __ldg(base + static_cast<size_t>(r) * w + x));
    rm = warp_max(rm);
    if (lane == 0) s_row[r] = rm;
  }
  for (int x = tid; x < w; x += KP_THREADS) {
    float cmx = -CUDART_INF_F;
    for (int r = 0; r < h; ++r) cmx = fmaxf(cmx, __ldg(base + static_cast<size_t>(r) * w + x));
    s_col[x] = cmx;
  }
  __syncthreads();
  kp_finish(s_col, s_row, h, w, H_img, W_img, out + static_cast<size_t>(plane) * 3, s_red, s_redi);
}

// ------------------------------------------------------------------------ lines
constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;

struct Peak { float v; int i; };

__device__ __forceinline__ Peak better(Peak a, Peak b) {   // max value, ties -> lower index
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ Peak block_best(Peak p, Peak* s_pk) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Peak q;
    q.v = __shfl_xor_sync(0xffffffffu, p.v, o);
    q.i = __shfl_xor_sync(0xffffffffu, p.i, o);
    p = better(p, q);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                       // s_pk reuse between the two passes
  if (lane == 0) s_pk[warp] = p;
  __syncthreads();
  Peak r = s_pk[0];
  for (int i = 1; i < LN_WARPS; ++i) r = better(r, s_pk[i]);
  return r;
}

// value of relu(h) * (1 - G) at (x, y), all roundings as torch's fp32 ops
__device__ __forceinline__ float suppressed(float v, int x, int y, float x1, float y1, float denom) {
  const float hv = fmaxf(v, 0.0f);
  const float dx = __fsub_rn(static_cast<float>(x), x1);
  const float dy = __fsub_rn(static_cast<float>(y), y1);
  const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  const float q = __fdiv_rn(-d2, denom);
  // exp(q) < 2^-25 -> (1 - mask) rounds to exactly 1.0f: skip the fp64 exp
  if (q < -17.5f) return hv;
  const float mask = exp_rn(q);
  return __fmul_rn(hv, __fsub_rn(1.0f, mask));
}

template <bool VEC>
__global__ void __launch_bounds__(LN_THREADS) line_decode_kernel(
    const float* __restrict__ heat, int h, int w, float denom, float scale,
    float* __restrict__ out) {
  __shared__ Peak s_pk[LN_WARPS];
  const int plane = blockIdx.x;
  const float* base = heat + static_cast<size_t>(plane) * h * w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // pass 1: flat first-argmax of relu(h)
  Peak p1{-1.0f, 0x7fffffff};
  for (int r = warp; r < h; r += LN_WARPS) {
    const float* row = base + static_cast<size_t>(r) * w;
    if (VEC) {
      for (int g = lane; g < (w >> 2); g += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + g);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hv = fmaxf(e[j], 0.0f);
          if (hv > p1.v) { p1.v = hv; p1.i = r * w + 4 * g + j; }
        }
      }
    } else {
      for (int x = lane; x < w; x += 32) {
        const float hv = fmaxf(__ldg(row + x), 0.0f);
        if (hv > p1.v) { p1.v = hv; p1.i = r * w + x; }
      }
    }
  }
  p1 = block_best(p1, s_pk);
  const int ix1 = p1.i % w, iy1 = p1.i / w;
  const float x1 = static_cast<float>(ix1), y1 = static_cast<float>(iy1);

  // pass 2 (L2-resident re-read): first-argmax of relu(h) * (1 - gaussian)
  Peak p2{-1.0f, 0x7fffffff};
  for (int r = warp; r < h; r += LN_WARPS) {
    const float* row = base + static_cast<size_t>(r) * w;
    if (VEC) {
      for (int g = lane; g < (w >> 2); g += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + g);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float s = suppressed(e[j], 4 * g + j, r, x1, y1, denom);
          if (s > p2.v) { p2.v = s; p2.i = r * w + 4 * g + j; }
        }
      }
    } else {
      for (int x = lane; x < w; x += 32) {
        const float s = suppressed(__ldg(row + x), x, r, x1, y1, denom);
        if (s > p2.v) { p2.v = s; p2.i = r * w + x; }
      }
    }
  }
  p2 = block_best(p2, s_pk);
  if (tid == 0) {
    float* o = out + static_cast<size_t>(plane) * 6;
    o[0] = __fmul_rn(x1, scale);
    o[1] = __fmul_rn(y1, scale);
    o[2] = p1.v;
    o[3] = __fmul_rn(static_cast<float>(p2.i % w), scale);
    o[4] = __fmul_rn(static_cast<float>(p2.i / w), scale);
    o[5] = p2.v;
  }
}

}  // namespace
}  // namespace cal

extern "C" int cal_kp_decode(const float* logp, int B, int C, int h, int w, int H_img, int W_img,
                             float* out, void* stream) {
  using namespace cal;
  CAL_REQUIRE(B >= 0 && C >= 1 && h >= 1 && w >= 1, CAL_E_INVALID, "cal_kp_decode: bad shape B=%d C=%d h=%d w=%d", B, C, h, w);
  if (B == 0 || C == 1) return CAL_OK;
  CAL_REQUIRE(logp && out, CAL_E_INVALID, "cal_kp_decode: null pointer");
  CAL_REQUIRE(static_cast<long long>(h) * w < (1ll << 30) && static_cast<long long>(w) * W_img < (1ll << 24) &&
                  static_cast<long long>(h) * H_img < (1ll << 24),
              CAL_E_UNSUPPORTED, "cal_kp_decode: plane too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int planes = B * (C - 1);
  const size_t smem = static_cast<size_t>(h + w + 4) * sizeof(float);
  CAL_REQUIRE(smem <= 48 * 1024, CAL_E_UNSUPPORTED, "cal_kp_decode: h+w too large");
  const bool vec = (w % 4 == 0) && (reinterpret_cast<uintptr_t>(logp) % 16 == 0) && w <= 1024;
  if (vec && w <= 512) {
    kp_decode_vec_kernel<4><<<planes, KP_THREADS, smem, st>>>(logp, C, h, w, H_img, W_img, out);
  } else if (vec) {
    kp_decode_vec_kernel<8><<<planes, KP_THREADS, smem, st>>>(logp, C, h, w, H_img, W_img, out);
  } else {
    kp_decode_generic_kernel<<<planes, KP_THREADS, smem, st>>>(logp, C, h, w, H_img, W_img, out);
  }
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_line_decode(const float* heat, int B, int C, int h, int w, double sigma,
                               float scale, float* out, void* stream) {
  using namespace cal;
  CAL_REQUIRE(B >= 0 && C >= 0 && h >= 1 && w >= 1, CAL_E_INVALID, "cal_line_decode: bad shape");
  if (B == 0 || C == 0) return CAL_OK;
  CAL_REQUIRE(heat && out, CAL_E_INVALID, "cal_line_decode: null pointer");
  CAL_REQUIRE(static_cast<long long>(h) * w < (1ll << 30), CAL_E_UNSUPPORTED, "cal_line_decode: plane too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 2.0 * sigma ** 2 is a python float (fp64) that torch casts to fp32 for the division
  const float denom = static_cast<float>(2.0 * (sigma * sigma));
  const bool vec = (w % 4 == 0) && (reinterpret_cast<uintptr_t>(heat) % 16 == 0);
  if (vec)
    line_decode_kernel<true><<<B * C, LN_THREADS, 0, st>>>(heat, h, w, denom, scale, out);
  else
    line_decode_kernel<false><<<B * C, LN_THREADS, 0, st>>>(heat, h, w, denom, scale, out);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}
