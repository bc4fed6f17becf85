// elementwise.cu - the non-GEMM pieces of the HRNet forward.
//
//  cal_stem_conv    : first 3x3/s2 conv 3->64 (+BN+ReLU) straight from the fp32 NCHW
//                     frame (src/models/hrnet/hrnet.py:450-452).  Cin = 3 gives K = 27,
//                     too thin for the tensor cores: CUDA-core direct conv, two output
//                     pixels per thread, weights broadcast from shared memory.
//  cal_fuse_combine : y = [relu](bias + sum_i up_i(src_i)) over fp16 NHWC tensors, the
//                     multi-resolution exchange of HighResolutionModule.forward
//                     (hrnet.py:229-246) and the head's upsample (hrnet.py:489-509);
//                     bilinear, align_corners=True, coordinates as in
//                     torch.nn.functional.interpolate.  HBM-bound, 128-bit accesses.
#define CAL_TU "elementwise.cu"
#include "common.cuh"

namespace cal {
namespace {

constexpr int STEM_CO = 64;
constexpr int STEM_K = 27;

// Two horizontally adjacent output pixels per thread: every weight quad read from shared memory
// feeds eight FMAs instead of four (the kernel is bound by shared-memory weight reads + FMA issue,
// not by HBM), and a 256-thread block amortises its 7 KB weight load over 512 pixels. Per output the
// FMA order is unchanged (k ascending), so results are bit-identical to the one-pixel version.
// (Four pixels per thread need 215 registers and measured slower: 2.8 vs 1.9 ms per step.)
constexpr int STEM_THREADS = 256;

// U8: the frame as cv2.imread leaves it - uint8 HWC BGR - with ToTensor's float32 division by 255
// (transforms.py:59-68, make_submit.py:66) folded into the load: a quarter of the host->device bytes.
template <bool U8>
__global__ void __launch_bounds__(STEM_THREADS) stem_conv_kernel(const void* __restrict__ xin,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias,
                                                                 __half* __restrict__ y, int B, int H, int W,
                                                                 int Ho, int Wo) {
  __shared__ __align__(16) float s_w[STEM_K * STEM_CO];   // [k][co]
  __shared__ __align__(16) float s_b[STEM_CO];
  for (int i = threadIdx.x; i < STEM_K * STEM_CO; i += blockDim.x) {
    const int k = i / STEM_CO, co = i - k * STEM_CO;
    s_w[i] = w[co * STEM_K + k];
  }
  for (int i = threadIdx.x; i < STEM_CO; i += blockDim.x) s_b[i] = bias[i];
  __syncthreads();
  const int Wp = (Wo + 1) >> 1;                            // pixel pairs per output row
  const long long total = static_cast<long long>(B) * Ho * Wp;
  const long long pr = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pr >= total) return;
  const int ox = 2 * static_cast<int>(pr % Wp);
  const int oy = static_cast<int>((pr / Wp) % Ho);
  const int b = static_cast<int>(pr / (static_cast<long long>(Wp) * Ho));
  const bool two = ox + 1 < Wo;
  // input columns 2*ox-1 .. 2*ox+3 (the two pixels share column 2*ox+1), rows 2*oy-1 .. 2*oy+1
  float in[3][3][5];
  if (U8) {
    const uint8_t* img = static_cast<const uint8_t*>(xin) + static_cast<size_t>(b) * H * W * 3;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy + ky - 1;
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int ix = 2 * ox + kx - 1;
        const bool ok = (iy >= 0) && (iy < H) && (ix >= 0) && (ix < W);
        const uint8_t* px = img + (static_cast<size_t>(iy) * W + ix) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) in[ci][ky][kx] = ok ? __fdiv_rn(static_cast<float>(__ldg(px + ci)), 255.0f) : 0.0f;
      }
    }
  } else {
    const float* x = static_cast<const float*>(xin);
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* plane = x + (static_cast<size_t>(b) * 3 + ci) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const int ix = 2 * ox + kx - 1;
          const bool ok = (iy >= 0) && (iy < H) && (ix >= 0) && (ix < W);
          in[ci][ky][kx] = ok ? __ldg(plane + static_cast<size_t>(iy) * W + ix) : 0.0f;
        }
      }
    }
  }
  __half* out = y + ((static_cast<size_t>(b) * Ho + oy) * Wo + ox) * STEM_CO;
#pragma unroll
  for (int c0 = 0; c0 < STEM_CO; c0 += 16) {
    float acc0[16], acc1[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc0[j] = s_b[c0 + j]; acc1[j] = acc0[j]; }
#pragma unroll
    for (int k = 0; k < STEM_K; ++k) {
      const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
      const float v0 = in[ci][ky][kx], v1 = in[ci][ky][kx + 2];
      const float4* wr = reinterpret_cast<const float4*>(s_w + k * STEM_CO + c0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 ww = wr[q];
        acc0[4 * q + 0] = fmaf(v0, ww.x, acc0[4 * q + 0]);
        acc0[4 * q + 1] = fmaf(v0, ww.y, acc0[4 * q + 1]);
        acc0[4 * q + 2] = fmaf(v0, ww.z, acc0[4 * q + 2]);
        acc0[4 * q + 3] = fmaf(v0, ww.w, acc0[4 * q + 3]);
        acc1[4 * q + 0] = fmaf(v1, ww.x, acc1[4 * q + 0]);
        acc1[4 * q + 1] = fmaf(v1, ww.y, acc1[4 * q + 1]);
        acc1[4 * q + 2] = fmaf(v1, ww.z, acc1[4 * q + 2]);
        acc1[4 * q + 3] = fmaf(v1, ww.w, acc1[4 * q + 3]);
      }
    }
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __half2 h = __floats2half2_rn(fmaxf(acc0[2 * j], 0.0f), fmaxf(acc0[2 * j + 1], 0.0f));
      o[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    stg_v8(out + c0, o);                                   // 16 channels = one full 32-byte sector
    if (two) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __half2 h = __floats2half2_rn(fmaxf(acc1[2 * j], 0.0f), fmaxf(acc1[2 * j + 1], 0.0f));
        o[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      stg_v8(out + STEM_CO + c0, o);
    }
  }
}

struct CombineParams {
  __half* y;
  int B, H, W, C8;       // C8 = C_pad / 8 (uint4 lanes per pixel)
  int C8r;               // lanes that hold real channels, rounded up to the lanes per thread: the rest are written as zeros
  int n_src;
  const __half* src[CAL_MAX_SOURCES];
  int sh[CAL_MAX_SOURCES], sw[CAL_MAX_SOURCES];
  float scale_y[CAL_MAX_SOURCES], scale_x[CAL_MAX_SOURCES];
  const float* bias;
  int relu;
};

__device__ __forceinline__ void acc8(float* a, const uint4 q, float wgt) {
  const uint32_t r[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r[j]));
    a[2 * j] = fmaf(wgt, f.x, a[2 * j]);
    a[2 * j + 1] = fmaf(wgt, f.y, a[2 * j + 1]);
  }
}

// One block per (frame, FC_TY x FC_TX output patch, all channel lanes): the bilinear footprint of
// the patch in every source (a few KB to ~100 KB) stays in L1 while the block walks the patch,
// so each source line crosses the L2 -> SM fabric about once instead of once per output that
// touches it (a row-wise version of this kernel ran at the L2 throughput cap: 16 gathers per
// output).  The kernel is instruction-bound (ncu: 27 % of DRAM peak at 78 % SM throughput; per 16
// output bytes 13 gathers, 104 half->float conversions and 104 FMAs are inherent), so everything
// else is kept off the per-pixel path: a thread owns LANES adjacent 8-channel lanes (16 bytes
// each) for the whole block - its source base pointers and bias are set up once - and walks the
// patch's pixels; pixel coordinates come from shifts (FC_TX = 16), offsets are 32-bit.
constexpr int FC_THREADS = 256;
constexpr int FC_TY = 8, FC_TX = 16;

template <int LANES>
__global__ void __launch_bounds__(FC_THREADS, 4) fuse_combine_kernel(const CombineParams p) {
  const int b = blockIdx.z;
  const int y_base = blockIdx.y * FC_TY, x_base = blockIdx.x * FC_TX;
  const int lanes = p.C8r / LANES;                // threads per pixel (the pad lanes are not worth threads: they would sit
                                                  // masked in every warp - the kernel is bound by instruction issue)
  const int ppb = FC_THREADS / lanes;             // pixels per pass
  const int tid = threadIdx.x;
  if (tid >= ppb * lanes) return;
  const int slot = tid / lanes, c8 = (tid - slot * lanes) * LANES;
  const uint4* base[CAL_MAX_SOURCES];
#pragma unroll
  for (int s = 0; s < CAL_MAX_SOURCES; ++s)
    base[s] = s < p.n_src ? reinterpret_cast<const uint4*>(p.src[s]) + static_cast<size_t>(b) * p.sh[s] * p.sw[s] * p.C8 + c8
                          : nullptr;
  float a0[8 * LANES];
#pragma unroll
  for (int l = 0; l < LANES; ++l) {
    if (p.bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias) + 2 * (c8 + l));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias) + 2 * (c8 + l) + 1);
      a0[8 * l + 0] = b0.x; a0[8 * l + 1] = b0.y; a0[8 * l + 2] = b0.z; a0[8 * l + 3] = b0.w;
      a0[8 * l + 4] = b1.x; a0[8 * l + 5] = b1.y; a0[8 * l + 6] = b1.z; a0[8 * l + 7] = b1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) a0[8 * l + j] = 0.0f;
    }
  }
  uint4* out = reinterpret_cast<uint4*>(p.y) + static_cast<size_t>(b) * p.H * p.W * p.C8 + c8;
  for (int pix = slot; pix < FC_TY * FC_TX; pix += ppb) {
    const int y = y_base + (pix >> 4), x = x_base + (pix & (FC_TX - 1));
    if (y >= p.H || x >= p.W) continue;
    float a[8 * LANES];
#pragma unroll
    for (int j = 0; j < 8 * LANES; ++j) a[j] = a0[j];
#pragma unroll
    for (int s = 0; s < CAL_MAX_SOURCES; ++s) {
      if (s >= p.n_src) break;
      const int sh = p.sh[s], sw = p.sw[s];
      if (sh == p.H && sw == p.W) {
        const uint4* q = base[s] + (y * sw + x) * p.C8;
#pragma unroll
        for (int l = 0; l < LANES; ++l) acc8(a + 8 * l, __ldg(q + l), 1.0f);
      } else {
        // align_corners=True source coordinates (ATen area_pixel_compute_source_index)
        const float fy = p.scale_y[s] * static_cast<float>(y);
        const float fx = p.scale_x[s] * static_cast<float>(x);
        const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
        const int dy = (y0 < sh - 1 ? sw : 0) * p.C8, dx = (x0 < sw - 1 ? p.C8 : 0);
        const float ly1 = fy - static_cast<float>(y0), lx1 = fx - static_cast<float>(x0);
        const float ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
        const float w00 = ly0 * lx0, w01 = ly0 * lx1, w10 = ly1 * lx0, w11 = ly1 * lx1;
        const uint4* q = base[s] + (y0 * sw + x0) * p.C8;
#pragma unroll
        for (int l = 0; l < LANES; ++l) {
          const uint4 q00 = __ldg(q + l), q01 = __ldg(q + dx + l), q10 = __ldg(q + dy + l), q11 = __ldg(q + dy + dx + l);
          acc8(a + 8 * l, q00, w00);
          acc8(a + 8 * l, q01, w01);
          acc8(a + 8 * l, q10, w10);
          acc8(a + 8 * l, q11, w11);
        }
      }
    }
#pragma unroll
    for (int l = 0; l < LANES; ++l) {
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float u = a[8 * l + 2 * j], v = a[8 * l + 2 * j + 1];
        if (p.relu) { u = fmaxf(u, 0.0f); v = fmaxf(v, 0.0f); }
        __half2 h = __floats2half2_rn(u, v);
        o[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      out[(y * p.W + x) * p.C8 + l] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (c8 + LANES == p.C8r) {                    // the pixel's last thread: pad lanes stay zero
      for (int l = LANES; c8 + l < p.C8; ++l) out[(y * p.W + x) * p.C8 + l] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

}  // namespace
}  // namespace cal

static int stem_launch(const void* x, bool u8, const float* w, const float* bias, void* y, int B, int H, int W, int Ho, int Wo,
                       void* stream) {
  using namespace cal;
  CAL_REQUIRE(x && w && bias && y, CAL_E_INVALID, "cal_stem_conv: null pointer");
  CAL_REQUIRE(B >= 1 && H >= 1 && W >= 1 && Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1, CAL_E_INVALID,
              "cal_stem_conv: bad shape %dx%d -> %dx%d", H, W, Ho, Wo);
  const long long total = static_cast<long long>(B) * Ho * ((Wo + 1) / 2);     // pixel pairs
  const int threads = STEM_THREADS;
  const long long blocks = (total + threads - 1) / threads;
  CAL_REQUIRE(blocks < (1ll << 31), CAL_E_UNSUPPORTED, "cal_stem_conv: too many pixels");
  if (u8)
    stem_conv_kernel<true><<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w, bias, reinterpret_cast<__half*>(y), B, H, W, Ho, Wo);
  else
    stem_conv_kernel<false><<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w, bias, reinterpret_cast<__half*>(y), B, H, W, Ho, Wo);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_stem_conv(const float* x, const float* w, const float* bias, void* y, int B,
                             int H, int W, int Ho, int Wo, void* stream) {
  return stem_launch(x, false, w, bias, y, B, H, W, Ho, Wo, stream);
}

extern "C" int cal_stem_conv_u8(const uint8_t* x, const float* w, const float* bias, void* y, int B,
                                int H, int W, int Ho, int Wo, void* stream) {
  return stem_launch(x, true, w, bias, y, B, H, W, Ho, Wo, stream);
}

extern "C" int cal_fuse_combine(const CalCombineArgs* a, void* stream) {
  using namespace cal;
  CAL_REQUIRE(a && a->y, CAL_E_INVALID, "cal_fuse_combine: null args");
  CAL_REQUIRE(a->n_src >= 1 && a->n_src <= CAL_MAX_SOURCES, CAL_E_INVALID, "cal_fuse_combine: n_src %d", a->n_src);
  CAL_REQUIRE(a->C_pad % 8 == 0 && a->B >= 1 && a->H >= 1 && a->W >= 1, CAL_E_INVALID, "cal_fuse_combine: bad shape");
  CombineParams p{};
  p.y = reinterpret_cast<__half*>(a->y);
  p.B = a->B; p.H = a->H; p.W = a->W; p.C8 = a->C_pad / 8;
  p.n_src = a->n_src;
  for (int i = 0; i < a->n_src; ++i) {
    CAL_REQUIRE(a->src[i] && a->src_h[i] >= 1 && a->src_w[i] >= 1, CAL_E_INVALID, "cal_fuse_combine: bad source %d", i);
    p.src[i] = reinterpret_cast<const __half*>(a->src[i]);
    p.sh[i] = a->src_h[i];
    p.sw[i] = a->src_w[i];
    // fp32 scale exactly as ATen's area_pixel_compute_scale<float>(in, out, align_corners=true)
    p.scale_y[i] = a->H > 1 ? static_cast<float>(a->src_h[i] - 1) / static_cast<float>(a->H - 1) : 0.0f;
    p.scale_x[i] = a->W > 1 ? static_cast<float>(a->src_w[i] - 1) / static_cast<float>(a->W - 1) : 0.0f;
  }
  p.bias = a->bias;
  p.relu = a->relu;
  CAL_REQUIRE(a->B <= 65535 && (a->H + FC_TY - 1) / FC_TY <= 65535 &&
                  static_cast<long long>(a->H) * a->W < (1ll << 30),
              CAL_E_UNSUPPORTED, "cal_fuse_combine: tensor too large");
  for (int i = 0; i < a->n_src; ++i)
    CAL_REQUIRE(static_cast<long long>(a->src_h[i]) * a->src_w[i] < (1ll << 30), CAL_E_UNSUPPORTED,
                "cal_fuse_combine: source %d too large", i);
  const dim3 grid(static_cast<unsigned>((a->W + FC_TX - 1) / FC_TX), static_cast<unsigned>((a->H + FC_TY - 1) / FC_TY),
                  static_cast<unsigned>(a->B));
  CAL_REQUIRE(static_cast<long long>(a->H) * a->W * p.C8 < (1ll << 31) && p.C8 <= FC_THREADS, CAL_E_UNSUPPORTED,
              "cal_fuse_combine: tensor too large for 32-bit offsets");
  for (int i = 0; i < a->n_src; ++i)
    CAL_REQUIRE(static_cast<long long>(a->src_h[i]) * a->src_w[i] * p.C8 < (1ll << 31), CAL_E_UNSUPPORTED,
                "cal_fuse_combine: source %d too large for 32-bit offsets", i);
  {
    const int c = (a->C > 0 && a->C <= a->C_pad) ? a->C : a->C_pad;
    const int per = p.C8 % 2 == 0 ? 2 : 1;          // lanes per thread
    p.C8r = ((c + 7) / 8 + per - 1) / per * per;
  }
  if (p.C8 % 2 == 0)
    fuse_combine_kernel<2><<<grid, FC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else
    fuse_combine_kernel<1><<<grid, FC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}
