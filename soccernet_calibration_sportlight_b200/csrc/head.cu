// head.cu - the first head convolution of both HRNets without its 784-channel input tensor.
//
// Reference (src/models/hrnet/hrnet.py:489-511, src/models/line/hrnet.py:236-249):
//   z = ReLU(BN(W1 * concat[full, up(y_1), ..., up(y_n)] + b))        (1x1 conv, 784 / 720 channels)
// with up() = bilinear interpolation (align_corners=True) to the head resolution.  A 1x1 conv
// commutes with the interpolation, so with the low-resolution projections p_i = W1_i y_i
// (computed by the ordinary conv kernel, small):
//   z[pix] = ReLU( W1_full * full[pix] + sum_i sum_taps w_tap(pix) * p_i[tap] + b )
// This kernel evaluates the whole bracket as ONE tensor-core accumulation per 8x16-pixel tile:
//   D[128 px][N] = A_full[128][64] * W1_full^T                (K-major operands, TMA)
//                + U_a[128][64] * P_a[64][N] + U_b[128][64] * P_b[64][N]
// where P_* are the source patches under the tile (TMA boxes of the NHWC projections - N is
// the contiguous dimension, i.e. MN-major B operands; convention pinned by
// tools/gpu_mn_probe.py) and U_* hold the (<= 4 per row and source) bilinear weights.  U
// depends only on the tile position, so each persistent CTA builds it once per position and
// sweeps the batch.  Neither the upsampled projections (13.8 GB at batch 64) nor the
// concatenated tensor ever exist.
#define CAL_TU "head.cu"
#include "common.cuh"

namespace cal {
namespace {

constexpr int HD_THREADS = 320;          // warp 0 producer, warp 1 MMA, warps 2..9 epilogue / U builders
constexpr int HD_TY = 8, HD_TX = 16;     // tile: 128 pixels, m = py * 16 + px
constexpr int HD_B_STAGES = 4;
constexpr int HD_B_STAGE_BYTES = 32 * 1024;   // 4 blocks of [64 rows][128 B]
constexpr int HD_A_BYTES = 16 * 1024;
constexpr int HD_TMEM_COLS = 512;
constexpr int HD_ACC_STRIDE = 256;
constexpr int HD_MAX_BIAS = 1024;
constexpr int HD_MAX_LOW = 4;

struct HeadParams {
  int B, H, W, Cout_pad;
  int tiles_x, tiles_y, n_pos, n_tiles, N_tile, nblk;
  int n_low, n_chunks;
  int low_h[HD_MAX_LOW], low_w[HD_MAX_LOW], fh[HD_MAX_LOW], fw[HD_MAX_LOW];
  int chunk[HD_MAX_LOW], rowoff[HD_MAX_LOW];
  float scale_y[HD_MAX_LOW], scale_x[HD_MAX_LOW];
  uint32_t chunk_tx[2];
  uint32_t w_tx;
  const float* bias;
  __half* z;
};

struct Maps {
  CUtensorMap full, w, low[HD_MAX_LOW];
};

__device__ __forceinline__ uint32_t hd_pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// element (row, col) of a [128][64] fp16 K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t u_offset(int row, int col) {
  return static_cast<uint32_t>(row * 128 + ((((col >> 3) ^ row) & 7) << 4) + ((col & 7) << 1));
}

__global__ void __launch_bounds__(HD_THREADS, 1) head_fused_kernel(const __grid_constant__ Maps maps, const HeadParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                   // HD_B_STAGES x 32 KB
  uint8_t* sFull = sB + HD_B_STAGES * HD_B_STAGE_BYTES;   // 2 x 16 KB
  uint8_t* sU = sFull + 2 * HD_A_BYTES;                  // 2 chunks x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sU + 2 * HD_A_BYTES);
  uint64_t* fullB = bars;
  uint64_t* emptyB = fullB + HD_B_STAGES;
  uint64_t* fullS = emptyB + HD_B_STAGES;
  uint64_t* emptyS = fullS + 2;
  uint64_t* tfull = emptyS + 2;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.full);
    prefetch_tmap(&maps.w);
    for (int i = 0; i < p.n_low; ++i) prefetch_tmap(&maps.low[i]);
    for (int s = 0; s < HD_B_STAGES; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&fullS[s], 1); mbar_init(&emptyS[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, HD_TMEM_COLS);
  for (int i = threadIdx.x; i < p.Cout_pad; i += HD_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.0f;
  // patch rows no TMA box ever writes are multiplied by zero weights: they must be finite
  for (int i = threadIdx.x; i < HD_B_STAGES * HD_B_STAGE_BYTES / 16; i += HD_THREADS)
    reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int steps = 1 + p.n_chunks;                      // K steps per (tile, N tile)

  // epilogue-side per-thread constants
  const int quarter = warp & 3;
  const int grp = (warp - 2) >> 2;
  const int m = quarter * 32 + lane;
  const int py = m >> 4, px = m & 15;

  int sb = 0, ss = 0;                 // ring positions (producer and MMA keep their own copies)
  uint32_t phb = 0, phs = 0;
  uint32_t aphase = 0;                // epilogue group's accumulator phase
  int acc_it = 0;                     // running (tile, N tile) counter -> accumulator stage

  for (int pos = blockIdx.x; pos < p.n_pos; pos += gridDim.x) {
    const int tyi = pos / p.tiles_x, txi = pos - tyi * p.tiles_x;
    const int y0 = tyi * HD_TY, x0 = txi * HD_TX;
    // ---------------------------------------------------------------- U for this position
    // (previous position fully consumed: every epilogue thread has seen the last tfull)
    __syncthreads();
    if (warp >= 2) {
      const int bt = threadIdx.x - 64;                   // 0..255
      for (int i = bt; i < 2 * HD_A_BYTES / 16; i += 256) reinterpret_cast<uint4*>(sU)[i] = make_uint4(0, 0, 0, 0);
      named_bar_sync(3, 256);
      // thread (row, source): rows 0..127, sources interleaved over the two halves
      const int row = bt & 127;
      const int rpy = row >> 4, rpx = row & 15;
      const int y = min(y0 + rpy, p.H - 1), x = min(x0 + rpx, p.W - 1);
      for (int s = (bt >> 7); s < p.n_low; s += 2) {
        const float fy = p.scale_y[s] * static_cast<float>(y), fx = p.scale_x[s] * static_cast<float>(x);
        const int sy0 = static_cast<int>(fy), sx0 = static_cast<int>(fx);
        const int sy1 = sy0 + (sy0 < p.low_h[s] - 1 ? 1 : 0), sx1 = sx0 + (sx0 < p.low_w[s] - 1 ? 1 : 0);
        const float ly1 = fy - static_cast<float>(sy0), lx1 = fx - static_cast<float>(sx0);
        const float ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
        const int ry0 = static_cast<int>(p.scale_y[s] * static_cast<float>(y0));
        const int cx0 = static_cast<int>(p.scale_x[s] * static_cast<float>(x0));
        int col[4] = {(sy0 - ry0) * p.fw[s] + (sx0 - cx0), (sy0 - ry0) * p.fw[s] + (sx1 - cx0),
                      (sy1 - ry0) * p.fw[s] + (sx0 - cx0), (sy1 - ry0) * p.fw[s] + (sx1 - cx0)};
        float wgt[4] = {ly0 * lx0, ly0 * lx1, ly1 * lx0, ly1 * lx1};
        // merge coinciding taps (clamped last row / column)
#pragma unroll
        for (int a = 1; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < a; ++b)
            if (col[a] == col[b] && wgt[a] != 0.0f) { wgt[b] += wgt[a]; wgt[a] = 0.0f; }
        uint8_t* tile = sU + p.chunk[s] * HD_A_BYTES;
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (wgt[a] != 0.0f)
            *reinterpret_cast<__half*>(tile + u_offset(row, p.rowoff[s] + col[a])) = __float2half_rn(wgt[a]);
      }
      fence_proxy_async();
    }
    __syncthreads();

    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer
      if (lane == 0) {
        for (int b = 0; b < p.B; ++b) {
          mbar_wait(&emptyS[ss], phs ^ 1);
          mbar_expect_tx(&fullS[ss], HD_A_BYTES);
          tma_load_4d(sFull + ss * HD_A_BYTES, &maps.full, &fullS[ss], 0, x0, y0, b);
          if (++ss == 2) { ss = 0; phs ^= 1; }
          for (int nt = 0; nt < p.n_tiles; ++nt) {
            const int n0 = nt * p.N_tile;
            // step 0: W1_full rows n0 .. n0 + N_tile (K-major)
            mbar_wait(&emptyB[sb], phb ^ 1);
            mbar_expect_tx(&fullB[sb], p.w_tx);
            tma_load_2d(sB + sb * HD_B_STAGE_BYTES, &maps.w, &fullB[sb], 0, n0);
            if (++sb == HD_B_STAGES) { sb = 0; phb ^= 1; }
            // steps 1..: source patches, 64-channel blocks side by side (MN-major)
            for (int c = 0; c < p.n_chunks; ++c) {
              mbar_wait(&emptyB[sb], phb ^ 1);
              mbar_expect_tx(&fullB[sb], p.chunk_tx[c]);
              for (int s = 0; s < p.n_low; ++s) {
                if (p.chunk[s] != c) continue;
                const int ry0 = static_cast<int>(p.scale_y[s] * static_cast<float>(y0));
                const int cx0 = static_cast<int>(p.scale_x[s] * static_cast<float>(x0));
                for (int blk = 0; blk < p.nblk; ++blk)
                  tma_load_4d(sB + sb * HD_B_STAGE_BYTES + blk * 8192 + p.rowoff[s] * 128, &maps.low[s], &fullB[sb],
                              n0 + blk * 64, cx0, ry0, b);
              }
              if (++sb == HD_B_STAGES) { sb = 0; phb ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // -------------------------------------------------------------- MMA issuer
      const bool issuer = elect_one();
      const uint32_t idesc_k = make_idesc_f16(128, p.N_tile);
      const uint32_t idesc_mn = idesc_k | (1u << 16);
      const uint64_t desc_k = make_smem_desc(0, 128, 2);
      const uint64_t desc_mn = make_smem_desc_ex(0, 8192, 1024, 2);
      const uint32_t u_addr = (smem_u32(sU) & 0x3FFFF) >> 4;
      for (int b = 0; b < p.B; ++b) {
        mbar_wait(&fullS[ss], phs);
        tc_fence_after();
        const uint64_t a_full = desc_k | static_cast<uint64_t>((smem_u32(sFull + ss * HD_A_BYTES) & 0x3FFFF) >> 4);
        for (int nt = 0; nt < p.n_tiles; ++nt, ++acc_it) {
          const int as = acc_it & 1;
          mbar_wait(&tempty[as], ((acc_it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * HD_ACC_STRIDE;
          for (int st = 0; st < steps; ++st) {
            mbar_wait(&fullB[sb], phb);
            tc_fence_after();
            const uint32_t b_addr = (smem_u32(sB + sb * HD_B_STAGE_BYTES) & 0x3FFFF) >> 4;
            if (issuer) {
              if (st == 0) {
                const uint64_t b0 = desc_k | b_addr;
                umma_f16(d_tmem, a_full, b0, idesc_k, 0);
                umma_f16(d_tmem, a_full + 2, b0 + 2, idesc_k, 1);
                umma_f16(d_tmem, a_full + 4, b0 + 4, idesc_k, 1);
                umma_f16(d_tmem, a_full + 6, b0 + 6, idesc_k, 1);
              } else {
                const uint64_t a0 = desc_k | static_cast<uint64_t>(u_addr + (st - 1) * (HD_A_BYTES >> 4));
                const uint64_t b0 = desc_mn | b_addr;
                umma_f16(d_tmem, a0, b0, idesc_mn, 1);
                umma_f16(d_tmem, a0 + 2, b0 + 128, idesc_mn, 1);
                umma_f16(d_tmem, a0 + 4, b0 + 256, idesc_mn, 1);
                umma_f16(d_tmem, a0 + 6, b0 + 384, idesc_mn, 1);
              }
              umma_commit(&emptyB[sb]);
            }
            __syncwarp();
            if (++sb == HD_B_STAGES) { sb = 0; phb ^= 1; }
          }
          if (issuer) umma_commit(&tfull[as]);
          __syncwarp();
        }
        if (issuer) umma_commit(&emptyS[ss]);
        __syncwarp();
        if (++ss == 2) { ss = 0; phs ^= 1; }
      }
    } else {
      // ---------------------------------------------------------------- epilogue
      const int y = y0 + py, x = x0 + px;
      const bool valid = (y < p.H) && (x < p.W);
      const int groups_total = (p.N_tile + 31) >> 5;
      for (int b = 0; b < p.B; ++b) {
        __half* zrow = p.z + ((static_cast<size_t>(b) * p.H + y) * p.W + x) * p.Cout_pad;
        for (int nt = 0; nt < p.n_tiles; ++nt, ++acc_it) {
          if ((acc_it & 1) != grp) continue;
          const int n0 = nt * p.N_tile;
          mbar_wait(&tfull[grp], aphase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + grp * HD_ACC_STRIDE + (static_cast<uint32_t>(quarter * 32) << 16);
          for (int g = 0; g < groups_total; ++g) {
            uint32_t acc[32];
            tmem_ld32(taddr + g * 32, acc);
            tmem_ld_wait();
            if (valid) {
              const float* bb = s_bias + n0 + g * 32;
              uint32_t o[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a = fmaxf(__uint_as_float(acc[2 * j]) + bb[2 * j], 0.0f);
                const float bq = fmaxf(__uint_as_float(acc[2 * j + 1]) + bb[2 * j + 1], 0.0f);
                o[j] = hd_pack_half2(a, bq);
              }
              // 32 channels = 64 bytes = two full 32-byte sectors per pixel row (N_tile is a multiple of 16)
              if (g * 32 + 16 <= p.N_tile) stg_v8(zrow + n0 + g * 32, o);
              if (g * 32 + 32 <= p.N_tile) stg_v8(zrow + n0 + g * 32 + 16, o + 8);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[grp]);
          aphase ^= 1;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, HD_TMEM_COLS);
}


// ----------------------------------------------------------------------------------------------
// Chained variant: the final 1x1 conv (784 -> 58 / 720 -> 23, hrnet.py:325-329) and its
// (Log)Softmax run on the ReLU'd tile while it is still on chip, so z never reaches HBM.
//   GEMM1 (as above) -> D1 in TMEM, 192 channels at a time
//   epilogue: D1 -> bias/ReLU -> fp16 -> shared memory in K-major SWIZZLE_128B form (A operand)
//   GEMM2: D2[128 px][64] += z_tile[128][192] * W2[:, 192-slice]^T      (accumulates over the slices)
//   last epilogue: D2 -> + bias2 -> (Log)Softmax over the classes -> fp32 NCHW
constexpr int HC_NT = 192;                    // channels of z per GEMM1 tile
constexpr int HC_B_STAGES = 3;
constexpr int HC_B_STAGE_BYTES = 24 * 1024;   // W1 slice (192 x 128 B) / 3 patch blocks / 3 W2 blocks
constexpr int HC_A2_BYTES = 3 * HD_A_BYTES;   // z tile: 3 blocks of [128 rows][64 ch]
constexpr int HC_D2_COL = 2 * HC_NT;          // TMEM: D1 stages at 0 and 192, D2 at 384

struct ChainParams {
  HeadParams h;
  const float* bias2;
  float* heat;
  int n_classes, mode;       // mode 1 LogSoftmax, 2 Softmax
};

struct ChainMaps {
  CUtensorMap full, w, w2, low[HD_MAX_LOW];
};

__global__ void __launch_bounds__(HD_THREADS, 1) head_chain_kernel(const __grid_constant__ ChainMaps maps,
                                                                  const ChainParams cp) {
  const HeadParams& p = cp.h;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                      // HC_B_STAGES x 24 KB
  uint8_t* sFull = sB + HC_B_STAGES * HC_B_STAGE_BYTES;      // 16 KB
  uint8_t* sU = sFull + HD_A_BYTES;                          // 2 chunks x 16 KB
  uint8_t* sA2 = sU + 2 * HD_A_BYTES;                        // 2 groups x 48 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA2 + 2 * HC_A2_BYTES);
  uint64_t* fullB = bars;
  uint64_t* emptyB = fullB + HC_B_STAGES;
  uint64_t* fullS = emptyB + HC_B_STAGES;
  uint64_t* emptyS = fullS + 1;
  uint64_t* tfull = emptyS + 1;
  uint64_t* tempty = tfull + 2;
  uint64_t* zfull = tempty + 2;
  uint64_t* zempty = zfull + 2;
  uint64_t* d2full = zempty + 2;
  uint64_t* d2empty = d2full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2empty + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_bias2 = s_bias + HD_MAX_BIAS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.full);
    prefetch_tmap(&maps.w);
    prefetch_tmap(&maps.w2);
    for (int i = 0; i < p.n_low; ++i) prefetch_tmap(&maps.low[i]);
    for (int s = 0; s < HC_B_STAGES; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    mbar_init(fullS, 1); mbar_init(emptyS, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4);
      mbar_init(&zfull[a], 4); mbar_init(&zempty[a], 1);
    }
    mbar_init(d2full, 1); mbar_init(d2empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, HD_TMEM_COLS);
  for (int i = threadIdx.x; i < p.Cout_pad; i += HD_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.0f;
  for (int i = threadIdx.x; i < 64; i += HD_THREADS) s_bias2[i] = cp.bias2 ? cp.bias2[i] : 0.0f;
  for (int i = threadIdx.x; i < HC_B_STAGES * HC_B_STAGE_BYTES / 16; i += HD_THREADS)
    reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = (p.Cout_pad + HC_NT - 1) / HC_NT;

  const int quarter = warp & 3;
  const int grp = (warp - 2) >> 2;
  const int m = quarter * 32 + lane;
  const int py = m >> 4, px = m & 15;

  int sb = 0;                      // B ring position / phase (producer and MMA keep their own copies)
  uint32_t phb = 0;
  uint32_t phS = 0;                // phase of the single full-resolution A buffer
  int unit = 0;                    // running unit counter: D1 tiles and D2 tiles; group = unit & 1
  int d1c = 0;                     // D1 tiles so far: TMEM stage = d1c & 1
  int zc[2] = {0, 0};              // z tiles produced / consumed per group
  int dc = 0;                      // D2 tiles so far

  for (int pos = blockIdx.x; pos < p.n_pos; pos += gridDim.x) {
    const int tyi = pos / p.tiles_x, txi = pos - tyi * p.tiles_x;
    const int y0 = tyi * HD_TY, x0 = txi * HD_TX;
    __syncthreads();
    if (warp >= 2) {
      const int bt = threadIdx.x - 64;
      for (int i = bt; i < 2 * HD_A_BYTES / 16; i += 256) reinterpret_cast<uint4*>(sU)[i] = make_uint4(0, 0, 0, 0);
      named_bar_sync(3, 256);
      const int row = bt & 127;
      const int rpy = row >> 4, rpx = row & 15;
      const int y = min(y0 + rpy, p.H - 1), x = min(x0 + rpx, p.W - 1);
      for (int s = (bt >> 7); s < p.n_low; s += 2) {
        const float fy = p.scale_y[s] * static_cast<float>(y), fx = p.scale_x[s] * static_cast<float>(x);
        const int sy0 = static_cast<int>(fy), sx0 = static_cast<int>(fx);
        const int sy1 = sy0 + (sy0 < p.low_h[s] - 1 ? 1 : 0), sx1 = sx0 + (sx0 < p.low_w[s] - 1 ? 1 : 0);
        const float ly1 = fy - static_cast<float>(sy0), lx1 = fx - static_cast<float>(sx0);
        const float ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
        const int ry0 = static_cast<int>(p.scale_y[s] * static_cast<float>(y0));
        const int cx0 = static_cast<int>(p.scale_x[s] * static_cast<float>(x0));
        int col[4] = {(sy0 - ry0) * p.fw[s] + (sx0 - cx0), (sy0 - ry0) * p.fw[s] + (sx1 - cx0),
                      (sy1 - ry0) * p.fw[s] + (sx0 - cx0), (sy1 - ry0) * p.fw[s] + (sx1 - cx0)};
        float wgt[4] = {ly0 * lx0, ly0 * lx1, ly1 * lx0, ly1 * lx1};
#pragma unroll
        for (int a = 1; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < a; ++b)
            if (col[a] == col[b] && wgt[a] != 0.0f) { wgt[b] += wgt[a]; wgt[a] = 0.0f; }
        uint8_t* tile = sU + p.chunk[s] * HD_A_BYTES;
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (wgt[a] != 0.0f)
            *reinterpret_cast<__half*>(tile + u_offset(row, p.rowoff[s] + col[a])) = __float2half_rn(wgt[a]);
      }
      fence_proxy_async();
    }
    __syncthreads();

    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer
      if (lane == 0) {
        auto load_w2 = [&](int j) {
          const int n0 = j * HC_NT;
          const int nb = (min(HC_NT, p.Cout_pad - n0)) / 64;
          mbar_wait(&emptyB[sb], phb ^ 1);
          mbar_expect_tx(&fullB[sb], static_cast<uint32_t>(nb * 8192));
          for (int kb = 0; kb < nb; ++kb)
            tma_load_2d(sB + sb * HC_B_STAGE_BYTES + kb * 8192, &maps.w2, &fullB[sb], n0 + kb * 64, 0);
          if (++sb == HC_B_STAGES) { sb = 0; phb ^= 1; }
        };
        for (int b = 0; b < p.B; ++b) {
          mbar_wait(emptyS, phS ^ 1);
          mbar_expect_tx(fullS, HD_A_BYTES);
          tma_load_4d(sFull, &maps.full, fullS, 0, x0, y0, b);
          phS ^= 1;
          for (int j = 0; j < n_tiles; ++j) {
            const int n0 = j * HC_NT;
            const int nt_w = min(HC_NT, p.Cout_pad - n0);
            const int nb = nt_w / 64;
            mbar_wait(&emptyB[sb], phb ^ 1);
            mbar_expect_tx(&fullB[sb], static_cast<uint32_t>(nt_w * 128));
            for (int kb = 0; kb < nb; ++kb)       // W1_full rows in 64-row boxes (rows past Cout_rows: zero fill)
              tma_load_2d(sB + sb * HC_B_STAGE_BYTES + kb * 8192, &maps.w, &fullB[sb], 0, n0 + kb * 64);
            if (++sb == HC_B_STAGES) { sb = 0; phb ^= 1; }
            for (int c = 0; c < p.n_chunks; ++c) {
              uint32_t tx = 0;
              for (int s = 0; s < p.n_low; ++s)
                if (p.chunk[s] == c) tx += static_cast<uint32_t>(nb * p.fh[s] * p.fw[s] * 128);
              mbar_wait(&emptyB[sb], phb ^ 1);
              mbar_expect_tx(&fullB[sb], tx);
              for (int s = 0; s < p.n_low; ++s) {
                if (p.chunk[s] != c) continue;
                const int ry0 = static_cast<int>(p.scale_y[s] * static_cast<float>(y0));
                const int cx0 = static_cast<int>(p.scale_x[s] * static_cast<float>(x0));
                for (int blk = 0; blk < nb; ++blk)
                  tma_load_4d(sB + sb * HC_B_STAGE_BYTES + blk * 8192 + p.rowoff[s] * 128, &maps.low[s], &fullB[sb],
                              n0 + blk * 64, cx0, ry0, b);
              }
              if (++sb == HC_B_STAGES) { sb = 0; phb ^= 1; }
            }
            if (j > 0) load_w2(j - 1);
          }
          load_w2(n_tiles - 1);
        }
      }
    } else if (warp == 1) {
      // -------------------------------------------------------------- MMA issuer
      const bool issuer = elect_one();
      const uint64_t desc_k = make_smem_desc(0, 128, 2);
      const uint64_t desc_mn = make_smem_desc_ex(0, 8192, 1024, 2);
      const uint32_t u_addr = (smem_u32(sU) & 0x3FFFF) >> 4;
      const uint64_t a_full = desc_k | static_cast<uint64_t>((smem_u32(sFull) & 0x3FFFF) >> 4);
      const uint32_t idesc2 = make_idesc_f16(128, 64);
      const uint32_t d2_tmem = tmem_base + HC_D2_COL;
      int prev_unit = 0, prev_nb = 0;
      auto gemm2 = [&](int j, int u, int nb) {
        const int g = u & 1;
        mbar_wait(&zfull[g], static_cast<uint32_t>(zc[g] & 1));       // z tile staged by the epilogue
        if (j == 0) mbar_wait(d2empty, static_cast<uint32_t>(dc & 1) ^ 1);   // D2 drained
        mbar_wait(&fullB[sb], phb);                                     // W2 slice landed
        tc_fence_after();
        const uint32_t a2 = (smem_u32(sA2 + g * HC_A2_BYTES) & 0x3FFFF) >> 4;
        const uint32_t b2 = (smem_u32(sB + sb * HC_B_STAGE_BYTES) & 0x3FFFF) >> 4;
        if (issuer) {
          for (int kb = 0; kb < nb; ++kb) {
            const uint64_t ad = desc_k | static_cast<uint64_t>(a2 + kb * (HD_A_BYTES >> 4));
            const uint64_t bd = desc_k | static_cast<uint64_t>(b2 + kb * (8192 >> 4));
            umma_f16(d2_tmem, ad, bd, idesc2, (j | kb) != 0);
            umma_f16(d2_tmem, ad + 2, bd + 2, idesc2, 1);
            umma_f16(d2_tmem, ad + 4, bd + 4, idesc2, 1);
            umma_f16(d2_tmem, ad + 6, bd + 6, idesc2, 1);
          }
          umma_commit(&emptyB[sb]);
          umma_commit(&zempty[g]);
        }
        __syncwarp();
        ++zc[g];
        if (++sb == HC_B_STAGES) { sb = 0; phb ^= 1; }
      };
      for (int b = 0; b < p.B; ++b) {
        mbar_wait(fullS, phS);
        tc_fence_after();
        for (int j = 0; j < n_tiles; ++j) {
          const int n0 = j * HC_NT;
          const int nt_w = min(HC_NT, p.Cout_pad - n0);
          const int u = unit++;
          const int s1 = d1c & 1;
          mbar_wait(&tempty[s1], static_cast<uint32_t>((d1c >> 1) & 1) ^ 1);
          tc_fence_after();
          ++d1c;
          const uint32_t d_tmem = tmem_base + s1 * HC_NT;
          const uint32_t idesc_k = make_idesc_f16(128, nt_w);
          const uint32_t idesc_mn = idesc_k | (1u << 16);
          for (int st = 0; st < 1 + p.n_chunks; ++st) {
            mbar_wait(&fullB[sb], phb);
            tc_fence_after();
            const uint32_t b_addr = (smem_u32(sB + sb * HC_B_STAGE_BYTES) & 0x3FFFF) >> 4;
            if (issuer) {
              if (st == 0) {
                const uint64_t b0 = desc_k | b_addr;
                umma_f16(d_tmem, a_full, b0, idesc_k, 0);
                umma_f16(d_tmem, a_full + 2, b0 + 2, idesc_k, 1);
                umma_f16(d_tmem, a_full + 4, b0 + 4, idesc_k, 1);
                umma_f16(d_tmem, a_full + 6, b0 + 6, idesc_k, 1);
              } else {
                const uint64_t a0 = desc_k | static_cast<uint64_t>(u_addr + (st - 1) * (HD_A_BYTES >> 4));
                const uint64_t b0 = desc_mn | b_addr;
                umma_f16(d_tmem, a0, b0, idesc_mn, 1);
                umma_f16(d_tmem, a0 + 2, b0 + 128, idesc_mn, 1);
                umma_f16(d_tmem, a0 + 4, b0 + 256, idesc_mn, 1);
                umma_f16(d_tmem, a0 + 6, b0 + 384, idesc_mn, 1);
              }
              umma_commit(&emptyB[sb]);
            }
            __syncwarp();
            if (++sb == HC_B_STAGES) { sb = 0; phb ^= 1; }
          }
          if (issuer) umma_commit(&tfull[s1]);
          __syncwarp();
          if (j > 0) gemm2(j - 1, prev_unit, prev_nb);     // overlaps the epilogue of tile j
          prev_unit = u;
          prev_nb = nt_w / 64;
        }
        if (issuer) umma_commit(emptyS);                    // the full-resolution tile is consumed
        __syncwarp();
        phS ^= 1;
        gemm2(n_tiles - 1, prev_unit, prev_nb);
        if (issuer) umma_commit(d2full);
        __syncwarp();
        ++dc;
        ++unit;                                             // the D2 unit
      }
    } else {
      // ---------------------------------------------------------------- epilogue
      const int y = y0 + py, x = x0 + px;
      const bool valid = (y < p.H) && (x < p.W);
      const uint32_t bias_u = smem_u32(s_bias);
      for (int b = 0; b < p.B; ++b) {
        for (int j = 0; j < n_tiles; ++j) {
          const int u = unit++;
          const int s1 = d1c & 1;
          const uint32_t ph1 = static_cast<uint32_t>((d1c >> 1) & 1);
          ++d1c;
          if ((u & 1) != grp) continue;
          const int n0 = j * HC_NT;
          const int nt_w = min(HC_NT, p.Cout_pad - n0);
          mbar_wait(&tfull[s1], ph1);
          mbar_wait(&zempty[grp], static_cast<uint32_t>(zc[grp] & 1) ^ 1);   // GEMM2 done with this buffer
          tc_fence_after();
          ++zc[grp];
          const uint32_t taddr = tmem_base + s1 * HC_NT + (static_cast<uint32_t>(quarter * 32) << 16);
          const uint32_t zt_u = smem_u32(sA2) + grp * HC_A2_BYTES + m * 128;
          for (int g = 0; g < (nt_w >> 5); ++g) {
            uint32_t acc[32];
            tmem_ld32(taddr + g * 32, acc);
            tmem_ld_wait();
            const uint32_t bb_u = bias_u + (n0 + g * 32) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b0 = lds_v4f(bb_u + q * 32), b1 = lds_v4f(bb_u + q * 32 + 16);
              const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint32_t o[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int c = q * 8 + 2 * jj;
                o[jj] = hd_pack_half2(fmaxf(__uint_as_float(acc[c]) + bv[2 * jj], 0.0f),
                                      fmaxf(__uint_as_float(acc[c + 1]) + bv[2 * jj + 1], 0.0f));
              }
              const int chunk = ((g & 1) * 4 + q) ^ (m & 7);
              sts_v4(zt_u + (g >> 1) * HD_A_BYTES + (chunk << 4), make_uint4(o[0], o[1], o[2], o[3]));
            }
          }
          tc_fence_before();
          fence_proxy_async();                        // z tile -> visible to the tensor core's smem reads
          __syncwarp();
          if (lane == 0) { mbar_arrive(&tempty[s1]); mbar_arrive(&zfull[grp]); }
        }
        // the D2 unit of this (position, frame)
        const int u2 = unit++;
        const uint32_t ph2 = static_cast<uint32_t>(dc & 1);
        ++dc;
        if ((u2 & 1) != grp) continue;
        mbar_wait(d2full, ph2);
        tc_fence_after();
        const uint32_t taddr2 = tmem_base + HC_D2_COL + (static_cast<uint32_t>(quarter * 32) << 16);
        float v[64];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t r[32];
          tmem_ld32(taddr2 + g * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[g * 32 + jj] = __uint_as_float(r[jj]) + s_bias2[g * 32 + jj];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2empty);          // D2 is in registers: the next tile may overwrite it
        if (valid) {
          float mx = -INFINITY;
#pragma unroll
          for (int jj = 0; jj < 64; ++jj) if (jj < cp.n_classes) mx = fmaxf(mx, v[jj]);
          float sum = 0.0f;
#pragma unroll
          for (int jj = 0; jj < 64; ++jj) if (jj < cp.n_classes) sum += expf(v[jj] - mx);
          const float lse = logf(sum), inv = 1.0f / sum;
          const size_t plane = static_cast<size_t>(p.H) * p.W;
          float* out = cp.heat + static_cast<size_t>(b) * cp.n_classes * plane + static_cast<size_t>(y) * p.W + x;
#pragma unroll
          for (int jj = 0; jj < 64; ++jj) {
            if (jj < cp.n_classes) {
              const float d = v[jj] - mx;
              out[jj * plane] = (cp.mode == 1) ? (d - lse) : (expf(d) * inv);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, HD_TMEM_COLS);
}

}  // namespace
}  // namespace cal

extern "C" int cal_head_fused(const CalHeadArgs* a, void* stream) {
  using namespace cal;
  CAL_REQUIRE(a && a->full && a->w_full && (a->z || a->w2), CAL_E_INVALID, "cal_head_fused: null pointer");
  if (a->w2)
    CAL_REQUIRE(a->heat && a->n_classes >= 1 && a->n_classes <= 64 && (a->mode == 1 || a->mode == 2), CAL_E_INVALID,
                "cal_head_fused: chained tail needs heat, n_classes <= 64, mode 1|2");
  CAL_REQUIRE(a->n_low >= 1 && a->n_low <= HD_MAX_LOW, CAL_E_INVALID, "cal_head_fused: n_low %d", a->n_low);
  CAL_REQUIRE(a->Cf_pad == 64, CAL_E_UNSUPPORTED, "cal_head_fused: full-resolution source must have 64 (padded) channels");
  CAL_REQUIRE(a->Cout_pad % 64 == 0 && a->Cout_pad >= 64 && a->Cout_pad <= HD_MAX_BIAS && a->Cout_rows % 16 == 0 &&
                  a->Cout_rows <= a->Cout_pad,
              CAL_E_INVALID, "cal_head_fused: Cout_pad %d / Cout_rows %d", a->Cout_pad, a->Cout_rows);
  CAL_REQUIRE(a->B >= 1 && a->H >= 2 && a->W >= 2, CAL_E_INVALID, "cal_head_fused: bad shape");
  HeadParams p{};
  p.B = a->B; p.H = a->H; p.W = a->W; p.Cout_pad = a->Cout_pad;
  p.tiles_x = (a->W + HD_TX - 1) / HD_TX;
  p.tiles_y = (a->H + HD_TY - 1) / HD_TY;
  p.n_pos = p.tiles_x * p.tiles_y;
  int n_tiles = 1;
  while (a->Cout_pad % (16 * n_tiles) != 0 || a->Cout_pad / n_tiles > 256) ++n_tiles;
  p.n_tiles = n_tiles;
  p.N_tile = a->Cout_pad / n_tiles;
  p.nblk = (p.N_tile + 63) / 64;
  p.n_low = a->n_low;
  p.bias = a->bias;
  p.z = reinterpret_cast<__half*>(a->z);
  // bilinear footprints of a tile in every source, and their packing into <= 2 chunks of 64 rows
  int used[2] = {0, 0};
  p.n_chunks = 0;
  for (int s = 0; s < a->n_low; ++s) {
    CAL_REQUIRE(a->low[s] && a->low_h[s] >= 1 && a->low_w[s] >= 1, CAL_E_INVALID, "cal_head_fused: bad source %d", s);
    p.low_h[s] = a->low_h[s]; p.low_w[s] = a->low_w[s];
    // fp32 scale exactly as ATen's area_pixel_compute_scale<float>(in, out, align_corners=true)
    p.scale_y[s] = static_cast<float>(a->low_h[s] - 1) / static_cast<float>(a->H - 1);
    p.scale_x[s] = static_cast<float>(a->low_w[s] - 1) / static_cast<float>(a->W - 1);
    int fh = 1, fw = 1;
    for (int o = 0; o < a->H; o += HD_TY) {
      const int last = (o + HD_TY - 1 < a->H - 1) ? o + HD_TY - 1 : a->H - 1;
      const int lo = static_cast<int>(p.scale_y[s] * static_cast<float>(o));
      int hi = static_cast<int>(p.scale_y[s] * static_cast<float>(last));
      hi += (hi < a->low_h[s] - 1) ? 1 : 0;
      if (hi - lo + 1 > fh) fh = hi - lo + 1;
    }
    for (int o = 0; o < a->W; o += HD_TX) {
      const int last = (o + HD_TX - 1 < a->W - 1) ? o + HD_TX - 1 : a->W - 1;
      const int lo = static_cast<int>(p.scale_x[s] * static_cast<float>(o));
      int hi = static_cast<int>(p.scale_x[s] * static_cast<float>(last));
      hi += (hi < a->low_w[s] - 1) ? 1 : 0;
      if (hi - lo + 1 > fw) fw = hi - lo + 1;
    }
    p.fh[s] = fh; p.fw[s] = fw;
    const int rows = (fh * fw + 7) & ~7;
    int c = -1;
    for (int k = 0; k < 2; ++k)
      if (c < 0 && used[k] + rows <= 64) c = k;
    if (c < 0) {
      set_error("cal_head_fused: source %d footprint %dx%d does not fit the interpolation chunks", s, fh, fw);
      return CAL_E_UNSUPPORTED;
    }
    p.chunk[s] = c; p.rowoff[s] = used[c];
    used[c] += rows;
    if (c + 1 > p.n_chunks) p.n_chunks = c + 1;
  }
  for (int c = 0; c < 2; ++c) {
    uint32_t tx = 0;
    for (int s = 0; s < a->n_low; ++s)
      if (p.chunk[s] == c) tx += static_cast<uint32_t>(p.nblk * p.fh[s] * p.fw[s] * 128);
    p.chunk_tx[c] = tx;
  }
  p.w_tx = static_cast<uint32_t>(p.N_tile * 128);

  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    CAL_CHECK_CUDA(cudaGetDevice(&dev));
    CAL_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(head_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int grid = p.n_pos < num_sms ? p.n_pos : num_sms;

  if (a->w2) {
    // ---- chained variant: 192-channel GEMM1 tiles feeding the final conv on chip
    ChainParams cp{};
    cp.h = p;
    cp.bias2 = a->bias2; cp.heat = a->heat; cp.n_classes = a->n_classes; cp.mode = a->mode;
    ChainMaps cm;
    {
      const uint64_t dims[4] = {64, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
      const uint64_t strides[3] = {128, (uint64_t)a->W * 128, (uint64_t)a->H * a->W * 128};
      const uint32_t box[4] = {64, HD_TX, HD_TY, 1};
      int rc = encode_tmap_f16(&cm.full, a->full, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != CAL_OK) return rc;
    }
    {
      const uint64_t dims[2] = {64, (uint64_t)a->Cout_rows};
      const uint64_t strides[1] = {128};
      const uint32_t box[2] = {64, 64};
      int rc = encode_tmap_f16(&cm.w, a->w_full, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != CAL_OK) return rc;
    }
    {
      const uint64_t dims[2] = {(uint64_t)a->Cout_pad, 64};
      const uint64_t strides[1] = {(uint64_t)a->Cout_pad * 2};
      const uint32_t box[2] = {64, 64};
      int rc = encode_tmap_f16(&cm.w2, a->w2, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != CAL_OK) return rc;
    }
    for (int s = 0; s < a->n_low; ++s) {
      const uint64_t C = (uint64_t)a->Cout_pad;
      const uint64_t dims[4] = {C, (uint64_t)a->low_w[s], (uint64_t)a->low_h[s], (uint64_t)a->B};
      const uint64_t strides[3] = {C * 2, (uint64_t)a->low_w[s] * C * 2, (uint64_t)a->low_h[s] * a->low_w[s] * C * 2};
      const uint32_t box[4] = {64, (uint32_t)p.fw[s], (uint32_t)p.fh[s], 1};
      int rc = encode_tmap_f16(&cm.low[s], a->low[s], 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != CAL_OK) return rc;
    }
    for (int s = a->n_low; s < HD_MAX_LOW; ++s) cm.low[s] = cm.low[0];
    const size_t smem_c = 1024 + HC_B_STAGES * HC_B_STAGE_BYTES + 3 * HD_A_BYTES + 2 * HC_A2_BYTES + 24 * 8 + 16 +
                          (HD_MAX_BIAS + 64) * 4;
    head_chain_kernel<<<grid, HD_THREADS, smem_c, static_cast<cudaStream_t>(stream)>>>(cm, cp);
    CAL_CHECK_CUDA(cudaGetLastError());
    return CAL_OK;
  }

  Maps maps;
  {
    const uint64_t dims[4] = {64, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t strides[3] = {128, (uint64_t)a->W * 128, (uint64_t)a->H * a->W * 128};
    const uint32_t box[4] = {64, HD_TX, HD_TY, 1};
    int rc = encode_tmap_f16(&maps.full, a->full, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  {
    const uint64_t dims[2] = {64, (uint64_t)a->Cout_rows};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, (uint32_t)p.N_tile};
    int rc = encode_tmap_f16(&maps.w, a->w_full, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  for (int s = 0; s < a->n_low; ++s) {
    const uint64_t C = (uint64_t)a->Cout_pad;
    const uint64_t dims[4] = {C, (uint64_t)a->low_w[s], (uint64_t)a->low_h[s], (uint64_t)a->B};
    const uint64_t strides[3] = {C * 2, (uint64_t)a->low_w[s] * C * 2, (uint64_t)a->low_h[s] * a->low_w[s] * C * 2};
    const uint32_t box[4] = {64, (uint32_t)p.fw[s], (uint32_t)p.fh[s], 1};
    int rc = encode_tmap_f16(&maps.low[s], a->low[s], 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  for (int s = a->n_low; s < HD_MAX_LOW; ++s) maps.low[s] = maps.low[0];

  const size_t smem = 1024 + HD_B_STAGES * HD_B_STAGE_BYTES + 4 * HD_A_BYTES + (2 * HD_B_STAGES + 8) * 8 + 16 +
                      HD_MAX_BIAS * 4;
  head_fused_kernel<<<grid, HD_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(maps, p);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}
