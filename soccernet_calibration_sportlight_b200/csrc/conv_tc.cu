// conv_tc.cu - convolution (+folded BN)(+residual)(+ReLU / (Log)Softmax) as an
// implicit GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA.  Replaces the Conv2d+BatchNorm2d(+add)+ReLU groups of
// src/models/hrnet/hrnet.py:42-58, 79-99, 184-213, 260-266, 316-330, 366-388.
//
// GEMM view: D[M=128 pixels][N=Cout] += A[M][K] * W[N][K]^T with K = taps * Cin_pad.
//   * One CTA tile = a TH x TW rectangle of output pixels of one frame (TH*TW <= 128)
//     times up to 256 output channels.
//   * A operand: for every filter tap the (Cin chunk of 64) x TW x TH box of the
//     NHWC fp16 activation tensor is fetched by ONE 4-D TMA load whose start
//     coordinate is shifted by the tap offset; out-of-image elements are zero-filled
//     by the TMA unit (the conv padding), stride-2 convs use the tensor map's
//     element strides.  Rows land as 128-byte SWIZZLE_128B lines = the canonical
//     K-major UMMA layout, so the box IS the MMA operand (no im2col buffer).
//   * B operand: (64 K-elements) x N box of the packed weights [Cout][tap][Cin].
//   * Persistent CTAs (one per SM), static round-robin tile schedule, 3 roles:
//     warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
//     warps 2..9 = epilogue (TMEM -> registers -> bias/residual/ReLU -> fp16 NHWC),
//     mbarrier ring between producer and MMA, double-buffered TMEM accumulators
//     between MMA and epilogue so tile i+1's MMAs overlap tile i's epilogue.
#include <math_constants.h>
#include <stdlib.h>

#define CAL_TU "conv_tc.cu"
#include "common.cuh"

namespace cal {
namespace {

constexpr int CONV_THREADS = 352;          // 11 warps: producer, MMA issuer, 8 epilogue warps, second MMA issuer
constexpr int CONV_MMA2_WARP = 10;
constexpr int A_STAGE_BYTES = 128 * 128;   // 128 rows x 64 fp16
constexpr int KC = 64;                     // K elements per pipeline stage
constexpr int MAX_STAGES = 8;
constexpr int TMEM_COLS = 512;
constexpr int MAX_ACC = 8;                 // TMEM accumulator ring: up to 8 stages (512 columns / N_tile)
constexpr int MAX_BIAS = 1024;

struct ConvParams {
  int B, Hout, Wout, Cout_pad;
  int TW, TH, tiles_x, tiles_y, n_tiles, total_tiles;
  int N_tile;        // output-channel span of one tile in the padded channel space
  int mma_n;         // N of the MMA instruction (<= N_tile, multiple of 16)
  int kc_per_tap, taps, stride, pad, num_k;
  int ksteps_last;   // K = 16 steps of the last channel chunk that hold real channels
  int relu, mode, n_classes;
  int dual;          // two MMA-issuing warps take alternate tiles (the issuing thread's barrier round trips are the
                     // bottleneck of short K loops: csrc/probe.cu, tools/gpu_mma_pattern.py); CAL_TC_DUAL=0 turns it off
  int wait_ahead;    // MMA issuer: next stage's wait before this stage's commit (experiments: CAL_TC_WAIT_AHEAD=0)
  int stages, b_stage_bytes;
  int n_acc, acc_stride;
  uint32_t tx_bytes;
  const float* bias;
  const __half* res;
  void* y;
};

struct TileCoord { int n0, b, y0, x0; };

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t) {
  TileCoord c;
  const int nt = t % p.n_tiles;
  int mt = t / p.n_tiles;
  c.n0 = nt * p.N_tile;
  const int txi = mt % p.tiles_x;
  mt /= p.tiles_x;
  const int tyi = mt % p.tiles_y;
  c.b = mt / p.tiles_y;
  c.y0 = tyi * p.TH;
  c.x0 = txi * p.TW;
  return c;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][barriers][tmem slot][bias]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.stages * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.stages * p.b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tfull = bars + 2 * MAX_STAGES;
  uint64_t* tempty = tfull + MAX_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + MAX_ACC);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < MAX_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  for (int i = threadIdx.x; i < p.Cout_pad; i += CONV_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const int xs = tc.x0 * p.stride - p.pad, ys = tc.y0 * p.stride - p.pad;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? tap / 3 : 0, dx = (p.taps == 9) ? tap - 3 * dy : 0;
          for (int cc = 0; cc < p.kc_per_tap; ++cc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], p.tx_bytes);
            tma_load_4d(sA + stage * A_STAGE_BYTES, &tmA, &full[stage], cc * KC, xs + dx, ys + dy, tc.b);
            tma_load_2d(sB + stage * p.b_stage_bytes, &tmB, &full[stage],
                        (tap * p.kc_per_tap + cc) * KC, tc.n0);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || (warp == CONV_MMA2_WARP && p.dual)) {
    // -------------------------------------------------------------- MMA issuer(s)
    // warp-uniform schedule, one elected lane issues: descriptors stay in uniform registers
    const int iw = warp == 1 ? 0 : 1;
    int stage = 0, as = 0, it = 0;
    uint32_t phase = 0, aphase = 0;
    const uint32_t idesc = make_idesc_f16(128, p.mma_n);
    const uint64_t desc0 = make_smem_desc(0, 128, 2);
    const uint32_t dhi = static_cast<uint32_t>(desc0 >> 32), dlo = static_cast<uint32_t>(desc0);
    const uint32_t a_lo0 = dlo + ((smem_u32(sA) & 0x3FFFF) >> 4), b_lo0 = dlo + ((smem_u32(sB) & 0x3FFFF) >> 4);
    const uint32_t a_step = A_STAGE_BYTES >> 4, b_step = static_cast<uint32_t>(p.b_stage_bytes) >> 4;
    const int num_k = p.num_k, stages = p.stages, kc_per_tap = p.kc_per_tap, ksteps_last = p.ksteps_last, n_acc = p.n_acc;
    const bool issuer = elect_one();
    bool ready = false;          // this stage's full barrier was already waited for
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      if (p.dual && (it & 1) != iw) {
        // the other issuer's tile: step the rings past it
        for (int k = 0; k < num_k; ++k)
          if (++stage == stages) { stage = 0; phase ^= 1; }
        if (++as == n_acc) { as = 0; aphase ^= 1; }
        continue;
      }
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * p.acc_stride;
      int cc_k = 0;
      for (int k = 0; k < num_k; ++k) {
        if (!ready) mbar_wait(&full[stage], phase);
        const uint32_t a_lo = a_lo0 + stage * a_step, b_lo = b_lo0 + stage * b_step;
        if (issuer) {
          // +32 bytes per K=16 step inside the 128-byte swizzle span; steps over pad lanes skipped
          const int nk = (cc_k == kc_per_tap - 1) ? ksteps_last : 4;
          umma_f16_lo(d_tmem, a_lo, b_lo, dhi, idesc, k != 0);
          if (nk > 1) umma_f16_lo(d_tmem, a_lo + 2, b_lo + 2, dhi, idesc, 1);
          if (nk > 2) umma_f16_lo(d_tmem, a_lo + 4, b_lo + 4, dhi, idesc, 1);
          if (nk > 3) umma_f16_lo(d_tmem, a_lo + 6, b_lo + 6, dhi, idesc, 1);
        }
        const int sn = (stage + 1 == stages) ? 0 : stage + 1;
        const uint32_t pn = (sn == 0) ? (phase ^ 1) : phase;
        // A shared-memory access of this thread right behind its own tcgen05.commit stalls ~230
        // cycles (tools/gpu_mma_rate.py): inside a tile the wait for the next stage goes before
        // this stage's commit, while the MMAs just issued are still queued. (Not across tiles: the
        // accumulator-complete commit must not wait for the next tile's loads.)
        ready = p.wait_ahead && k + 1 < num_k;
        if (ready) mbar_wait(&full[sn], pn);
        if (issuer) umma_commit(&empty[stage]);   // frees the smem slot when these MMAs retire
        __syncwarp();
        if (++cc_k == kc_per_tap) cc_k = 0;
        stage = sn; phase = pn;
      }
      if (issuer) umma_commit(&tfull[as]);        // accumulator complete -> epilogue
      __syncwarp();
      if (++as == n_acc) { as = 0; aphase ^= 1; }
    }
  } else if (warp < CONV_MMA2_WARP) {
    // ---------------------------------------------------------------- epilogue
    // Two groups of four warps ping-pong over the tiles (group g owns accumulator stage g), so
    // one tile's TMEM-load / residual-read / store chain overlaps the other group's.
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;
    const int groups_total = (p.N_tile + 31) >> 5;   // 32-column groups
    const int m = quarter * 32 + lane;
    const int ty = m / p.TW, tx = m - ty * p.TW;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int as = it % p.n_acc;
      const uint32_t aphase = static_cast<uint32_t>(it / p.n_acc) & 1u;
      const TileCoord tc = decode_tile(p, t);
      const int y = tc.y0 + ty, x = tc.x0 + tx;
      const bool valid = (m < p.TW * p.TH) && (y < p.Hout) && (x < p.Wout);
      const size_t pix = (static_cast<size_t>(tc.b) * p.Hout + y) * p.Wout + x;
      const __half* rrow = (p.res && valid) ? p.res + pix * p.Cout_pad + tc.n0 : nullptr;
      uint4 rpre[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) rpre[q] = make_uint4(0, 0, 0, 0);
      const bool prefetch_res = groups_total <= 2;
      if (prefetch_res && rrow) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q * 8 < p.N_tile) rpre[q] = __ldg(reinterpret_cast<const uint4*>(rrow) + q);
      }
      // wide tiles: the residual of column group g + 1 is in flight while group g is processed
      // (and group 0's while the accumulator is still being computed)
      uint4 rnext[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) rnext[q] = make_uint4(0, 0, 0, 0);
      if (!prefetch_res && rrow) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q * 8 < p.N_tile) rnext[q] = __ldg(reinterpret_cast<const uint4*>(rrow) + q);
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * p.acc_stride + (static_cast<uint32_t>(quarter * 32) << 16);
      if (p.mode == 0) {
        __half* yrow = reinterpret_cast<__half*>(p.y) + pix * p.Cout_pad + tc.n0;
        for (int g = 0; g < groups_total; ++g) {
          uint4 rq[4];
          if (prefetch_res) {
#pragma unroll
            for (int q = 0; q < 4; ++q) rq[q] = (g == 0) ? rpre[q] : rpre[4 + q];
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) { rq[q] = rnext[q]; rnext[q] = make_uint4(0, 0, 0, 0); }
            if (rrow && g + 1 < groups_total) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if ((g + 1) * 32 + q * 8 < p.N_tile) rnext[q] = __ldg(reinterpret_cast<const uint4*>(rrow + (g + 1) * 32) + q);
            }
          }
          uint32_t acc[32];
          if (g * 32 < p.mma_n) {
            tmem_ld32(taddr + g * 32, acc);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0u;
          }
          if (valid) {
            const uint32_t bb_u = smem_u32(s_bias) + (tc.n0 + g * 32) * 4;
            uint32_t o[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t rr[4] = {rq[q].x, rq[q].y, rq[q].z, rq[q].w};
              const float4 b0 = lds_v4f(bb_u + q * 32), b1 = lds_v4f(bb_u + q * 32 + 16);
              const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = q * 8 + 2 * j;
                const __half2 rh = *reinterpret_cast<const __half2*>(&rr[j]);
                float a = (g * 32 + c < p.mma_n) ? __uint_as_float(acc[c]) : 0.0f;
                float b = (g * 32 + c + 1 < p.mma_n) ? __uint_as_float(acc[c + 1]) : 0.0f;
                a += bv[2 * j] + __low2float(rh);
                b += bv[2 * j + 1] + __high2float(rh);
                if (p.relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                o[q * 4 + j] = pack_half2(a, b);
              }
            }
            // 32 channels = 64 bytes = two full sectors per pixel row
            if (g * 32 + 16 <= p.N_tile) stg_v8(yrow + g * 32, o);
            if (g * 32 + 32 <= p.N_tile) stg_v8(yrow + g * 32 + 16, o + 8);
          }
        }
      } else {
        // (Log)Softmax over the first n_classes columns of this pixel row, fp32 NCHW out
        float v[64];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g * 32 < p.mma_n) {
            uint32_t r[32];
            tmem_ld32(taddr + g * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[g * 32 + j] = (g * 32 + j < p.mma_n) ? __uint_as_float(r[j]) + s_bias[g * 32 + j] : 0.0f;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[g * 32 + j] = 0.0f;
          }
        }
        if (valid) {
          float mx = -CUDART_INF_F;
#pragma unroll
          for (int j = 0; j < 64; ++j) if (j < p.n_classes) mx = fmaxf(mx, v[j]);
          float sum = 0.0f;
#pragma unroll
          for (int j = 0; j < 64; ++j) if (j < p.n_classes) sum += expf(v[j] - mx);
          const float lse = logf(sum), inv = 1.0f / sum;
          float* out = reinterpret_cast<float*>(p.y);
          const size_t plane = static_cast<size_t>(p.Hout) * p.Wout;
          const size_t base = static_cast<size_t>(tc.b) * p.n_classes * plane + static_cast<size_t>(y) * p.Wout + x;
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            if (j < p.n_classes) {
              const float d = v[j] - mx;
              out[base + j * plane] = (p.mode == 1) ? (d - lse) : (expf(d) * inv);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// smem image of one TMA box, for pinning the tensor-map conventions in tests
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int x0, int y0,
                                 int n0, uint32_t bytes, uint4* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < A_STAGE_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes);
    tma_load_4d(smem, &tm, &bar, c0, x0, y0, n0);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < A_STAGE_BYTES / 16; i += blockDim.x) out[i] = reinterpret_cast<uint4*>(smem)[i];
}


// Experiment: is an M-row-shifted start address legal for a SWIZZLE_128B K-major operand?
// A_full = 256 rows x 64 fp16 loaded by one TMA box; D = A_full[shift : shift+128] * W^T.
__global__ void shift_mma_probe_kernel(const __grid_constant__ CUtensorMap tmX,
                                       const __grid_constant__ CUtensorMap tmW, int shift,
                                       int base_offset_mode, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 256 x 128 B
  uint8_t* sB = smem + 256 * 128;     // 64 x 128 B
  __shared__ uint64_t bar, mbar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&mbar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tslot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tslot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 256 * 128 + 64 * 128);
    tma_load_2d(sA, &tmX, &bar, 0, 0);
    tma_load_2d(sB, &tmW, &bar, 0, 0);
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + shift * 128;
    uint64_t a_desc = make_smem_desc(a_addr, 128, 2);
    if (base_offset_mode == 1) a_desc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t b_desc = make_smem_desc(smem_u32(sB), 128, 2);
    const uint32_t idesc = make_idesc_f16(128, 64);
    for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base, a_desc + 2 * kk, b_desc + 2 * kk, idesc, kk != 0);
    umma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tc_fence_after();
  if (warp < 4) {
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int g = 0; g < 4; ++g) {
      uint32_t r[16];
      tmem_ld16(taddr + g * 16, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + g * 16 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// Experiment: MN-major B operand (N contiguous, the layout of an NHWC activation tile used as
// the "weights" of an interpolation GEMM).  D(128x128) = X(128x64, K-major) * Y(64x128, N contiguous).
__global__ void mn_mma_probe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                                    int mode, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 x 128 B
  uint8_t* sB = smem + 128 * 128;     // 2 blocks of [64 K-rows][64 N-elements]
  __shared__ uint64_t bar, mbar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&mbar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tslot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tslot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 128 * 128 + 2 * 64 * 128);
    tma_load_2d(sA, &tmX, &bar, 0, 0);
    tma_load_2d(sB, &tmY, &bar, 0, 0);
    tma_load_2d(sB + 8192, &tmY, &bar, 64, 0);
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t lbo = (mode & 1) ? 1024 : 8192, sbo = (mode & 1) ? 8192 : 1024;
    const uint32_t idesc = make_idesc_f16(128, 128) | (1u << 16);     // B is MN-major
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t a_desc = make_smem_desc(smem_u32(sA), 128, 2) + 2 * kk;
      const uint64_t b_desc = make_smem_desc_ex(smem_u32(sB) + kk * 16 * 128, lbo, sbo, 2);
      umma_f16(tmem_base, a_desc, b_desc, idesc, kk != 0);
    }
    umma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tc_fence_after();
  if (warp < 4) {
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int g = 0; g < 8; ++g) {
      uint32_t r[16];
      tmem_ld16(taddr + g * 16, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 128 + g * 16 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// Experiment: issue rate of tcgen05.mma (M = 128, K = 16, kind::f16) from shared-memory operands
// as a function of N and of the row alignment of the A operand's start address.
__global__ void mma_rate_probe_kernel(int N, int shift_rows, int iters, int flags, long long* out) {
  const int b_mn = flags & 1, fence_too = (flags >> 1) & 1, no_mma = (flags >> 2) & 1, commit_every = (flags >> 8) & 0xFF;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;           // 256 rows x 128 B
  __shared__ uint64_t mbar, mbar2, mbar3;
  __shared__ uint32_t tslot[2];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) tslot[1] = 1u;
  for (int i = threadIdx.x; i < 2 * 256 * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_init(&mbar2, 1u << 20); mbar_init(&mbar3, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tslot[0], 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tslot[0];
  if (warp == 1) {
    const bool issuer = elect_one();
    const uint32_t idesc = make_idesc_f16(128, N) | (b_mn ? (1u << 16) : 0u);
    const uint64_t a0 = make_smem_desc(smem_u32(sA) + shift_rows * 128, 128, 2);
    const uint64_t b0 = b_mn ? make_smem_desc_ex(smem_u32(sB), 8192, 1024, 2) : make_smem_desc(smem_u32(sB), 128, 2);
    long long t0 = 0, t1 = 0;
    if (issuer) {
      // warm-up, then a timed train of dependent-free MMAs into two alternating accumulators
      for (int i = 0; i < 16; ++i) umma_f16(tmem_base, a0 + 2 * (i & 3), b0 + 2 * (i & 3), idesc, i != 0);
      umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after();
    if (issuer) {
      t0 = clock64();
      // groups of 4 MMAs (one 64-wide K slice) into two alternating accumulators; commit_every counts groups
      const uint32_t kadv = b_mn ? 128 : 2;
      const int cmask = commit_every ? commit_every - 1 : 0;      // power of two
      for (int g = 0; g < iters / 4; ++g) {
        const uint32_t d = tmem_base + (g & 1) * 256;
        if (!no_mma) {
          umma_f16(d, a0, b0, idesc, 1);
          umma_f16(d, a0 + 2, b0 + kadv, idesc, 1);
          umma_f16(d, a0 + 4, b0 + 2 * kadv, idesc, 1);
          umma_f16(d, a0 + 6, b0 + 3 * kadv, idesc, 1);
        }
        if (commit_every && (g & cmask) == cmask) {
          if ((flags & 64) && fence_too) mbar_wait(&mbar, 0);   // the wait BEFORE the commit
          umma_commit(&mbar2);                               // never completes: 2^20 arrivals pending
          if (flags & 64) continue;
          if (fence_too) {
            if (flags & 16) {                                // mbarrier.test_wait instead of try_wait
              uint32_t ok = 0;
              while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
            } else if (flags & 32) {                         // a plain shared-memory flag poll
              uint32_t v = 0;
              while (!v) asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&tslot[1])) : "memory");
            } else {
              mbar_wait(&mbar, 0);                           // an already-completed phase
            }
          }
          if (flags & 8) tc_fence_after();
        }
      }
      umma_commit(&mbar);
    }
    mbar_wait(&mbar, 1);
    if (issuer) {
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int make_act_tmap(CUtensorMap* tm, const void* x, int B, int H, int W, int C, int box_w, int box_h,
                  int estride) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[4] = {(uint32_t)KC, (uint32_t)box_w, (uint32_t)box_h, 1};
  const uint32_t es[4] = {1, (uint32_t)estride, (uint32_t)estride, 1};
  return encode_tmap_f16(tm, x, 4, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace
}  // namespace cal

extern "C" int cal_conv2d(const CalConvArgs* a, void* stream) {
  using namespace cal;
  CAL_REQUIRE(a != nullptr, CAL_E_INVALID, "cal_conv2d: null args");
  CAL_REQUIRE(a->x && a->w && a->y, CAL_E_INVALID, "cal_conv2d: null tensor pointer");
  CAL_REQUIRE(a->ksize == 1 || a->ksize == 3, CAL_E_UNSUPPORTED, "cal_conv2d: ksize %d", a->ksize);
  CAL_REQUIRE(a->stride == 1 || a->stride == 2, CAL_E_UNSUPPORTED, "cal_conv2d: stride %d", a->stride);
  CAL_REQUIRE(a->B >= 1 && a->Hin >= 1 && a->Win >= 1 && a->Hout >= 1 && a->Wout >= 1, CAL_E_INVALID,
              "cal_conv2d: bad spatial shape");
  CAL_REQUIRE(a->Cin_pad % KC == 0 && a->Cin_pad >= KC, CAL_E_INVALID, "cal_conv2d: Cin_pad %d not a multiple of 64", a->Cin_pad);
  CAL_REQUIRE(a->Cout_pad % 64 == 0 && a->Cout_pad >= 64 && a->Cout_pad <= MAX_BIAS, CAL_E_INVALID,
              "cal_conv2d: Cout_pad %d", a->Cout_pad);
  CAL_REQUIRE(a->Cout_rows % 16 == 0 && a->Cout_rows >= 16 && a->Cout_rows <= a->Cout_pad, CAL_E_INVALID,
              "cal_conv2d: Cout_rows %d", a->Cout_rows);
  const int pad = a->ksize / 2;
  CAL_REQUIRE(a->Hout == (a->Hin + 2 * pad - a->ksize) / a->stride + 1 &&
                  a->Wout == (a->Win + 2 * pad - a->ksize) / a->stride + 1,
              CAL_E_INVALID, "cal_conv2d: output size %dx%d inconsistent with input %dx%d k%d s%d", a->Hout,
              a->Wout, a->Hin, a->Win, a->ksize, a->stride);
  CAL_REQUIRE(a->mode >= 0 && a->mode <= 2, CAL_E_INVALID, "cal_conv2d: mode %d", a->mode);
  if (a->mode != 0)
    CAL_REQUIRE(a->Cout_pad == 64 && a->n_classes >= 1 && a->n_classes <= a->Cout_rows && !a->res, CAL_E_UNSUPPORTED,
                "cal_conv2d: softmax modes need Cout_pad == 64, n_classes <= Cout_rows, no residual");

  CAL_REQUIRE(!a->w_slices || (a->ksize == 3 && a->mode == 0), CAL_E_INVALID, "cal_conv2d: slice-major weights are for 3x3 convs");
  if (a->ksize == 3 && a->stride == 2 && a->w_slices) {
    // stride-2 3x3 layers: CTA-pair kernel with parity-phase patches (conv3x3_pair.cu); the generic kernel below takes
    // K-major weights
    const int rcp = launch_conv3x3_pair(a, stream);
    if (rcp != CAL_E_UNSUPPORTED) return rcp;
    set_error("cal_conv2d: shape needs the generic kernel, which takes K-major weights");
    return CAL_E_UNSUPPORTED;
  }
  {
    // 3x3 stride-1 layers: halo-tile kernel (conv3x3.cu); CAL_CONV_HALO=0 forces the generic one
    static const bool use_halo = [] { const char* e = getenv("CAL_CONV_HALO"); return !(e && e[0] == '0'); }();
    if (use_halo) {
      const int rcp = launch_conv3x3_pair(a, stream);
      if (rcp != CAL_E_UNSUPPORTED) return rcp;
      const int rc = launch_conv3x3_halo(a, stream);
      if (rc != CAL_E_UNSUPPORTED) return rc;
    }
    CAL_REQUIRE(!a->w_slices, CAL_E_UNSUPPORTED, "cal_conv2d: shape needs the generic kernel, which takes K-major weights");
  }
  ConvParams p{};
  p.B = a->B; p.Hout = a->Hout; p.Wout = a->Wout; p.Cout_pad = a->Cout_pad;
  // tile rectangle: minimise the tile count, prefer wide tiles
  int best_tw = 1, best_th = 1;
  long best = -1;
  for (int tw = 1; tw <= 128 && tw <= a->Wout; ++tw) {
    int th = 128 / tw;
    if (th > a->Hout) th = a->Hout;
    if (th < 1) continue;
    const long tiles = (long)((a->Wout + tw - 1) / tw) * ((a->Hout + th - 1) / th);
    if (best < 0 || tiles < best || (tiles == best && tw > best_tw)) { best = tiles; best_tw = tw; best_th = th; }
  }
  p.TW = best_tw; p.TH = best_th;
  p.tiles_x = (a->Wout + p.TW - 1) / p.TW;
  p.tiles_y = (a->Hout + p.TH - 1) / p.TH;
  // channel tiling in the padded space
  int n_tiles = 1;
  while (a->Cout_pad % (16 * n_tiles) != 0 || a->Cout_pad / n_tiles > 256) ++n_tiles;
  p.n_tiles = n_tiles;
  p.N_tile = a->Cout_pad / n_tiles;
  p.mma_n = p.N_tile < a->Cout_rows ? p.N_tile : a->Cout_rows;
  p.total_tiles = a->B * p.tiles_x * p.tiles_y * n_tiles;
  p.kc_per_tap = a->Cin_pad / KC;
  {
    const int cin = (a->Cin > 0 && a->Cin <= a->Cin_pad) ? a->Cin : a->Cin_pad;
    p.ksteps_last = (cin - (p.kc_per_tap - 1) * KC + 15) / 16;
    if (p.ksteps_last < 1) p.ksteps_last = 1;
    if (p.ksteps_last > 4) p.ksteps_last = 4;
  }
  p.taps = a->ksize * a->ksize;
  p.stride = a->stride; p.pad = pad;
  p.num_k = p.taps * p.kc_per_tap;
  p.relu = a->relu; p.mode = a->mode; p.n_classes = a->n_classes;
  p.bias = a->bias; p.res = reinterpret_cast<const __half*>(a->res); p.y = a->y;
  p.b_stage_bytes = ((p.mma_n * 128) + 1023) & ~1023;
  const int stage_bytes = A_STAGE_BYTES + p.b_stage_bytes;
  p.acc_stride = (p.N_tile + 31) & ~31;
  if (p.mode != 0) p.acc_stride = 64;
  p.n_acc = TMEM_COLS / p.acc_stride;
  if (p.n_acc > MAX_ACC) p.n_acc = MAX_ACC;
  const int tail_bytes = (2 * MAX_STAGES + 2 * MAX_ACC) * 8 + 16 + MAX_BIAS * 4;
  const int avail = 224 * 1024 - smem_headroom() < 200 * 1024 ? 224 * 1024 - smem_headroom() : 200 * 1024;
  int stages = (avail - 1024 - tail_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  CAL_REQUIRE(stages >= 2, CAL_E_UNSUPPORTED, "cal_conv2d: tile does not fit shared memory");
  p.stages = stages;
  // (opt-in, CAL_TC_DUAL=1: the full pipeline failed intermittently with it - unspecified launch failure in the first
  // processes on a box - and the 1x1 convs it applies to are bandwidth-bound anyway)
  { static const bool du = [] { const char* e = getenv("CAL_TC_DUAL"); return e && e[0] == '1'; }(); p.dual = (du && p.n_acc >= 2) ? 1 : 0; }
  // the second issuer waits on barriers up to num_k positions ahead of the first's, and a parity wait is only
  // meaningful within one pass of the ring: short K loops only (the 1x1 convs)
  if (p.num_k + 1 > stages) p.dual = 0;
  { static const bool wa = [] { const char* e = getenv("CAL_TC_WAIT_AHEAD"); return !(e && e[0] == '0'); }(); p.wait_ahead = wa ? 1 : 0; }
  p.tx_bytes = (uint32_t)(p.TW * p.TH * 128 + p.mma_n * 128);
  const size_t smem = 1024 + (size_t)stages * stage_bytes + tail_bytes;

  CUtensorMap tmA, tmB;
  int rc = make_act_tmap(&tmA, a->x, a->B, a->Hin, a->Win, a->Cin_pad, p.TW * a->stride, p.TH * a->stride, a->stride);
  if (rc != CAL_OK) return rc;
  {
    const uint64_t ktot = (uint64_t)p.taps * a->Cin_pad;
    const uint64_t dims[2] = {ktot, (uint64_t)a->Cout_rows};
    const uint64_t strides[1] = {ktot * 2};
    const uint32_t box[2] = {(uint32_t)KC, (uint32_t)p.mma_n};
    rc = encode_tmap_f16(&tmB, a->w, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    CAL_CHECK_CUDA(cudaGetDevice(&dev));
    CAL_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  conv_tc_kernel<<<grid, CONV_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_debug_tma_probe(const void* x, int B, int H, int W, int C, int box_w, int box_h,
                                   int estride, int c0, int x0, int y0, int n0, void* out_smem_16k,
                                   void* stream) {
  using namespace cal;
  CAL_REQUIRE(x && out_smem_16k, CAL_E_INVALID, "cal_debug_tma_probe: null pointer");
  CAL_REQUIRE(C % 64 == 0 && estride >= 1 && estride <= 2, CAL_E_INVALID, "cal_debug_tma_probe: bad args");
  const int nx = (box_w + estride - 1) / estride, ny = (box_h + estride - 1) / estride;
  CAL_REQUIRE(nx * ny <= 128 && nx >= 1 && ny >= 1, CAL_E_INVALID, "cal_debug_tma_probe: box too large");
  CUtensorMap tm;
  int rc = make_act_tmap(&tm, x, B, H, W, C, box_w, box_h, estride);
  if (rc != CAL_OK) return rc;
  CAL_CHECK_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  tma_probe_kernel<<<1, 128, A_STAGE_BYTES + 1024, static_cast<cudaStream_t>(stream)>>>(
      tm, c0, x0, y0, n0, (uint32_t)(nx * ny * 128), reinterpret_cast<uint4*>(out_smem_16k));
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_debug_shift_mma(const void* x_256x64, const void* w_64x64, int shift,
                                   int base_offset_mode, float* out_128x64, void* stream) {
  using namespace cal;
  CAL_REQUIRE(x_256x64 && w_64x64 && out_128x64 && shift >= 0 && shift <= 128, CAL_E_INVALID,
              "cal_debug_shift_mma: bad args");
  CUtensorMap tmX, tmW;
  const uint64_t dx[2] = {64, 256}, dw[2] = {64, 64}, st[1] = {128};
  const uint32_t bx[2] = {64, 256}, bw[2] = {64, 64};
  int rc = encode_tmap_f16(&tmX, x_256x64, 2, dx, st, bx, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != CAL_OK) return rc;
  rc = encode_tmap_f16(&tmW, w_64x64, 2, dw, st, bw, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != CAL_OK) return rc;
  CAL_CHECK_CUDA(cudaFuncSetAttribute(shift_mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  shift_mma_probe_kernel<<<1, 128, 256 * 128 + 64 * 128 + 1024, static_cast<cudaStream_t>(stream)>>>(
      tmX, tmW, shift, base_offset_mode, out_128x64);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_debug_mn_mma(const void* x_128x64, const void* y_64x128, int mode, float* out_128x128,
                                void* stream) {
  using namespace cal;
  CAL_REQUIRE(x_128x64 && y_64x128 && out_128x128, CAL_E_INVALID, "cal_debug_mn_mma: bad args");
  CUtensorMap tmX, tmY;
  const uint64_t dx[2] = {64, 128}, sx[1] = {128}, dy[2] = {128, 64}, sy[1] = {256};
  const uint32_t bx[2] = {64, 128}, by[2] = {64, 64};
  int rc = encode_tmap_f16(&tmX, x_128x64, 2, dx, sx, bx, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != CAL_OK) return rc;
  rc = encode_tmap_f16(&tmY, y_64x128, 2, dy, sy, by, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != CAL_OK) return rc;
  CAL_CHECK_CUDA(cudaFuncSetAttribute(mn_mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  mn_mma_probe_kernel<<<1, 128, 128 * 128 + 2 * 8192 + 1024, static_cast<cudaStream_t>(stream)>>>(tmX, tmY, mode,
                                                                                                  out_128x128);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}

extern "C" int cal_debug_mma_rate(int N, int shift_rows, int iters, int flags, long long* out_cycles, void* stream) {
  using namespace cal;
  CAL_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && shift_rows >= 0 && shift_rows <= 64 && iters >= 1 && out_cycles,
              CAL_E_INVALID, "cal_debug_mma_rate: bad args");
  CAL_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  const int grid = ((flags >> 16) & 0xFF) ? ((flags >> 16) & 0xFF) : 1;
  mma_rate_probe_kernel<<<grid, 64, 2 * 256 * 128 + 1024, static_cast<cudaStream_t>(stream)>>>(N, shift_rows, iters, flags,
                                                                                              out_cycles);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}
