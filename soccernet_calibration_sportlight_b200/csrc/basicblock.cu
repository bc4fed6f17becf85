// basicblock.cu - a whole BasicBlock (src/models/hrnet/hrnet.py:29-58: conv3x3 - BN - ReLU - conv3x3 - BN - add - ReLU)
// of the full-resolution branch (C <= 48 channels at H/4 x W/4: 136 of the step's conv launches, HBM-bound when
// run one conv at a time) in ONE kernel: the intermediate tensor never reaches HBM.
//
//   * a persistent CTA walks a vertical strip of 28 output columns down the image, 4 rows a step.  Per step
//     conv1 produces a 4 x 32-pixel chunk of the intermediate (with its one-pixel halo) from a 6 x 32-pixel view
//     of the input, conv2 produces 4 x 28 output pixels from a 6-row view of the intermediate: the halo-tile
//     formulation of conv3x3.cu (GEMM M index = flat pixel index of the patch, the nine taps are row-shifted views
//     of one SWIZZLE_128B tile, filter-row grouping: taps dx = 1 | 0 in one N = 2 * Cout MMA, dx = 2 into the same
//     columns, the remaining one-pixel shift a warp shuffle in the epilogue) - same MMAs in the same order as the
//     two separate launches, so the results are bit-identical to them;
//   * input and intermediate live in shared-memory rings of three 4-row chunks; a step's view covers its chunk and
//     the first two rows (+ 2 pixels) of the next, which is physically contiguous except behind the last slot:
//     there a 72-row mirror holds a copy of the head of slot 0 (a second small TMA load / a second store of the
//     first epilogue), so no view ever wraps;
//   * conv1's epilogue (TMEM -> bias, ReLU, zero outside the image = conv2's padding -> fp16) writes the chunk in the
//     K-major SWIZZLE_128B form the second conv's A descriptors read (fence.proxy.async + mbarrier hand-over, as
//     head.cu's chained tail);
//   * roles: warp 0 TMA producer, warp 1 issues conv1's MMAs, warp 2 conv2's (two issuing threads keep the
//     tensor pipe fed while one of them sits in a barrier round trip: tools/gpu_mma_pattern.py), warps 4-7
//     epilogue of conv1, warps 8-11 epilogue of conv2 (+ residual from the block input, ReLU, 32-byte-sector
//     stores); both convs' weights (2 x 9 x Cout x 128 B) stay resident.
#include <stdlib.h>

#define CAL_TU "basicblock.cu"
#include "common.cuh"

namespace cal {
namespace {

constexpr int BB_THREADS = 384;
constexpr int BB_TWP = 32, BB_TW = 28, BB_R = 4;
constexpr int BB_CH = BB_R * BB_TWP * 128;          // one ring chunk: 4 rows x 32 pixels x 64 channels fp16
constexpr int BB_MIR_ROWS = 72;                     // pixel rows mirrored behind the last slot (a view overruns its chunk by 2 * 32 + 2)
constexpr int BB_MIR = BB_MIR_ROWS * 128;
constexpr int BB_SLOTS = 3;
constexpr int BB_RING = BB_SLOTS * BB_CH + BB_MIR;  // 57 KB
constexpr int BB_ACC = 128;                         // TMEM columns per accumulator stage: [G1 | G0] = 2 * Cout <= 96
constexpr int BB_ACC3 = 160;                        // ... with all three taps of a filter row side by side: [G0 | G1 | G2] <= 144

struct BBParams {
  int B, H, W;
  int tiles_x, strips, K;        // K = ceil(H / 4) output steps per strip
  int nk;                        // K = 16 steps that hold real input channels
  int rows, w_bytes;             // weight rows per tap (Cout rounded to 16); one conv's weights in shared memory
  int relu_out;
  const float* bias1;
  const float* bias2;
  const __half* x;
  __half* y;
  int ablate;                    // profiling experiments only (CAL_DEBUG_ABLATE): 4 = issue no MMAs, 2 = no global stores, 32 = no TMEM loads
  long long* dbg;                // profiling experiments only (CAL_DEBUG_TIMELINE): cycles each role spent waiting, per CTA
};

// mbarrier wait that, in profiling runs, adds the cycles spent in it to the role's counter `idx`
#define BB_WAIT(bar, parity, idx)                                  \
  do {                                                             \
    if (p.dbg) {                                                   \
      const long long t0_ = clock64();                             \
      mbar_wait(bar, parity);                                      \
      wcyc[idx] += clock64() - t0_;                                \
    } else {                                                       \
      mbar_wait(bar, parity);                                      \
    }                                                              \
  } while (0)

__device__ __forceinline__ uint32_t bb_pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}


// The NB 16-channel blocks of this thread's pixel, software-pipelined: the TMEM loads of block cb + 1 are in
// flight while block cb is combined and finished (the TMEM read port delivers 64 B/clk per SM - 768 cycles for a
// step's [G1 | G0] accumulator - and both epilogues share it: it must not idle while an epilogue does arithmetic).
// `loaded()` runs once the last load has landed, `block(cb, v)` gets v = G0[p] + G1[p + 1] (+ G2[p + 2]).
template <int NB, bool TAP3, typename FL, typename FB>
__device__ __forceinline__ void bb_acc_blocks(uint32_t taddr, int ablate, FL&& loaded, FB&& block, long long* ldw = nullptr) {
  const bool no_ld = (ablate & 32) != 0;
  constexpr int G = TAP3 ? 3 : 2;
  uint32_t g[2][G][16];
  if (no_ld) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int c = 0; c < 16; ++c) g[i][gi][c] = 0u;
  }
  if (!no_ld) {
#pragma unroll
    for (int gi = 0; gi < G; ++gi) tmem_ld16(taddr + gi * 16 * NB, g[0][gi]);
  }
#pragma unroll
  for (int cb = 0; cb < NB; ++cb) {
    if (!no_ld) {
      if (ldw) { const long long t0_ = clock64(); tmem_ld_wait(); *ldw += clock64() - t0_; } else tmem_ld_wait();
    }
    if (cb + 1 < NB) {
      if (!no_ld) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi) tmem_ld16(taddr + gi * 16 * NB + (cb + 1) * 16, g[(cb + 1) & 1][gi]);
      }
    } else {
      loaded();
    }
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (TAP3) {
        const float a = __uint_as_float(g[cb & 1][0][c]) + __shfl_down_sync(0xffffffffu, __uint_as_float(g[cb & 1][1][c]), 1);
        v[c] = a + __shfl_down_sync(0xffffffffu, __uint_as_float(g[cb & 1][2][c]), 2);
      } else {
        // accumulator columns [G1 | G0] (conv3x3.cu)
        v[c] = __uint_as_float(g[cb & 1][1][c]) + __shfl_down_sync(0xffffffffu, __uint_as_float(g[cb & 1][0][c]), 1);
      }
    }
    block(cb, v);
  }
}

// NB = Cout rows / 16.  TAP3 (experiment, CAL_BB_TAP3=1): the three taps of a filter row in ONE MMA of N = 3 * Cout
// (accumulator columns [G0 | G1 | G2], out[p] = G0[p] + G1[p + 1] + G2[p + 2]: two shuffles per value) - nine MMAs
// a step instead of eighteen, 830 instead of 1150 tensor-pipe cycles (tools/gpu_mma_pattern.py); TMEM then holds
// two accumulator stages for conv1 and one for conv2.  Slower in this kernel: see cal_basicblock.
template <int NB, bool TAP3>
__global__ void __launch_bounds__(BB_THREADS, 1)
basicblock_kernel(const __grid_constant__ CUtensorMap tmX4, const __grid_constant__ CUtensorMap tmX2,
                  const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const BBParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sIn = smem;
  uint8_t* sMid = sIn + BB_RING;
  uint8_t* sW = sMid + BB_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 2 * p.w_bytes);
  uint64_t* wfull = bars;
  uint64_t* inFull = wfull + 1;
  uint64_t* inEmpty = inFull + BB_SLOTS;
  uint64_t* midFull = inEmpty + BB_SLOTS;
  uint64_t* midEmpty = midFull + BB_SLOTS;
  uint64_t* t1full = midEmpty + BB_SLOTS;
  uint64_t* t1empty = t1full + 2;
  uint64_t* t2full = t1empty + 2;
  uint64_t* t2empty = t2full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t2empty + 2);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [2][64], 16-byte aligned (float4 reads)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX4); prefetch_tmap(&tmX2); prefetch_tmap(&tmX1); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    mbar_init(wfull, 1);
    for (int s = 0; s < BB_SLOTS; ++s) {
      mbar_init(&inFull[s], 1); mbar_init(&inEmpty[s], 1);
      mbar_init(&midFull[s], 4); mbar_init(&midEmpty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&t1full[a], 1); mbar_init(&t1empty[a], 4);
      mbar_init(&t2full[a], 1); mbar_init(&t2empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 3) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 128; i += BB_THREADS) s_bias[i] = i < 64 ? p.bias1[i] : p.bias2[i - 64];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int K = p.K;
  long long wcyc[2] = {0, 0};
  const long long t_begin = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(wfull, 2u * static_cast<uint32_t>(p.w_bytes));
      for (int cv = 0; cv < 2; ++cv)
        for (int s = 0; s < 9; ++s) {
          // the taps of a filter row are stored dx = 1, 0, 2: [1 | 0] is one N = 2 * Cout operand
          const int slot = TAP3 ? s : (s / 3) * 3 + ((s % 3) == 0 ? 1 : ((s % 3) == 1 ? 0 : 2));
          tma_load_2d(sW + cv * p.w_bytes + slot * p.rows * 128, cv ? &tmW2 : &tmW1, wfull, 0, s * p.rows);
        }
      int slot = 0;
      uint32_t ph = 0;
      for (int s = blockIdx.x; s < p.strips; s += gridDim.x) {
        const int b = s / p.tiles_x, x0 = (s - b * p.tiles_x) * BB_TW;
        for (int i = 0; i <= K + 1; ++i) {
          BB_WAIT(&inEmpty[slot], ph ^ 1, 0);
          mbar_expect_tx(&inFull[slot], static_cast<uint32_t>(BB_CH + (slot == 0 ? BB_MIR : 0)));
          tma_load_4d(sIn + slot * BB_CH, &tmX4, &inFull[slot], 0, x0 - 2, 4 * i - 2, b);
          if (slot == 0) {
            // the head of this chunk again behind the last slot: 2 rows + 8 pixels of the third
            tma_load_4d(sIn + BB_SLOTS * BB_CH, &tmX2, &inFull[0], 0, x0 - 2, 4 * i - 2, b);
            tma_load_4d(sIn + BB_SLOTS * BB_CH + 64 * 128, &tmX1, &inFull[0], 0, x0 - 2, 4 * i, b);
          }
          if (++slot == BB_SLOTS) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // -------------------------------------------------------------- MMA issuers: warp 1 conv1, warp 2 conv2
    const int cv = warp - 1;
    uint64_t* full = cv ? midFull : inFull;
    uint64_t* empty = cv ? midEmpty : inEmpty;
    uint64_t* tfull = cv ? t2full : t1full;
    uint64_t* tempty = cv ? t2empty : t1empty;
    const int steps = cv ? K : K + 1;              // conv1 also produces the intermediate rows below the last output rows
    const uint32_t idesc = make_idesc_f16(128, 16 * NB), idesc2 = make_idesc_f16(128, 32 * NB), idesc3 = make_idesc_f16(128, 48 * NB);
    constexpr int ACC = TAP3 ? BB_ACC3 : BB_ACC;
    const int n_acc = (TAP3 && cv) ? 1 : 2;
    const uint64_t desc0 = make_smem_desc(0, 128, 2);
    const uint32_t dhi = static_cast<uint32_t>(desc0 >> 32), dlo = static_cast<uint32_t>(desc0);
    const uint32_t a_lo0 = dlo + ((smem_u32(cv ? sMid : sIn) & 0x3FFFF) >> 4);
    const uint32_t w_lo0 = dlo + ((smem_u32(sW + cv * p.w_bytes) & 0x3FFFF) >> 4);
    const uint32_t w_tap = static_cast<uint32_t>(p.rows * 128) >> 4;
    const uint32_t d0 = tmem_base + cv * 2 * ACC;
    const int nk = p.nk;
    const bool issuer = elect_one();
    const bool do_mma = !(p.ablate & 4);
    mbar_wait(wfull, 0);
    tc_fence_after();
    int slot = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int s = blockIdx.x; s < p.strips; s += gridDim.x) {
      BB_WAIT(&full[slot], ph, 0);                  // the strip's first chunk
      for (int j = 0; j < steps; ++j) {
        const int sn = (slot + 1 == BB_SLOTS) ? 0 : slot + 1;
        const uint32_t pn = (sn == 0) ? (ph ^ 1) : ph;
        BB_WAIT(&full[sn], pn, 0);                  // the view runs two rows into the next chunk
        BB_WAIT(&tempty[as], aph ^ 1, 1);
        tc_fence_after();
        if (issuer) {
          const uint32_t d = d0 + as * ACC;
          const uint32_t a_lo = a_lo0 + slot * (BB_CH >> 4);
          if (do_mma)
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t at = a_lo + ((dy * BB_TWP * 128) >> 4), at1 = at + (128 >> 4);
            const uint32_t bl = w_lo0 + dy * 3 * w_tap, bl2 = bl + 2 * w_tap;
            if (TAP3) {
              umma_f16_lo(d, at, bl, dhi, idesc3, dy != 0);
              if (nk > 1) umma_f16_lo(d, at + 2, bl + 2, dhi, idesc3, 1);
              if (nk > 2) umma_f16_lo(d, at + 4, bl + 4, dhi, idesc3, 1);
              if (nk > 3) umma_f16_lo(d, at + 6, bl + 6, dhi, idesc3, 1);
            } else {
              umma_f16_lo(d, at, bl, dhi, idesc2, dy != 0);
              if (nk > 1) umma_f16_lo(d, at + 2, bl + 2, dhi, idesc2, 1);
              if (nk > 2) umma_f16_lo(d, at + 4, bl + 4, dhi, idesc2, 1);
              if (nk > 3) umma_f16_lo(d, at + 6, bl + 6, dhi, idesc2, 1);
              umma_f16_lo(d, at1, bl2, dhi, idesc, 1);
              if (nk > 1) umma_f16_lo(d, at1 + 2, bl2 + 2, dhi, idesc, 1);
              if (nk > 2) umma_f16_lo(d, at1 + 4, bl2 + 4, dhi, idesc, 1);
              if (nk > 3) umma_f16_lo(d, at1 + 6, bl2 + 6, dhi, idesc, 1);
            }
          }
          umma_commit(&tfull[as]);
          umma_commit(&empty[slot]);               // this chunk is not read again
          if (j == steps - 1) umma_commit(&empty[sn]);   // nor is the strip's last one
        }
        __syncwarp();
        if (++as == n_acc) { as = 0; aph ^= 1; }
        slot = sn; ph = pn;
      }
      if (++slot == BB_SLOTS) { slot = 0; ph ^= 1; }   // past the strip's last chunk
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------- epilogue of conv1: accumulator -> intermediate chunk
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;               // pixel of the chunk: row = quarter, column = lane
    const uint32_t sMid_u = smem_u32(sMid), bias_u = smem_u32(s_bias);
    long long phc[5] = {0, 0, 0, 0, 0};      // profiling: cycles in tcgen05.wait::ld, fence+midEmpty wait, loads+math+stores, fence.proxy.async, arrive
    int slot = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int s = blockIdx.x; s < p.strips; s += gridDim.x) {
      const int b = s / p.tiles_x, x0 = (s - b * p.tiles_x) * BB_TW;
      const int mx = x0 - 1 + lane;
      const bool col_in = lane < BB_TW + 2 && mx >= 0 && mx < p.W;
      for (int j = 0; j <= K; ++j) {
        const int my = 4 * j - 1 + quarter;
        const bool inside = col_in && my >= 0 && my < p.H;     // outside the image the intermediate is conv2's zero padding
        BB_WAIT(&t1full[as], aph, 0);
        const long long s0 = p.dbg ? clock64() : 0;
        tc_fence_after();
        constexpr int ACC = TAP3 ? BB_ACC3 : BB_ACC;
        const uint32_t taddr = tmem_base + as * ACC + (static_cast<uint32_t>(quarter * 32) << 16);
        // the chunk's slot must have been read for the last time (conv2 three steps back)
        BB_WAIT(&midEmpty[slot], ph ^ 1, 1);
        const long long s1 = p.dbg ? clock64() : 0;
        const uint32_t row = sMid_u + slot * BB_CH + m * 128;
        const uint32_t mrow = sMid_u + BB_SLOTS * BB_CH + m * 128;
        const bool mirror = slot == 0 && m < BB_MIR_ROWS;
        bb_acc_blocks<NB, TAP3>(taddr, p.ablate, [&] {
          tc_fence_before();                                     // accumulator drained
          __syncwarp();
          if (lane == 0) mbar_arrive(&t1empty[as]);
        }, [&](int cb, const float (&v)[16]) {
          uint32_t o[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bq = lds_v4f(bias_u + (cb * 16 + j4 * 4) * 4);
            const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int jj = j4 * 2 + h2;
              float a = v[2 * jj], c = v[2 * jj + 1];
              a += bv[2 * h2] + 0.0f;                          // (the separate launch adds bias + residual(= 0) the same way)
              c += bv[2 * h2 + 1] + 0.0f;
              a = fmaxf(a, 0.0f); c = fmaxf(c, 0.0f);
              o[jj] = inside ? bb_pack_half2(a, c) : 0u;
            }
          }
          const uint4 lo = make_uint4(o[0], o[1], o[2], o[3]), hi = make_uint4(o[4], o[5], o[6], o[7]);
          sts_v4(row + (((2 * cb) ^ (m & 7)) << 4), lo);
          sts_v4(row + (((2 * cb + 1) ^ (m & 7)) << 4), hi);
          if (mirror) {
            sts_v4(mrow + (((2 * cb) ^ (m & 7)) << 4), lo);
            sts_v4(mrow + (((2 * cb + 1) ^ (m & 7)) << 4), hi);
          }
        }, p.dbg ? &phc[0] : nullptr);
        const long long s2 = p.dbg ? clock64() : 0;
        fence_proxy_async();                          // chunk -> visible to the tensor core's shared-memory reads
        const long long s3 = p.dbg ? clock64() : 0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&midFull[slot]);
        if (p.dbg) { const long long s4 = clock64(); phc[1] += s1 - s0; phc[2] += s2 - s1; phc[3] += s3 - s2; phc[4] += s4 - s3; }
        if (++slot == BB_SLOTS) { slot = 0; ph ^= 1; }
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
    if (p.dbg && warp == 4 && lane == 0)
      for (int i = 0; i < 5; ++i) p.dbg[148 * 15 + blockIdx.x * 5 + i] = phc[i];
  } else if (warp >= 8) {
    // ---------------------------------------------------------------- epilogue of conv2: + bias + block input, ReLU, store
    const int quarter = warp & 3;
    const uint32_t bias_u = smem_u32(s_bias);
    int as = 0;
    uint32_t aph = 0;
    for (int s = blockIdx.x; s < p.strips; s += gridDim.x) {
      const int b = s / p.tiles_x, x0 = (s - b * p.tiles_x) * BB_TW;
      const int x = x0 + lane;
      const bool col_ok = lane < BB_TW && x < p.W;
      for (int k = 0; k < K; ++k) {
        const int y = 4 * k + quarter;
        const bool valid = col_ok && y < p.H;
        const size_t pix = (static_cast<size_t>(b) * p.H + y) * p.W + x;
        uint4 rdx[2 * NB];
#pragma unroll
        for (int q = 0; q < 2 * NB; ++q) rdx[q] = make_uint4(0, 0, 0, 0);
        if (valid) {
          const __half* rrow = p.x + pix * 64;
#pragma unroll
          for (int q = 0; q < 2 * NB; ++q) rdx[q] = ldg_nc_v4(rrow + q * 8);
        }
        BB_WAIT(&t2full[as], aph, 0);
        tc_fence_after();
        constexpr int ACC = TAP3 ? BB_ACC3 : BB_ACC;
        const uint32_t taddr = tmem_base + (2 + as) * ACC + (static_cast<uint32_t>(quarter * 32) << 16);
        __half* yrow = p.y + pix * 64;
        bb_acc_blocks<NB, TAP3>(taddr, p.ablate, [&] {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t2empty[as]);
        }, [&](int cb, const float (&v)[16]) {
          uint32_t o[8];
          const uint4 r0 = rdx[2 * cb], r1 = rdx[2 * cb + 1];
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bq = lds_v4f(bias_u + (64 + cb * 16 + j4 * 4) * 4);
            const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int jj = j4 * 2 + h2;
              float a = v[2 * jj], c = v[2 * jj + 1];
              const __half2 rh = *reinterpret_cast<const __half2*>(&rr[jj]);
              a += bv[2 * h2] + __low2float(rh);
              c += bv[2 * h2 + 1] + __high2float(rh);
              if (p.relu_out) { a = fmaxf(a, 0.0f); c = fmaxf(c, 0.0f); }
              o[jj] = bb_pack_half2(a, c);
            }
          }
          if (valid && !(p.ablate & 2)) stg_v8(yrow + cb * 16, o);
        });
        if (valid) {
          const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};      // pad channels stay zero
#pragma unroll
          for (int cb = NB; cb < 4; ++cb) stg_v8(yrow + cb * 16, z);
        }
        if (TAP3) { aph ^= 1; } else { as ^= 1; if (as == 0) aph ^= 1; }
      }
    }
  }

  if (p.dbg && lane == 0 && (warp <= 2 || warp == 4 || warp == 8)) {
    // per CTA: [role][wait 0, wait 1, total]; roles: producer, MMA1, MMA2, epilogue 1, epilogue 2
    const int role = warp <= 2 ? warp : (warp == 4 ? 3 : 4);
    long long* d = p.dbg + (static_cast<size_t>(blockIdx.x) * 5 + role) * 3;
    d[0] = wcyc[0]; d[1] = wcyc[1]; d[2] = clock64() - t_begin;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace cal

extern "C" int cal_basicblock(const CalBasicBlockArgs* a, void* stream) {
  using namespace cal;
  CAL_REQUIRE(a != nullptr, CAL_E_INVALID, "cal_basicblock: null args");
  CAL_REQUIRE(a->x && a->w1 && a->w2 && a->bias1 && a->bias2 && a->y, CAL_E_INVALID, "cal_basicblock: null tensor pointer");
  CAL_REQUIRE(a->B >= 1 && a->H >= 1 && a->W >= 1, CAL_E_INVALID, "cal_basicblock: bad shape");
  if (a->C_pad != 64 || !(a->rows == 16 || a->rows == 32 || a->rows == 48)) {
    set_error("cal_basicblock: C_pad %d / %d weight rows (needs C_pad 64, Cout <= 48)", a->C_pad, a->rows);
    return CAL_E_UNSUPPORTED;
  }
  BBParams p{};
  p.B = a->B; p.H = a->H; p.W = a->W;
  p.tiles_x = (a->W + BB_TW - 1) / BB_TW;
  p.strips = a->B * p.tiles_x;
  p.K = (a->H + BB_R - 1) / BB_R;
  {
    const int cin = (a->C > 0 && a->C <= 64) ? a->C : 64;
    p.nk = (cin + 15) / 16;
    if (p.nk * 16 > a->rows) p.nk = a->rows / 16;     // (Cin == Cout: the intermediate holds `rows` channels)
  }
  p.rows = a->rows;
  p.w_bytes = 9 * a->rows * 128;
  p.relu_out = 1;
  p.bias1 = a->bias1; p.bias2 = a->bias2;
  p.x = reinterpret_cast<const __half*>(a->x);
  p.y = reinterpret_cast<__half*>(a->y);
  { const char* e = getenv("CAL_DEBUG_ABLATE"); p.ablate = e ? atoi(e) : 0; }
  { const char* e = getenv("CAL_DEBUG_TIMELINE"); p.dbg = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 16)) : nullptr; }
  const size_t smem = 1024 + 2 * static_cast<size_t>(BB_RING) + 2 * static_cast<size_t>(p.w_bytes) + 32 * 8 + 16 + 128 * 4;
  if (smem > static_cast<size_t>(227 * 1024 - smem_headroom())) {
    set_error("cal_basicblock: %zu bytes of shared memory do not fit next to the reserved headroom", smem);
    return CAL_E_UNSUPPORTED;
  }
  CUtensorMap tmX4, tmX2, tmX1, tmW1, tmW2;
  {
    const uint64_t dims[4] = {64ull, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t strides[3] = {128ull, (uint64_t)a->W * 128, (uint64_t)a->H * a->W * 128};
    const uint32_t box4[4] = {64, BB_TWP, 4, 1}, box2[4] = {64, BB_TWP, 2, 1}, box1[4] = {64, 8, 1, 1};
    int rc = encode_tmap_f16(&tmX4, a->x, 4, dims, strides, box4, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
    rc = encode_tmap_f16(&tmX2, a->x, 4, dims, strides, box2, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
    rc = encode_tmap_f16(&tmX1, a->x, 4, dims, strides, box1, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  for (int cv = 0; cv < 2; ++cv) {
    // slice-major weights: slice s = tap, rows [s * rows, (s + 1) * rows) of a (9 * rows, 64) matrix
    const uint64_t dims[2] = {64ull, 9ull * (uint64_t)a->rows};
    const uint64_t strides[1] = {128ull};
    const uint32_t box[2] = {64, (uint32_t)a->rows};
    const int rc = encode_tmap_f16(cv ? &tmW2 : &tmW1, cv ? a->w2 : a->w1, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    CAL_CHECK_CUDA(cudaGetDevice(&dev));
    CAL_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(basicblock_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(basicblock_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(basicblock_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(basicblock_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(basicblock_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(basicblock_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int grid = p.strips < num_sms ? p.strips : num_sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Default: filter-row grouping as conv3x3.cu (N = 2 * Cout + N = Cout per filter row: bit-identical to the two
  // separate launches).  CAL_BB_TAP3=1: all three taps of a filter row in one MMA - 28 % less tensor-pipe time, but
  // the epilogues then read 3 * Cout accumulator columns per pixel through the 64 B/clk TMEM read port, which is
  // what bounds this kernel (tools/gpu_block_waits.py): measured 273 vs 232 us per block at batch 64.
  static const bool tap3 = [] { const char* e = getenv("CAL_BB_TAP3"); return e && e[0] == '1'; }();
#define BB_LAUNCH(NBV, T3) basicblock_kernel<NBV, T3><<<grid, BB_THREADS, smem, st>>>(tmX4, tmX2, tmX1, tmW1, tmW2, p)
  if (tap3) {
    if (a->rows == 16) BB_LAUNCH(1, true); else if (a->rows == 32) BB_LAUNCH(2, true); else BB_LAUNCH(3, true);
  } else {
    if (a->rows == 16) BB_LAUNCH(1, false); else if (a->rows == 32) BB_LAUNCH(2, false); else BB_LAUNCH(3, false);
  }
#undef BB_LAUNCH
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}
