// conv3x3_pair.cu - 3x3 convolution, stride 1 or 2 (+folded BN)(+residual)(+ReLU) on CTA PAIRS (tcgen05 cta_group::2)
// for the layers whose weights do not fit one CTA's shared memory but whose HALF does: C = 96 at H/8 x W/8
// (128 launches a step).  In conv3x3.cu those layers stream their weights through a ring and spend as long in
// the ring's barrier hand-overs as in MMAs.  Here:
//
//   * two CTAs of a cluster (the two SMs of a TPC) work on two different 128-pixel tiles at a time; ONE thread
//     of the leader CTA issues M = 256 MMAs (cta_group::2): rows 0-127 are the leader's tile, 128-255 the peer's,
//     each CTA's tensor core reading the A operand from its own shared memory and accumulating in its own TMEM;
//   * the B operand (N = Cout weight rows of a (tap, 64-channel chunk) slice) is split between the two: each CTA
//     holds rows [rank * Cout/2, (rank + 1) * Cout/2) of EVERY slice - 9 * ncc * Cout/2 * 128 B = 108 KB for
//     C = 96 - resident for the whole kernel: no weight ring, no weight barriers, half the MMA instructions;
//   * input patches as in conv3x3.cu (halo tile, nine row-shifted views of one SWIZZLE_128B tile); each CTA's
//     TMA producer loads its own patches but signals the LEADER's barrier (cp.async.bulk.tensor ... cta_group::2),
//     tcgen05.commit multicasts the "patch consumed" / "accumulator complete" arrivals to both CTAs, and the peer's
//     epilogue warps arrive on the leader's "accumulator free" barrier across the cluster.
#include <stdlib.h>

#define CAL_TU "conv3x3_pair.cu"
#include "common.cuh"

namespace cal {
namespace {

constexpr int P_THREADS = 352;            // warp 0 TMA producer, warps 1 and 10 MMA issuers (leader CTA), warps 2..9 epilogue
constexpr int P_MMA2_WARP = 10;
constexpr int P_STAGES = 4;               // input patch ring (at most; stride 2: two stages of four phase patches)
constexpr int P_ACC = 4;                  // accumulator stages in TMEM at most (512 columns / N tile)
constexpr int P_MAX_B = 6;
constexpr int P_TWP = 32, P_TW = 30, P_R = 4;
constexpr uint32_t P_PEER_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's leader CTA

struct PairParams {
  int B, H, W, Cout_pad;
  int tiles_x, tiles_y, total_tiles, items;
  int rows, half_rows, ncc, nk_last;   // rows = N of an MMA (the N tile); half_rows of every slice live in each CTA
  int rows_total, n_tiles;             // weight rows per slice in the tensor; N tiles (Cout = n_tiles * rows)
  int resident;                        // all half-slices resident, or streamed through a ring of `b_stages` stages of G slices
  int G;                               // weight slices per ring stage (3 or 1): one barrier pair per stage
  int stride;                          // 1, or 2: the patch is loaded as its four (row, column) parity phases
  int a_stages, phase_bytes;
  int b_stages, b_stage_bytes, n_acc, acc_stride;
  int relu;
  int dual;                // second MMA-issuing warp (alternate items); needs 2 * ncc <= P_STAGES: a parity wait is only meaningful within one pass of the ring
  int a_stage_bytes, half_bytes;
  uint32_t a_tx;
  const float* bias;
  const __half* res;
  __half* y;
};

// ---- cta_group::2 flavours of the PTX wrappers in common.cuh
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's smem] * B[half of the rows in each CTA's smem]^T
__device__ __forceinline__ void umma2_f16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, {%6, %6, %6, %6, %6, %6, %6, %6}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrives on the mbarrier at this offset in every CTA of `mask` once the MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// TMA loads into this CTA's shared memory whose bytes complete on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & P_PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & P_PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the mbarrier at this offset in the leader CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & P_PEER_MASK) : "memory");
}

struct PTile { int b, y0, x0, n0; bool real; };
// item -> (N tile, pair of pixel tiles); this CTA takes pixel tile 2 * (it / n_tiles) + rank
__device__ __forceinline__ PTile p_tile(const PairParams& p, int it, int rank) {
  PTile c;
  const int nt = it % p.n_tiles;
  const int t = 2 * (it / p.n_tiles) + rank;
  c.n0 = nt * p.rows;
  c.real = t < p.total_tiles;
  const int per = p.tiles_x * p.tiles_y;
  const int b = t / per, rem = t - b * per;
  const int ty = rem / p.tiles_x;
  c.b = c.real ? b : p.B;                 // past the last tile: a patch outside the tensor (TMA fills zeros), nothing stored
  c.y0 = ty * P_R;
  c.x0 = (rem - ty * p.tiles_x) * P_TW;
  return c;
}

__device__ __forceinline__ uint32_t p_pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(P_THREADS, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = sA + p.a_stages * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (p.resident ? 9 * p.ncc * p.half_bytes : p.b_stages * p.b_stage_bytes));
  uint64_t* wfull = bars;                  // (the leader's is used)
  uint64_t* fullA = wfull + 1;             // (the leader's)
  uint64_t* emptyA = fullA + P_STAGES;     // each CTA's own
  uint64_t* tfull = emptyA + P_STAGES;     // each CTA's own
  uint64_t* tempty = tfull + P_ACC;        // (the leader's: both CTAs' epilogues arrive)
  uint64_t* fullB = tempty + P_ACC;        // (the leader's)
  uint64_t* emptyB = fullB + P_MAX_B;      // each CTA's own
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(emptyB + P_MAX_B);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    mbar_init(wfull, 1);
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int a = 0; a < P_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
    for (int s = 0; s < P_MAX_B; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.Cout_pad; i += P_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.0f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer's barriers exist before anything is sent their way
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      if (p.resident) {
        if (leader) mbar_expect_tx(wfull, 2u * 9u * static_cast<uint32_t>(p.ncc) * static_cast<uint32_t>(p.half_bytes));
        for (int s = 0; s < 9 * p.ncc; ++s)
          tma_load_2d_pair(sW + s * p.half_bytes, &tmW, wfull, 0, s * p.rows_total + static_cast<int>(rank) * p.half_rows);
      }
      int sa = 0, sb = 0, g = 0;
      uint32_t pha = 0, phb = 0;
      for (int it = pair; it < p.items; it += n_pairs) {
        const PTile tc = p_tile(p, it, static_cast<int>(rank));
        for (int cc = 0; cc < p.ncc; ++cc) {
          mbar_wait(&emptyA[sa], pha ^ 1);
          if (leader) mbar_expect_tx(&fullA[sa], 2u * p.a_tx);
          if (p.stride == 1) {
            tma_load_4d_pair(sA + sa * p.a_stage_bytes, &tmA, &fullA[sa], cc * 64, tc.x0 - 1, tc.y0 - 1, tc.b);
          } else {
            // stride 2: the four (row parity, column parity) phases of the input patch, each a dense 5 x 32-pixel tile
            // (tensor-map element strides of 2); phase 1 = odd rows / columns starts one input pixel before the tile
#pragma unroll
            for (int ph4 = 0; ph4 < 4; ++ph4)
              tma_load_4d_pair(sA + sa * p.a_stage_bytes + ph4 * p.phase_bytes, &tmA, &fullA[sa], cc * 64,
                               2 * tc.x0 - (ph4 & 1), 2 * tc.y0 - (ph4 >> 1), tc.b);
          }
          if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
          if (!p.resident) {
            // this CTA's half of the nine (tap, chunk) slices, p.G to a ring stage, in the order the issuer consumes them
            for (int t9 = 0; t9 < 9; ++t9) {
              if (g == 0) {
                mbar_wait(&emptyB[sb], phb ^ 1);
                if (leader) mbar_expect_tx(&fullB[sb], 2u * static_cast<uint32_t>(p.G) * static_cast<uint32_t>(p.half_bytes));
              }
              tma_load_2d_pair(sW + sb * p.b_stage_bytes + g * p.half_bytes, &tmW, &fullB[sb], 0,
                               (t9 * p.ncc + cc) * p.rows_total + tc.n0 + static_cast<int>(rank) * p.half_rows);
              if (++g == p.G) { g = 0; if (++sb == p.b_stages) { sb = 0; phb ^= 1; } }
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == P_MMA2_WARP) {
    // -------------------------------------------------------------- MMA issuers (leader CTA only)
    // two warps take alternate items (own patch stages and accumulator stage each): while one sits in a barrier
    // round trip the other keeps the tensor pipes of both SMs fed (conv3x3.cu, tools/gpu_mma_pattern.py)
    if (leader && (warp == 1 || p.dual)) {
      const int iw = warp == 1 ? 0 : 1;
      const bool issuer = elect_one();
      const uint32_t idesc = make_idesc_f16(256, p.rows);
      const uint64_t desc0 = make_smem_desc(0, 128, 2);
      const uint32_t dhi = static_cast<uint32_t>(desc0 >> 32), dlo = static_cast<uint32_t>(desc0);
      const uint32_t a_lo0 = dlo + ((smem_u32(sA) & 0x3FFFF) >> 4), a_step = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
      const uint32_t w_lo0 = dlo + ((smem_u32(sW) & 0x3FFFF) >> 4), w_step = static_cast<uint32_t>(p.half_bytes) >> 4;
      uint32_t tap_off[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap % 3;
        if (p.stride == 1) {
          tap_off[tap] = static_cast<uint32_t>((dy * P_TWP + dx) * 128) >> 4;
        } else {
          // input row 2 y - 1 + dy: dy = 1 is row y of the even phase, dy = 0 / 2 rows y / y + 1 of the odd phase
          const int phase = (dy != 1 ? 2 : 0) + (dx != 1 ? 1 : 0);
          tap_off[tap] = static_cast<uint32_t>(phase * p.phase_bytes + ((dy == 2 ? P_TWP : 0) + (dx == 2 ? 1 : 0)) * 128) >> 4;
        }
      }
      const int ncc = p.ncc, nk_last = p.nk_last;
      if (p.resident) { mbar_wait(wfull, 0); tc_fence_after(); }
      const bool resident = p.resident != 0;
      const uint32_t b_stage_step = static_cast<uint32_t>(p.b_stage_bytes) >> 4;
      const int b_stages = p.b_stages, n_acc = p.n_acc;
      int sa = 0, as = 0, sb = 0, g = 0;
      uint32_t pha = 0, aph = 0, phb = 0;
      int li = 0;
      for (int it = pair; it < p.items; it += n_pairs, ++li) {
        if (p.dual && (li & 1) != iw) {
          // the other issuer's item: step the rings past it
          for (int cc = 0; cc < ncc; ++cc)
            if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
          if (++as == n_acc) { as = 0; aph ^= 1; }
          continue;
        }
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * p.acc_stride;
        for (int cc = 0; cc < ncc; ++cc) {
          mbar_wait(&fullA[sa], pha);
          tc_fence_after();
          if (!resident) {
            const int nk = (cc == ncc - 1) ? nk_last : 4;
            const uint32_t a_lo = a_lo0 + sa * a_step;
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              if (g == 0) { mbar_wait(&fullB[sb], phb); tc_fence_after(); }
              if (issuer) {
                const uint32_t at = a_lo + tap_off[t9];
                const uint32_t bl = w_lo0 + sb * b_stage_step + g * w_step;
                umma2_f16_lo(d_tmem, at, bl, dhi, idesc, (cc | t9) != 0);
                if (nk > 1) umma2_f16_lo(d_tmem, at + 2, bl + 2, dhi, idesc, 1);
                if (nk > 2) umma2_f16_lo(d_tmem, at + 4, bl + 4, dhi, idesc, 1);
                if (nk > 3) umma2_f16_lo(d_tmem, at + 6, bl + 6, dhi, idesc, 1);
              }
              if (++g == p.G) {
                if (issuer) umma2_commit_mcast(&emptyB[sb], 0b11);     // both CTAs' halves of the stage are consumed
                g = 0;
                if (++sb == b_stages) { sb = 0; phb ^= 1; }
              }
            }
            if (issuer) umma2_commit_mcast(&emptyA[sa], 0b11);
          } else if (issuer) {
            const int nk = (cc == ncc - 1) ? nk_last : 4;     // pad lanes of the last chunk are zero: skip them
            const uint32_t a_lo = a_lo0 + sa * a_step;
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              const uint32_t at = a_lo + tap_off[t9];
              const uint32_t bl = w_lo0 + (t9 * ncc + cc) * w_step;
              umma2_f16_lo(d_tmem, at, bl, dhi, idesc, (cc | t9) != 0);
              if (nk > 1) umma2_f16_lo(d_tmem, at + 2, bl + 2, dhi, idesc, 1);
              if (nk > 2) umma2_f16_lo(d_tmem, at + 4, bl + 4, dhi, idesc, 1);
              if (nk > 3) umma2_f16_lo(d_tmem, at + 6, bl + 6, dhi, idesc, 1);
            }
            umma2_commit_mcast(&emptyA[sa], 0b11);          // both CTAs' patches are consumed
          }
          __syncwarp();
          if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
        }
        if (issuer) umma2_commit_mcast(&tfull[as], 0b11);   // both CTAs' accumulators are complete
        __syncwarp();
        if (++as == n_acc) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp < P_MMA2_WARP) {
    // ---------------------------------------------------------------- epilogue (both CTAs, own tile)
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int m = quarter * 32 + lane;
    const int r = m / P_TWP, xx = m - r * P_TWP;
    const uint32_t bias_u = smem_u32(s_bias);
    int li = 0;
    for (int it = pair; it < p.items; it += n_pairs, ++li) {
      if ((li & 1) != grp) continue;
      const int as = li % p.n_acc;
      const uint32_t aph = static_cast<uint32_t>(li / p.n_acc) & 1u;
      const PTile tc = p_tile(p, it, static_cast<int>(rank));
      const int y = tc.y0 + r, x = tc.x0 + xx;
      const bool valid = tc.real && xx < P_TW && y < p.H && x < p.W;
      const size_t pix = (static_cast<size_t>(tc.real ? tc.b : 0) * p.H + (valid ? y : 0)) * p.W + (valid ? x : 0);
      const __half* rrow = (p.res && valid) ? p.res + pix * p.Cout_pad + tc.n0 : nullptr;
      __half* yrow = p.y + pix * p.Cout_pad + tc.n0;
      uint4 rnext[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) rnext[q] = make_uint4(0, 0, 0, 0);
      // N <= 96: the whole residual row of this pixel goes in flight before the accumulator is waited for (a group-ahead
      // prefetch hides a few hundred cycles of an HBM latency of well over a thousand)
      const bool pre_all = p.rows <= 96;
      uint4 rall[12];
#pragma unroll
      for (int q = 0; q < 12; ++q) rall[q] = make_uint4(0, 0, 0, 0);
      if (rrow) {
        if (pre_all) {
#pragma unroll
          for (int q = 0; q < 12; ++q)
            if (q * 8 < p.rows) rall[q] = ldg_nc_v4(rrow + q * 8);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) rnext[q] = ldg_nc_v4(rrow + q * 8);
        }
      }
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * p.acc_stride + (static_cast<uint32_t>(quarter * 32) << 16);
      const int real_groups = p.rows >> 5;                  // 32-column groups of real channels (+ one 16-column group when Cout % 32 == 16)
      const bool tail16 = (p.rows & 16) != 0;
      for (int g = 0; g < real_groups; ++g) {
        uint4 rq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { rq[q] = rnext[q]; rnext[q] = make_uint4(0, 0, 0, 0); }
        if (pre_all) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rq[q] = g == 0 ? rall[q] : (g == 1 ? rall[4 + q] : rall[8 + q]);
        } else if (rrow && g + 1 < real_groups) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rnext[q] = ldg_nc_v4(rrow + (g + 1) * 32 + q * 8);
        }
        uint32_t acc[32];
        tmem_ld32(taddr + g * 32, acc);
        tmem_ld_wait();
        if (g + 1 == real_groups && !tail16) {
          // the last columns are in registers: hand the accumulator stage back to the leader's issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty[as]);
        }
        uint32_t o[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t rr[4] = {rq[q].x, rq[q].y, rq[q].z, rq[q].w};
          float4 b0, b1;                                     // bias of channels g*32 + q*8 .. + 8 (explicit ld.shared: 2 instead of 8 loads)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(bias_u + (tc.n0 + g * 32 + q * 8) * 4));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(bias_u + (tc.n0 + g * 32 + q * 8 + 4) * 4));
          const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = q * 8 + 2 * j;
            const __half2 rh = *reinterpret_cast<const __half2*>(&rr[j]);
            float a = __uint_as_float(acc[c]), b2 = __uint_as_float(acc[c + 1]);
            a += bv[2 * j] + __low2float(rh);
            b2 += bv[2 * j + 1] + __high2float(rh);
            if (p.relu) { a = fmaxf(a, 0.0f); b2 = fmaxf(b2, 0.0f); }
            o[q * 4 + j] = p_pack_half2(a, b2);
          }
        }
        if (valid) {
          stg_v8(yrow + g * 32, o);
          stg_v8(yrow + g * 32 + 16, o + 8);
        }
      }
      if (tail16) {
        // a trailing 16-channel group (Cout = 48: N = 48)
        const int c0 = real_groups * 32;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        if (pre_all) { r0 = real_groups == 1 ? rall[4] : rall[8]; r1 = real_groups == 1 ? rall[5] : rall[9]; }
        else if (rrow) { r0 = ldg_nc_v4(rrow + c0); r1 = ldg_nc_v4(rrow + c0 + 8); }
        uint32_t acc[16];
        tmem_ld16(taddr + c0, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty[as]);
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 bq;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bq.x), "=f"(bq.y), "=f"(bq.z), "=f"(bq.w) : "r"(bias_u + (tc.n0 + c0 + q * 4) * 4));
          const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = q * 4 + 2 * j;
            const __half2 rh = *reinterpret_cast<const __half2*>(&rr[q * 2 + j]);
            float a = __uint_as_float(acc[c]) + (bv[2 * j] + __low2float(rh));
            float b2 = __uint_as_float(acc[c + 1]) + (bv[2 * j + 1] + __high2float(rh));
            if (p.relu) { a = fmaxf(a, 0.0f); b2 = fmaxf(b2, 0.0f); }
            o[q * 2 + j] = p_pack_half2(a, b2);
          }
        }
        if (valid) stg_v8(yrow + c0, o);
      }
      if (valid && p.n_tiles == 1) {
        const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};        // pad channels stay zero
        for (int c = p.rows; c < p.Cout_pad; c += 16) stg_v8(yrow + c, z);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves (or frees TMEM) while its peer may still reach into it
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

}  // namespace

// CAL_E_UNSUPPORTED (no error set): the shape is served by conv3x3.cu
int launch_conv3x3_pair(const CalConvArgs* a, void* stream) {
  // bit 0: stride 1 with resident halves, bit 1: stride 1 with streamed halves, bit 2: stride 2
  static const int enabled = [] { const char* e = getenv("CAL_CONV_PAIR"); return e ? atoi(e) : 7; }();
  if (!enabled || a->ksize != 3 || (a->stride != 1 && a->stride != 2) || a->mode != 0 || !a->w_slices) return CAL_E_UNSUPPORTED;
  if (a->stride == 2 && !(enabled & 4)) return CAL_E_UNSUPPORTED;
  if (a->Cout_rows % 16 != 0 || a->Cout_rows < 32 || a->Cout_pad > 1024) return CAL_E_UNSUPPORTED;
  PairParams p{};
  p.B = a->B; p.H = a->Hout; p.W = a->Wout; p.Cout_pad = a->Cout_pad;
  p.rows_total = a->Cout_rows;
  p.ncc = a->Cin_pad / 64;
  p.stride = a->stride;
  {
    const int cin = (a->Cin > 0 && a->Cin <= a->Cin_pad) ? a->Cin : a->Cin_pad;
    p.nk_last = (cin - (p.ncc - 1) * 64 + 15) / 16;
    if (p.nk_last < 1) p.nk_last = 1;
    if (p.nk_last > 4) p.nk_last = 4;
  }
  if (a->stride == 1) {
    p.phase_bytes = 0;
    p.a_stage_bytes = (P_R + 2) * P_TWP * 128 + 1024;               // + pad rows read by the last taps of halo columns
    p.a_tx = static_cast<uint32_t>((P_R + 2) * P_TWP * 128);
    p.a_stages = P_STAGES;
  } else {
    p.phase_bytes = (P_R + 1) * P_TWP * 128;                        // one parity phase of the patch: 5 rows x 32 pixels
    p.a_stage_bytes = 4 * p.phase_bytes + 1024;
    p.a_tx = static_cast<uint32_t>(4 * p.phase_bytes);
    p.a_stages = 2;
  }
  const size_t tail = 64 * 8 + 32 + static_cast<size_t>(a->Cout_pad) * 4;
  const size_t avail = static_cast<size_t>(227 * 1024 - smem_headroom());
  const size_t fixed = 1024 + static_cast<size_t>(p.a_stages) * p.a_stage_bytes + tail;
  // N tile: the whole Cout when it fits one MMA (and a double-buffered accumulator), else its widest divisor
  p.n_tiles = 1;
  while (a->Cout_rows % p.n_tiles != 0 || a->Cout_rows / p.n_tiles > 256 || (a->Cout_rows / p.n_tiles) % 16 != 0) {
    if (++p.n_tiles > 8) return CAL_E_UNSUPPORTED;
  }
  if (p.n_tiles > 1 && (a->Cout_rows != a->Cout_pad || (a->Cout_rows / p.n_tiles) % 32 != 0)) return CAL_E_UNSUPPORTED;
  p.rows = a->Cout_rows / p.n_tiles;
  p.half_rows = p.rows / 2;
  p.half_bytes = p.half_rows * 128;
  if (p.half_bytes % 1024 != 0) return CAL_E_UNSUPPORTED;           // slices stay on swizzle-atom boundaries
  p.acc_stride = (p.rows + 31) & ~31;
  p.n_acc = 512 / p.acc_stride;
  if (p.n_acc > P_ACC) p.n_acc = P_ACC;
  if (p.n_acc < 2) return CAL_E_UNSUPPORTED;
  p.n_acc &= ~1;                                                     // stage parity == epilogue group
  const size_t w_all = static_cast<size_t>(9) * p.ncc * p.half_bytes;
  // stride 1: worth a pair only where one CTA cannot hold the weights (those layers run with resident weights in
  // conv3x3.cu); stride 2: always - the generic kernel fetches every patch nine times (once per tap) and runs at
  // the L2 -> SM cap, the four parity phases here are 80 KB instead of 144 KB per tile and chunk
  if (a->stride == 1 && p.n_tiles == 1 && 2 * w_all + 2 * static_cast<size_t>(p.a_stage_bytes) <= 200 * 1024) return CAL_E_UNSUPPORTED;
  size_t smem;
  p.G = 3;
  if (p.n_tiles == 1 && fixed + w_all <= avail) {
    if (a->stride == 1 && !(enabled & 1)) return CAL_E_UNSUPPORTED;
    p.resident = 1;
    p.b_stages = 0; p.b_stage_bytes = 0;
    smem = fixed + w_all;
  } else {
    if (a->stride == 1 && !(enabled & 2)) return CAL_E_UNSUPPORTED;
    if (fixed >= avail) return CAL_E_UNSUPPORTED;
    p.resident = 0;
    if ((avail - fixed) / (3 * static_cast<size_t>(p.half_bytes)) < 2) p.G = 1;       // wide slices next to the stride-2 patches: one per stage
    p.b_stage_bytes = p.G * p.half_bytes;
    p.b_stages = static_cast<int>((avail - fixed) / p.b_stage_bytes);
    if (p.b_stages > P_MAX_B) p.b_stages = P_MAX_B;
    if (p.b_stages < 2) return CAL_E_UNSUPPORTED;
    smem = fixed + static_cast<size_t>(p.b_stages) * p.b_stage_bytes;
  }
  p.tiles_x = (a->Wout + P_TW - 1) / P_TW;
  p.tiles_y = (a->Hout + P_R - 1) / P_R;
  p.total_tiles = a->B * p.tiles_x * p.tiles_y;
  p.items = (p.total_tiles + 1) / 2 * p.n_tiles;
  p.relu = a->relu;
  { static const bool du = [] { const char* e = getenv("CAL_PAIR_DUAL"); return !(e && e[0] == '0'); }();
    p.dual = (du && p.resident && 2 * p.ncc <= p.a_stages && p.n_acc >= 2) ? 1 : 0; }
  p.bias = a->bias;
  p.res = reinterpret_cast<const __half*>(a->res);
  p.y = reinterpret_cast<__half*>(a->y);

  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin_pad, (uint64_t)a->Win, (uint64_t)a->Hin, (uint64_t)a->B};
    const uint64_t strides[3] = {(uint64_t)a->Cin_pad * 2, (uint64_t)a->Win * a->Cin_pad * 2, (uint64_t)a->Hin * a->Win * a->Cin_pad * 2};
    // stride 2: one parity phase per load - every second pixel of every second row (element strides), 32 x 5 of them
    const uint32_t box1[4] = {64, (uint32_t)P_TWP, (uint32_t)(P_R + 2), 1};
    const uint32_t box2[4] = {64, (uint32_t)(2 * P_TWP), (uint32_t)(2 * (P_R + 1)), 1};
    const uint32_t es2[4] = {1, 2, 2, 1};
    const int rc = encode_tmap_f16(&tmA, a->x, 4, dims, strides, a->stride == 1 ? box1 : box2, a->stride == 1 ? nullptr : es2,
                                   CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  {
    const uint64_t dims[2] = {64ull, 9ull * p.ncc * (uint64_t)a->Cout_rows};   // slice-major: slice s = rows [s * Cout_rows, ...)
    const uint64_t strides[1] = {128ull};
    const uint32_t box[2] = {64, (uint32_t)p.half_rows};
    const int rc = encode_tmap_f16(&tmW, a->w, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    CAL_CHECK_CUDA(cudaGetDevice(&dev));
    CAL_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  {
    static const bool show = getenv("CAL_DEBUG_CONFIG") != nullptr;
    if (show)
      fprintf(stderr, "pair conv s%d %dx%d Cin_pad %d Cout %d: N tile %d x %d resident %d G %d b_stages %d n_acc %d smem %zu items %d dual %d\n", a->stride,
              a->Hout, a->Wout, a->Cin_pad, a->Cout_rows, p.rows, p.n_tiles, p.resident, p.G, p.b_stages, p.n_acc, smem, p.items, p.dual);
  }
  int grid = 2 * p.items < num_sms ? 2 * p.items : (num_sms & ~1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel, tmA, tmW, p));
  return CAL_OK;
}

}  // namespace cal
