// conv3x3.cu - 3x3 stride-1 convolution (+folded BN)(+residual)(+ReLU), the BasicBlock /
// Bottleneck work-horse of HRNet (src/models/hrnet/hrnet.py:29-58, 61-99: 86 % of the conv
// FLOPs outside the head), as an implicit GEMM on tcgen05 with a HALO tile:
//
//   * the input patch of a tile - (R+2) image rows x TWp = TW+2 pixels x 64 channels - is
//     fetched ONCE per 64-channel chunk by a single 4-D TMA load (zero fill outside the image =
//     the conv padding) and lands as (R+2)*TWp consecutive 128-byte SWIZZLE_128B rows;
//   * the GEMM's M dimension is the flat pixel index p = r*TWp + x of that patch pitch
//     (R*TWp = 128), so the A operand of filter tap (dy,dx) is the SAME shared-memory tile
//     read through a descriptor whose start address is advanced by (dy*TWp + dx) rows - nine
//     shifted views instead of nine loads (row-shifted SWIZZLE_128B descriptors are legal: the
//     swizzle is a function of the absolute address; pinned by tools/gpu_shift_probe.py).
//     Two of the TWp columns are halo: 6 % of the MMA rows compute values nobody stores.
//   * weights stay resident in shared memory for the whole (persistent) CTA when they fit
//     (C = 48 / 64), otherwise they stream through their own ring, G (tap, chunk) slices to a
//     stage. The single MMA-issuing thread pays ~400 cycles per barrier wait + tcgen05.commit
//     pair and the tensor pipe queues only a couple of MMAs (tools/gpu_mma_rate.py), so the issue
//     loop is a handful of instructions per MMA, a stage's wait comes before the previous
//     stage's commit, and work items of T = 2 tiles share every slice where TMEM allows;
//   * C <= 64 (the 135x240 branch: a quarter of the step): an MMA of M = 128, K = 16 reads its A operand
//     from shared memory in 32 cycles whatever N is, so N = 48 runs at 24 / 44 of the tensor rate and
//     nine of them per K step re-read the same patch nine times.  Template DX takes the column shift of
//     two taps out of the operand: per filter row dy, ONE MMA with N = 2 * Cout multiplies the view
//     shifted by dy rows with the taps dx = 1 and dx = 0 side by side (accumulator columns [G1 | G0]),
//     and the tap dx = 2 accumulates into the G1 columns from the view shifted one more pixel:
//     G1'[q] = sum_dy W[dy,1] X[q + dy*TWp] + W[dy,2] X[q + 1 + dy*TWp], G0[q] = sum_dy W[dy,0] X[q + dy*TWp],
//     out[p] = G0[p] + G1'[p + 1].  Six MMAs per K step instead of nine (300 instead of 396 operand-read
//     cycles), and the one shift left is a warp shuffle in the epilogue: with TWp <= 32 a patch row lies
//     inside one warp's 32 TMEM lanes.  (All three taps in one N = 3 * Cout MMA was tried first: two
//     shuffles per value and 160 accumulator columns - only two TMEM stages - made the epilogue and the
//     issuer wait for each other: 200 vs 171 us.)
//   * epilogue: TMEM -> registers (32 columns per tcgen05.ld) -> bias / residual / ReLU -> fp16
//     -> 32-byte-sector stores straight from registers (st.global.v8). A swizzled staging tile
//     drained by TMA stores remains as a build option (-DCAL_HALO_STAGED): its shared-memory traffic
//     competes with the MMA operand reads that bound the small-N layers.
//
// L2 -> SM traffic per tile drops from 9 x 16 KB (+ all weights) to 25 KB (+ nothing when the
// weights are resident): the old per-tap kernel ran these layers at the L2 throughput cap.
#include <stdlib.h>

#define CAL_TU "conv3x3.cu"
#include "common.cuh"

namespace cal {
namespace {

constexpr int H_THREADS = 352;            // warp 0 producer, warp 1 MMA, warps 2..9 epilogue, warp 10 second MMA issuer
constexpr int H_MMA2_WARP = 10;
constexpr int H_EPI_THREADS = 256;
constexpr int H_MAX_A_STAGES = 6;
constexpr int H_MAX_B_STAGES = 12;
constexpr int H_TMEM_COLS = 512;
constexpr int H_MAX_ACC = 8;              // TMEM accumulator stages (512 columns / N_tile)
constexpr int H_MAX_BIAS = 1024;
constexpr int H_STAGE_BLOCK = 128 * 128;  // staging: 128 rows x 64 fp16 per channel block
// The staged epilogue (swizzled staging tile drained by TMA stores) is compiled only with
// -DCAL_HALO_STAGED: carrying both store paths costs the resident variants register spills, and the
// direct path measured faster (see launch_conv3x3_halo).
#ifdef CAL_HALO_STAGED
constexpr bool H_STAGED_BUILD = true;
#else
constexpr bool H_STAGED_BUILD = false;
#endif

struct HaloParams {
  int B, H, W, Cout_pad;
  int TWp, TW, R;
  int cluster, part_rows, n_slowest;   // weight slices multicast over `cluster` CTAs, part_rows rows loaded by each
  int T, RI;               // a work item is T (1, 2 or 4) vertically adjacent tiles sharing every weight slice; RI = T*R rows
  int tiles_x, tiles_y, n_tiles, total_tiles;
  int N_tile, mma_n, nblk, ncc;
  int ksteps_last;         // K = 16 steps of the last 64-channel chunk that hold real channels
  int n_acc, acc_stride;   // accumulator ring in TMEM: n_acc stages of acc_stride columns
  int w_slices;            // weights are slice-major: slice s = rows [s*Cout_rows, (s+1)*Cout_rows) of a (.., 64) matrix
  int cout_rows;
  int relu, ablate;     // ablate: profiling experiments only (CAL_DEBUG_ABLATE), 0 in production
  int dual;             // two MMA-issuing warps: 1 = they alternate items (T = 1), 2 = they split the tiles of every item (T > 1)
  int dx;               // filter-row grouping: taps dx = 1, 0 in one N = 2 * mma_n MMA, the column shift in the epilogue (template DX)
  int a_stages, a_stage_bytes, out_bufs;
  uint32_t a_tx_bytes;
  int w_resident, b_stages, b_slice_bytes;
  int G, b_stage_bytes;    // streamed weights: a ring stage holds G consecutive (tap, chunk) slices
  uint32_t b_tx_bytes;
  const float* bias;
  const __half* res;
  __half* yptr;            // output tensor (direct stores of the streamed-weight variant)
  long long* dbg;          // profiling experiments only: per-role clock64 stamps of CTA 0 (CAL_DEBUG_TIMELINE)
};

#define H_STAMP(slot, tile_i, k)                                                              \
  do {                                                                                         \
    if (p.dbg && blockIdx.x == 0 && (tile_i) < 64) p.dbg[((slot) * 64 + (tile_i)) * 8 + (k)] = clock64(); \
  } while (0)

struct HTile { int n0, b, y0, x0; };

// Tile coordinates advanced by a fixed stride without divisions: the stride's mixed-radix digits
// (N tile, x tile, y tile, frame) are computed once, each step is an add with carries.
struct HTileIter {
  int d[4];                     // digits, fastest first: (nt, x, y, b), or (x, y, b, nt) when p.n_slowest
  int s[4];                     // digits of the stride
  int rad[3];                   // radices of the three lower digits
  __device__ __forceinline__ void init(const HaloParams& p, int t, int stride) {
    if (p.n_slowest) { rad[0] = p.tiles_x; rad[1] = p.tiles_y; rad[2] = p.B; }
    else { rad[0] = p.n_tiles; rad[1] = p.tiles_x; rad[2] = p.tiles_y; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { d[k] = t % rad[k]; t /= rad[k]; s[k] = stride % rad[k]; stride /= rad[k]; }
    d[3] = t; s[3] = stride;
  }
  __device__ __forceinline__ void advance(const HaloParams&) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      d[k] += s[k] + c;
      c = 0;
      if (d[k] >= rad[k]) { d[k] -= rad[k]; c = 1; }
    }
    d[3] += s[3] + c;
  }
  __device__ __forceinline__ HTile get(const HaloParams& p) const {
    HTile c;
    if (p.n_slowest) { c.x0 = d[0] * p.TW; c.y0 = d[1] * p.RI; c.b = d[2]; c.n0 = d[3] * p.N_tile; }
    else { c.n0 = d[0] * p.N_tile; c.x0 = d[1] * p.TW; c.y0 = d[2] * p.RI; c.b = d[3]; }
    return c;
  }
};

__device__ __forceinline__ uint32_t h_pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <bool RESIDENT, int T, int DXB = 0>   // DXB: filter-row grouping with mma_n = 16 * DXB output channels (0: off)
__global__ void __launch_bounds__(H_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmBp, const __grid_constant__ CUtensorMap tmY,
                    const HaloParams p) {
  constexpr bool DX = DXB > 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = sA + p.a_stages * p.a_stage_bytes;
  const int w_slots = RESIDENT ? 9 * p.ncc : p.b_stages * p.G;
  uint8_t* sOut = sW + w_slots * p.b_slice_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + p.out_bufs * p.nblk * H_STAGE_BLOCK);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + H_MAX_A_STAGES;
  uint64_t* fullB = emptyA + H_MAX_A_STAGES;
  uint64_t* emptyB = fullB + H_MAX_B_STAGES;
  uint64_t* wfull = emptyB + H_MAX_B_STAGES;
  uint64_t* tfull = wfull + 1;
  uint64_t* tempty = tfull + H_MAX_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + H_MAX_ACC);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // 16-byte aligned (float4 reads)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.dbg && threadIdx.x == 0) {            // per-CTA wall-clock span (entries after the role stamps)
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.dbg[4 * 64 * 8 + 2 * blockIdx.x] = static_cast<long long>(gt);
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmY);
    const uint32_t n_iss = p.dual == 2 ? 2u : 1u;         // issuers committing on every shared stage
    for (int s = 0; s < H_MAX_A_STAGES; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], n_iss); }
    for (int s = 0; s < H_MAX_B_STAGES; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], p.cluster * n_iss); }
    mbar_init(wfull, 1);
    for (int a = 0; a < H_MAX_ACC; ++a) { mbar_init(&tfull[a], n_iss); mbar_init(&tempty[a], T > 1 ? 4 * T : 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, H_TMEM_COLS);
  for (int i = threadIdx.x; i < p.Cout_pad; i += H_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.0f;
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();      // peers' barriers are initialised before any multicast reaches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();                 // the next kernel's prologue may overlap this kernel's tail
  const int items_cta = (p.total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const uint32_t crank = p.cluster > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << p.cluster) - 1u);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (RESIDENT) {
        mbar_expect_tx(wfull, p.b_tx_bytes * 9u * static_cast<uint32_t>(p.ncc));
        for (int s = 0; s < 9 * p.ncc; ++s) {
          // DX (one chunk): the taps of a filter row are stored dx = 1, 0, 2
          const int slot = DX ? (s / 3) * 3 + ((s % 3) == 0 ? 1 : ((s % 3) == 1 ? 0 : 2)) : s;
          if (p.w_slices) tma_load_2d(sW + slot * p.b_slice_bytes, &tmB, wfull, 0, s * p.cout_rows);
          else tma_load_2d(sW + slot * p.b_slice_bytes, &tmB, wfull, s * 64, 0);
        }
      }
      griddep_wait();                           // activations of the previous kernel from here on
      int sa = 0, sb = 0, g = 0;
      uint32_t pha = 0, phb = 0;
      int slices_left = items_cta * 9 * p.ncc;   // of this CTA, over all its items
      HTileIter ti;
      ti.init(p, blockIdx.x, gridDim.x);
      int pi = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ti.advance(p), ++pi) {
        const HTile tc = ti.get(p);
        for (int cc = 0; cc < p.ncc; ++cc) {
          H_STAMP(0, pi, 0);
          mbar_wait(&emptyA[sa], pha ^ 1);
          H_STAMP(0, pi, 1);
          if (p.ablate & 8) {
            mbar_arrive(&fullA[sa]);
          } else {
            mbar_expect_tx(&fullA[sa], p.a_tx_bytes);
            tma_load_4d(sA + sa * p.a_stage_bytes, &tmA, &fullA[sa], cc * 64, tc.x0 - 1, tc.y0 - 1, tc.b);
          }
          if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
          if (!RESIDENT) {
            // Weight slices go G to a ring stage (one barrier pair per stage: the MMA issuer pays
            // ~400 cycles per wait + commit); stages run through tile and chunk boundaries, only
            // the CTA's last one may be short.
            for (int t9 = 0; t9 < 9; ++t9) {
              if (g == 0) {
                if (cc == 0 && t9 == 3) H_STAMP(0, pi, 2);
                mbar_wait(&emptyB[sb], phb ^ 1);         // freed by every CTA of the cluster
                if (cc == 0 && t9 == 3) H_STAMP(0, pi, 3);
                const int n = slices_left < p.G ? slices_left : p.G;
                if (!(p.ablate & 16)) mbar_expect_tx(&fullB[sb], n * p.b_tx_bytes);
              }
              uint8_t* dst = sW + sb * p.b_stage_bytes + g * p.b_slice_bytes;
              const int sl = t9 * p.ncc + cc;
              const int kx = p.w_slices ? 0 : sl * 64, ry = p.w_slices ? sl * p.cout_rows + tc.n0 : tc.n0;
              if (p.ablate & 16) {
              } else if (p.cluster > 1) {                // this CTA's rows of the slice, to all peers
                tma_load_2d_mcast(dst + crank * p.part_rows * 128, &tmBp, &fullB[sb], kx,
                                  ry + static_cast<int>(crank) * p.part_rows, cmask);
              } else {
                tma_load_2d(dst, &tmB, &fullB[sb], kx, ry);
              }
              --slices_left;
              if (++g == p.G || slices_left == 0) {
                if (p.ablate & 16) mbar_arrive(&fullB[sb]);
                g = 0;
                if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 || (warp == H_MMA2_WARP && p.dual)) {
    // -------------------------------------------------------------- MMA issuer(s)
    // The issuing thread is the scarce resource of the small-N layers: the tensor pipe queues only a couple
    // of MMAs, so every barrier round trip of the issuing thread (~200 cycles) and every instruction between
    // two MMAs shows up as idle tensor time (tools/gpu_mma_pattern.py: the filter-row grouped train of a
    // C = 48 tile takes 1780 cycles from one thread with the loop's waits, 1150 from two threads on
    // alternate tiles = the tensor pipe's own time).  With p.dual a second warp issues as well: dual == 1,
    // the two alternate the CTA's items (own accumulator and operand stages, nothing shared); dual == 2,
    // both walk every item and every weight stage and issue one tile each of the item's T tiles (the
    // shared stages are released by both).
    // The whole warp walks the (warp-uniform) schedule and one elected lane issues. N is 48..192
    // here, an MMA retires every 44..96 cycles and the tensor pipe queues only a couple of them
    // (tools/gpu_mma_rate.py), so the loop between two MMAs has to be a handful of instructions:
    // everything is hoisted into registers, the descriptors advance by 32-bit adds on their low
    // word, the tap order and the tiles per item are compile-time constants.
    const int iw = warp == 1 ? 0 : 1;
    const bool alt = p.dual == 1, split = p.dual == 2;
    int sa = 0, sb = 0, as = 0;
    uint32_t pha = 0, phb = 0, aphase = 0;
    const uint32_t idesc = make_idesc_f16(128, p.mma_n);
    const uint32_t idesc2 = make_idesc_f16(128, 2 * p.mma_n);      // DX: two taps side by side
    const uint64_t desc0 = make_smem_desc(0, 128, 2);      // everything but the start address
    const uint32_t dhi = static_cast<uint32_t>(desc0 >> 32), dlo = static_cast<uint32_t>(desc0);
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tap_off[tap] = static_cast<uint32_t>(((tap / 3) * p.TWp + (tap % 3)) * 128) >> 4;
    const bool issuer = elect_one();
    const bool do_mma = issuer && !(p.ablate & 4);
    const bool mc = p.cluster > 1;
    const bool stamp = p.dbg != nullptr && issuer && iw == 0;
    if (RESIDENT) { mbar_wait(wfull, 0); tc_fence_after(); }
    const uint32_t a_lo0 = dlo + ((smem_u32(sA) & 0x3FFFF) >> 4), a_step = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
    const uint32_t w_lo0 = dlo + ((smem_u32(sW) & 0x3FFFF) >> 4), w_step = static_cast<uint32_t>(p.b_slice_bytes) >> 4;
    const uint32_t w_tap = static_cast<uint32_t>(p.ncc) * w_step;
    const uint32_t a_tile1 = static_cast<uint32_t>(p.R * p.TWp * 128) >> 4;   // next tile of the item: R image rows down
    const uint32_t acc_stride = p.acc_stride;
    const int ncc = p.ncc, n_acc = p.n_acc, a_stages = p.a_stages, b_stages = p.b_stages, ksteps_last = p.ksteps_last;
    const int step = gridDim.x, total = p.total_tiles;
    uint32_t a_lo = a_lo0, b_lo = w_lo0;                    // low descriptor words of stages sa / sb
    const int G = p.G;
    const uint32_t b_stage_step = static_cast<uint32_t>(p.b_stage_bytes) >> 4;
    int g = 0, slices_left = items_cta * 9 * ncc;           // position in the weight stage; slices this CTA still consumes
    bool b_ready = false;                                   // the current weight stage was already waited for
    int mi = 0;
    bool pre = false;        // this item's tempty / first fullA were already waited for (see below)
    for (int t = blockIdx.x; t < total; t += step, ++mi) {
      if (alt && (mi & 1) != iw) {
        // the other issuer's item: step the rings past it
        for (int cc = 0; cc < ncc; ++cc) { a_lo += a_step; if (++sa == a_stages) { sa = 0; pha ^= 1; a_lo = a_lo0; } }
        if (++as == n_acc) { as = 0; aphase ^= 1; }
        continue;
      }
      if (stamp) H_STAMP(1, mi, 0);
      const bool has_next = !p.dual && t + step < total;      // look-ahead waits: single issuer only (the other issuer covers them)
      bool pre_next = false;
      if (!pre) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
      }
      if (stamp) H_STAMP(1, mi, 1);
      const uint32_t d_tmem = tmem_base + as * (T * acc_stride);
      for (int cc = 0; cc < ncc; ++cc) {
        if (!(pre && cc == 0)) {
          mbar_wait(&fullA[sa], pha);
          tc_fence_after();
        }
        if (stamp) H_STAMP(1, mi, 2);
        const int nk = (cc == ncc - 1) ? ksteps_last : 4;       // pad lanes of the last chunk are zero: skip them
        const uint32_t w_cc = w_lo0 + cc * w_step;
        if (DX) {
          // per filter row: [taps dx = 1 | dx = 0] from the view shifted by dy rows, then tap dx = 2 from
          // the view one pixel further into the first group's accumulator columns
          if (p.ablate & 64) {
            if (do_mma) {
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                const uint32_t at = a_lo + tap_off[dy * 3];
                const uint32_t bl = w_cc + dy * 3 * w_tap;
                umma_f16_lo(d_tmem, at, bl, dhi, idesc2, (cc | dy) != 0);
                if (nk > 1) umma_f16_lo(d_tmem, at + 2, bl + 2, dhi, idesc2, 1);
                if (nk > 2) umma_f16_lo(d_tmem, at + 4, bl + 4, dhi, idesc2, 1);
                if (nk > 3) umma_f16_lo(d_tmem, at + 6, bl + 6, dhi, idesc2, 1);
              }
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                const uint32_t at1 = a_lo + tap_off[dy * 3 + 1];
                const uint32_t bl2 = w_cc + dy * 3 * w_tap + 2 * w_tap;
                umma_f16_lo(d_tmem, at1, bl2, dhi, idesc, 1);
                if (nk > 1) umma_f16_lo(d_tmem, at1 + 2, bl2 + 2, dhi, idesc, 1);
                if (nk > 2) umma_f16_lo(d_tmem, at1 + 4, bl2 + 4, dhi, idesc, 1);
                if (nk > 3) umma_f16_lo(d_tmem, at1 + 6, bl2 + 6, dhi, idesc, 1);
              }
            }
          } else
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            if (do_mma) {
              const uint32_t at = a_lo + tap_off[dy * 3];
              const uint32_t at1 = a_lo + tap_off[dy * 3 + 1];
              const uint32_t bl = w_cc + dy * 3 * w_tap;
              const uint32_t bl2 = bl + 2 * w_tap;
              const uint32_t first = (cc | dy) != 0;
              umma_f16_lo(d_tmem, at, bl, dhi, idesc2, first);
              if (nk > 1) umma_f16_lo(d_tmem, at + 2, bl + 2, dhi, idesc2, 1);
              if (nk > 2) umma_f16_lo(d_tmem, at + 4, bl + 4, dhi, idesc2, 1);
              if (nk > 3) umma_f16_lo(d_tmem, at + 6, bl + 6, dhi, idesc2, 1);
              umma_f16_lo(d_tmem, at1, bl2, dhi, idesc, 1);
              if (nk > 1) umma_f16_lo(d_tmem, at1 + 2, bl2 + 2, dhi, idesc, 1);
              if (nk > 2) umma_f16_lo(d_tmem, at1 + 4, bl2 + 4, dhi, idesc, 1);
              if (nk > 3) umma_f16_lo(d_tmem, at1 + 6, bl2 + 6, dhi, idesc, 1);
            }
            if (dy == 1 && cc == ncc - 1 && has_next) {
              const int as_n = (as + 1 == n_acc) ? 0 : as + 1;
              const uint32_t aph_n = (as_n == 0) ? (aphase ^ 1) : aphase;
              mbar_wait(&tempty[as_n], aph_n ^ 1);
              const int sa_n = (sa + 1 == a_stages) ? 0 : sa + 1;
              const uint32_t pha_n = (sa_n == 0) ? (pha ^ 1) : pha;
              mbar_wait(&fullA[sa_n], pha_n);
              tc_fence_after();
              pre_next = true;
            }
          }
        } else {
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
          if (!RESIDENT && g == 0 && !b_ready) mbar_wait(&fullB[sb], phb);
          if (do_mma) {
            const uint32_t at = a_lo + tap_off[t9];
            const uint32_t bl = RESIDENT ? w_cc + t9 * w_tap : b_lo + g * w_step;
            const uint32_t first = (cc | t9) != 0;
#pragma unroll
            for (int tl = 0; tl < T; ++tl) {                 // the tiles of the item share the slice
              if (T > 1 && split && (tl & 1) != iw) continue;
              const uint32_t al = at + tl * a_tile1;
              const uint32_t dt = d_tmem + tl * acc_stride;
              umma_f16_lo(dt, al, bl, dhi, idesc, first);
              if (nk > 1) umma_f16_lo(dt, al + 2, bl + 2, dhi, idesc, 1);
              if (nk > 2) umma_f16_lo(dt, al + 4, bl + 4, dhi, idesc, 1);
              if (nk > 3) umma_f16_lo(dt, al + 6, bl + 6, dhi, idesc, 1);
            }
          }
          if (!RESIDENT) {
            --slices_left;
            if (g == G - 1 || slices_left == 0) {
              // End of a ring stage. A shared-memory access of this thread right after its own
              // tcgen05.commit stalls ~230 cycles and the tensor pipe queues only a couple of MMAs
              // (tools/gpu_mma_rate.py), so the wait for the NEXT stage goes first, while this
              // stage's MMAs are still queued, and the commit after it.
              const int sb_n = (sb + 1 == b_stages) ? 0 : sb + 1;
              const uint32_t ph_n = (sb_n == 0) ? (phb ^ 1) : phb;
              b_ready = slices_left != 0;
              if (b_ready) mbar_wait(&fullB[sb_n], ph_n);
              if (issuer) { if (mc) umma_commit_mcast(&emptyB[sb], cmask); else umma_commit(&emptyB[sb]); }
              __syncwarp();
              sb = sb_n; phb = ph_n; g = 0;
              b_lo = w_lo0 + sb * b_stage_step;
            } else {
              ++g;
            }
          }
          if (RESIDENT && t9 == 5 && cc == ncc - 1 && has_next) {
            // The MMA queue still holds this item's last taps: take the next item's barrier
            // waits (accumulator stage free, first operand tile landed) off the tensor pipe's
            // critical path.
            const int as_n = (as + 1 == n_acc) ? 0 : as + 1;
            const uint32_t aph_n = (as_n == 0) ? (aphase ^ 1) : aphase;
            mbar_wait(&tempty[as_n], aph_n ^ 1);
            const int sa_n = (sa + 1 == a_stages) ? 0 : sa + 1;
            const uint32_t pha_n = (sa_n == 0) ? (pha ^ 1) : pha;
            mbar_wait(&fullA[sa_n], pha_n);
            tc_fence_after();
            pre_next = true;
          }
        }
        }
        if (issuer) umma_commit(&emptyA[sa]);
        __syncwarp();
        a_lo += a_step;
        if (++sa == a_stages) { sa = 0; pha ^= 1; a_lo = a_lo0; }
      }
      if (issuer) umma_commit(&tfull[as]);
      if (stamp) H_STAMP(1, mi, 3);
      __syncwarp();
      if (++as == n_acc) { as = 0; aphase ^= 1; }
      pre = pre_next;
    }
  } else if (warp < H_MMA2_WARP) {
    // ---------------------------------------------------------------- epilogue
    // Two groups of four warps ping-pong over the tiles: group g owns TMEM accumulator stage g
    // and staging buffer g, so the latency chain of one tile's epilogue (TMEM load, residual
    // read, staging, store issue) overlaps the other group's.
    griddep_wait();                             // residual reads and output writes follow the previous kernel
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int groups_total = p.N_tile >> 5;                 // 32-column groups
    const int m = quarter * 32 + lane;
    const int r = m / p.TWp, xx = m - r * p.TWp;
    const int ms = r * p.TW + xx;                            // row of this pixel in the dense R x TW staging tile
    const bool leader = (quarter == 2 && lane == 0);         // first warp of the group
    const uint32_t bias_u = smem_u32(s_bias);
    uint8_t* sStage = sOut + grp * p.nblk * H_STAGE_BLOCK;
    const bool prefetch_res = !DX && (p.res != nullptr) && groups_total <= 2 && T <= 2;
    // residual rows are fetched one of this group's tiles ahead (the accumulator ring lets the
    // MMAs run far ahead, so nothing else would hide the read latency)
    uint4 rnext[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) rnext[q] = make_uint4(0, 0, 0, 0);
    // single mode: the groups alternate over the items; multi-tile items: group g takes tiles g, g+2 of every item
    const bool staged = RESIDENT && H_STAGED_BUILD && p.out_bufs > 0;   // staging tile + TMA stores, or direct stores from registers
    constexpr bool multi = T > 1;
    constexpr int subs = multi ? T / 2 : 1;          // tiles of an item handled by this group
    const int t_start = blockIdx.x + (multi ? 0 : grp * gridDim.x);
    const int t_step = (multi ? 1 : 2) * gridDim.x;
    const int as_step = multi ? 1 : 2;
    const uint32_t item_cols = T * p.acc_stride;
    HTileIter ti, tn;                       // this group's current item and the one after it
    ti.init(p, t_start, t_step);
    tn = ti;
    auto residual_row = [&](int t, const HTileIter& itr) -> const __half* {
      if (t >= p.total_tiles) return nullptr;
      const HTile c = itr.get(p);
      const int yy = c.y0 + (multi ? grp * p.R : 0) + r, xq = c.x0 + xx;
      if (!((xx < p.TW) && (yy < p.H) && (xq < p.W))) return nullptr;
      return p.res + ((static_cast<size_t>(c.b) * p.H + yy) * p.W + xq) * p.Cout_pad + c.n0;
    };
    if (prefetch_res) {
      const __half* r0 = residual_row(t_start, tn);
      if (r0) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q < groups_total * 4) rnext[q] = __ldg(reinterpret_cast<const uint4*>(r0) + q);
      }
    }
    int as = multi ? 0 : grp;                // single mode: n_acc is even, this group sees stages grp, grp+2, ...
    uint32_t aphase = 0;
    int ei = 0;
    for (int t = t_start; t < p.total_tiles; t += t_step, ++ei) {
      if (leader) H_STAMP(2 + grp, ei, 0);
      const HTile tc = ti.get(p);
      ti.advance(p);
      tn.advance(p);
      for (int sub = 0; sub < subs; ++sub) {
      const int tl = multi ? grp + 2 * sub : 0;              // tile of the item this pass drains
      const uint32_t col_off = tl * p.acc_stride;
      const int ty0 = tc.y0 + tl * p.R;
      const int y = ty0 + r, x = tc.x0 + xx;
      const bool valid = (xx < p.TW) && (y < p.H) && (x < p.W);
      const size_t pix = (static_cast<size_t>(tc.b) * p.H + y) * p.W + x;
      const __half* rrow = (p.res && valid) ? p.res + pix * p.Cout_pad + tc.n0 : nullptr;
      uint4 rdx[8];                                          // DX: this pixel's residual channels, in flight while the accumulator completes
      if (DX) {
#pragma unroll
        for (int q = 0; q < 8; ++q) rdx[q] = make_uint4(0, 0, 0, 0);
        if (rrow) {
#pragma unroll
          for (int q = 0; q < 2 * DXB; ++q) rdx[q] = ldg_nc_v4(rrow + q * 8);
        }
        if (p.res) {
          // this group's next tile: its residual rows start their way from HBM to L2 now (one 128-byte line
          // per pixel), so the loads above find them there a tile from now
          const __half* rn = residual_row(t + t_step, tn);
          if (rn) asm volatile("prefetch.global.L2 [%0];" ::"l"(rn));
        }
      }
      uint4 rpre[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) rpre[q] = rnext[q];
      if (prefetch_res) {
#pragma unroll
        for (int q = 0; q < 8; ++q) rnext[q] = make_uint4(0, 0, 0, 0);
        const __half* rn = residual_row(t + t_step, tn);
        if (rn) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < groups_total * 4) rnext[q] = __ldg(reinterpret_cast<const uint4*>(rn) + q);
        }
      }
      // this group's previous TMA stores must have finished reading its staging buffer
      if (leader) H_STAMP(2 + grp, ei, 1);
      if (staged) {
        if (leader) bulk_wait_read();
        if (leader) H_STAMP(2 + grp, ei, 2);
        named_bar_sync(1 + grp, 128);
      }
      if (leader) H_STAMP(2 + grp, ei, 3);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if (leader) H_STAMP(2 + grp, ei, 4);
      const uint32_t taddr = tmem_base + as * item_cols + col_off + (static_cast<uint32_t>(quarter * 32) << 16);
      // bias / residual / ReLU on one 32-column group and its 4 swizzled 16-byte staging writes
      __half* yrow = p.yptr + pix * p.Cout_pad + tc.n0;
      auto finish_group = [&](const uint32_t (&acc)[32], const uint4 (&rq)[4], int g) {
        if ((p.ablate & 2) || xx >= p.TW) return;            // halo columns are not staged / stored
        const float* bb = s_bias + tc.n0 + g * 32;
        if (!staged) {
          // Streamed-weight layers: no staging tile - its shared memory buys a deeper weight ring,
          // which is what bounds them - each lane writes its pixel's 32 channels as two full sectors.
          if (!valid) return;
          uint32_t o[16];
          const uint32_t bb_u = bias_u + (tc.n0 + g * 32) * 4;
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
              o[q * 4 + j] = h_pack_half2(a, b);
            }
          }
          stg_v8(yrow + g * 32, o);
          stg_v8(yrow + g * 32 + 16, o + 8);
          return;
        }
        uint8_t* blk = sStage + (g >> 1) * H_STAGE_BLOCK + ms * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t rr[4] = {rq[q].x, rq[q].y, rq[q].z, rq[q].w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = q * 8 + 2 * j;
            const __half2 rh = *reinterpret_cast<const __half2*>(&rr[j]);
            float a = (g * 32 + c < p.mma_n) ? __uint_as_float(acc[c]) : 0.0f;
            float b = (g * 32 + c + 1 < p.mma_n) ? __uint_as_float(acc[c + 1]) : 0.0f;
            a += bb[c] + __low2float(rh);
            b += bb[c + 1] + __high2float(rh);
            if (p.relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
            o[j] = h_pack_half2(a, b);
          }
          const int chunk = ((g & 1) * 4 + q) ^ (ms & 7);      // SWIZZLE_128B: 16-byte chunk index
          *reinterpret_cast<uint4*>(blk + (chunk << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      };
      if constexpr (DX) {
        // out[p] = G0[p] + G1'[p + 1]: the two column groups of the accumulator, the shift by a warp
        // shuffle (a patch row of TWp <= 32 pixels lies inside this warp's lanes; the lanes whose
        // neighbour belongs to the next row are halo columns, which are not stored).  The epilogue is a
        // latency chain (TMEM load -> shuffle -> math -> store) that two groups of warps have to get
        // through once per tile time: all the tile's TMEM loads go out together, the accumulator stage is
        // handed back as soon as they have landed, and the arithmetic runs on registers only.
        constexpr int NBP = DXB == 1 ? 1 : 2;                  // 16-channel blocks per pass (64 accumulator registers; three blocks at once spilled and measured slower)
#pragma unroll
        for (int b0 = 0; b0 < DXB; b0 += NBP) {
          uint32_t g1[16 * NBP], g0[16 * NBP];
          if (p.ablate & 32) {
#pragma unroll
            for (int j = 0; j < 16 * NBP; ++j) { g1[j] = 0u; g0[j] = 0u; }
          } else {
#pragma unroll
            for (int b = 0; b < NBP; ++b) {
              if (b0 + b < DXB) {
                tmem_ld16(taddr + (b0 + b) * 16, g1 + 16 * b);
                tmem_ld16(taddr + 16 * DXB + (b0 + b) * 16, g0 + 16 * b);
              }
            }
            tmem_ld_wait();
          }
          if (b0 + NBP >= DXB) {                                // accumulator drained
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
          }
#pragma unroll
          for (int b = 0; b < NBP; ++b) {
            const int cb = b0 + b;
            if (cb >= DXB) break;
            uint32_t o[8];
            const uint4 r0 = rdx[2 * cb], r1 = rdx[2 * cb + 1];
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bq = lds_v4f(bias_u + (tc.n0 + cb * 16 + j4 * 4) * 4);
              const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const int j = j4 * 2 + h2;
                float a = __uint_as_float(g0[16 * b + 2 * j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(g1[16 * b + 2 * j]), 1);
                float bvv = __uint_as_float(g0[16 * b + 2 * j + 1]) + __shfl_down_sync(0xffffffffu, __uint_as_float(g1[16 * b + 2 * j + 1]), 1);
                const __half2 rh = *reinterpret_cast<const __half2*>(&rr[j]);
                a += bv[2 * h2] + __low2float(rh);
                bvv += bv[2 * h2 + 1] + __high2float(rh);
                if (p.relu) { a = fmaxf(a, 0.0f); bvv = fmaxf(bvv, 0.0f); }
                o[j] = h_pack_half2(a, bvv);
              }
            }
            if (valid && !(p.ablate & 2)) stg_v8(yrow + cb * 16, o);
          }
        }
        if (valid && !(p.ablate & 2)) {
          // pad channels stay zero
          const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
          for (int cb = DXB; cb < 4; ++cb) stg_v8(yrow + cb * 16, z);
        }
      } else {
        for (int g = 0; g < groups_total; ++g) {
          uint4 rq[4];
          if (prefetch_res) {
#pragma unroll
            for (int q = 0; q < 4; ++q) rq[q] = (g == 0) ? rpre[q] : rpre[4 + q];
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) rq[q] = make_uint4(0, 0, 0, 0);
            if (rrow) {
#pragma unroll
              for (int q = 0; q < 4; ++q) rq[q] = __ldg(reinterpret_cast<const uint4*>(rrow + g * 32) + q);
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
          finish_group(acc, rq, g);
        }
      }
      if (leader) H_STAMP(2 + grp, ei, 5);
      if (!DX) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);            // accumulator drained
      }
      if (staged) {
        fence_proxy_async();                              // staging writes -> visible to the TMA unit
        if (leader) H_STAMP(2 + grp, ei, 6);
        named_bar_sync(1 + grp, 128);
      }
      if (staged && leader && !(p.ablate & 1)) {
        // one store per 64-channel block: the R x TW box is dense in the staging tile; rows below
        // the image and columns right of it are clipped by the TMA unit
        if (ty0 < p.H)
          for (int kb = 0; kb < p.nblk; ++kb)
            tma_store_4d(&tmY, sStage + kb * H_STAGE_BLOCK, tc.n0 + kb * 64, tc.x0, ty0, tc.b);
        bulk_commit();
      }
      if (leader) H_STAMP(2 + grp, ei, 7);
      }
      as += as_step;
      if (as >= p.n_acc) { as -= p.n_acc; aphase ^= 1; }
    }
    if (leader) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();      // no CTA leaves while a peer may still write into it
  if (p.dbg && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.dbg[4 * 64 * 8 + 2 * blockIdx.x + 1] = static_cast<long long>(gt);
  }
  if (warp == 1) tmem_dealloc(tmem_base, H_TMEM_COLS);
}

}  // namespace

// Chooses the tile geometry and launches; returns CAL_E_UNSUPPORTED (without setting an error
// the caller must report) when the shape is better served by the generic kernel.
int launch_conv3x3_halo(const CalConvArgs* a, void* stream) {
  if (a->ksize != 3 || a->stride != 1 || a->mode != 0) return CAL_E_UNSUPPORTED;
  HaloParams p{};
  p.B = a->B; p.H = a->Hout; p.W = a->Wout; p.Cout_pad = a->Cout_pad;
  // tile geometry: few tiles (MMA rows are spent on every tile, used or not) and a small halo
  // patch (L2 -> SM bytes per tile)
  // filter-row grouping (template DX) applies to the single-chunk, single-N-tile layers (C <= 64 in and out);
  // it needs a patch row inside one warp (TWp <= 32)
  static const bool dx_enabled = [] { const char* e = getenv("CAL_CONV_DX"); return !(e && e[0] == '0'); }();
  const bool dx_shape = dx_enabled && a->Cin_pad == 64 && a->Cout_pad == 64 && (a->Cout_rows == 32 || a->Cout_rows == 48 || a->Cout_rows == 64);
  long best = -1;
  for (int twp = 16; twp <= (dx_shape ? 32 : 128); twp *= 2) {
    const int tw = twp - 2, r = 128 / twp;
    const long tiles = static_cast<long>((a->Wout + tw - 1) / tw) * ((a->Hout + r - 1) / r);
    const long cost = tiles * (128 * 4 + (r + 2) * twp);
    if (best < 0 || cost < best) { best = cost; p.TWp = twp; p.TW = tw; p.R = r; }
  }
  p.tiles_x = (a->Wout + p.TW - 1) / p.TW;
  int n_tiles = 1;
  while (a->Cout_pad % (64 * n_tiles) != 0 || a->Cout_pad / n_tiles > 256) ++n_tiles;
  auto set_ntiles = [&](int nt) {
    n_tiles = nt;
    p.n_tiles = nt;
    p.N_tile = a->Cout_pad / nt;
    p.nblk = p.N_tile / 64;
    p.mma_n = p.N_tile < a->Cout_rows ? p.N_tile : a->Cout_rows;   // weight rows of the (only) ragged N tile
    p.acc_stride = (p.N_tile + 31) & ~31;
    p.b_slice_bytes = ((p.mma_n * 128) + 1023) & ~1023;
    p.b_tx_bytes = static_cast<uint32_t>(p.mma_n * 128);
  };
  set_ntiles(n_tiles);
  if (n_tiles > 1 && a->Cout_rows != a->Cout_pad) return CAL_E_UNSUPPORTED;   // ragged last N tile: generic kernel
  p.ncc = a->Cin_pad / 64;
  {
    const int cin = (a->Cin > 0 && a->Cin <= a->Cin_pad) ? a->Cin : a->Cin_pad;
    p.ksteps_last = (cin - (p.ncc - 1) * 64 + 15) / 16;
    if (p.ksteps_last < 1) p.ksteps_last = 1;
    if (p.ksteps_last > 4) p.ksteps_last = 4;
  }
  p.relu = a->relu;
  p.w_slices = a->w_slices ? 1 : 0;
  p.cout_rows = a->Cout_rows;
  { const char* e = getenv("CAL_DEBUG_ABLATE"); p.ablate = e ? atoi(e) : 0; }
  {
    // CAL_DEBUG_TIMELINE=<device pointer, hex>: 4 roles x 64 tiles x 8 stamps of long long
    const char* e = getenv("CAL_DEBUG_TIMELINE");
    p.dbg = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 16)) : nullptr;
  }
  p.bias = a->bias;
  p.res = reinterpret_cast<const __half*>(a->res);
  p.yptr = reinterpret_cast<__half*>(a->y);
  // Work items of T vertically adjacent tiles share every weight slice (1/T of the weight traffic and,
  // for streamed weights, of the per-slice barrier wait + commit of the MMA issuer). T = 2 whenever two
  // items of two accumulators each fit the 512 TMEM columns, so the accumulator ring stays double
  // buffered; measured on B200 (tools/gpu_ablate.py, batch 64): C=96 68x120 199 -> 133 us, while giving
  // up the double buffering for T = 2 / 4 (C=192, 384; C=96 at T = 4) loses what the sharing gains.
  // Resident weights share nothing but the handshakes, and with a residual input the two epilogue
  // groups of a T = 2 item read their residual rows at the same moment: T = 1 there (178 vs 193 us).
  const int tail = (2 * H_MAX_A_STAGES + 2 * H_MAX_B_STAGES + 1 + 2 * H_MAX_ACC) * 8 + 16 + 16 + H_MAX_BIAS * 4;
  const int w_all = 9 * p.ncc * p.b_slice_bytes;
  const int avail = 224 * 1024 - smem_headroom();
  const int budget_res = avail - 1024 - tail - (H_STAGED_BUILD ? 2 * p.nblk * H_STAGE_BLOCK : 0);   // (evaluated before any N split)
  const int budget_ring = avail - 1024 - tail;
  auto set_tiles = [&](int T) {
    p.T = T;
    p.RI = p.R * T;
    p.tiles_y = (a->Hout + p.RI - 1) / p.RI;
    p.total_tiles = a->B * p.tiles_x * p.tiles_y * n_tiles;
    p.a_stage_bytes = (p.RI + 2) * p.TWp * 128 + 1024;   // + pad rows read by the last taps of halo columns
    p.a_tx_bytes = static_cast<uint32_t>((p.RI + 2) * p.TWp * 128);
    p.n_acc = H_TMEM_COLS / (p.acc_stride * T);
    if (p.n_acc > H_MAX_ACC) p.n_acc = H_MAX_ACC;
    if (T == 1) p.n_acc &= ~1;                          // single mode: stage parity == epilogue group
  };
  static const int force_tiles = [] { const char* e = getenv("CAL_CONV_TILES"); return e ? atoi(e) : 0; }();   // experiments
  const int max_tiles = force_tiles ? force_tiles : 2;
  const bool fits2 = 4 * p.acc_stride <= H_TMEM_COLS && a->Hout > p.R;
  set_tiles((max_tiles >= 2 && fits2 && (force_tiles || !a->res)) ? 2 : 1);
  if (n_tiles == 1 && w_all + 2 * p.a_stage_bytes <= budget_res) {
    p.w_resident = 1;
    if (dx_shape && p.b_slice_bytes == p.mma_n * 128) {
      // (slices of mma_n * 128 bytes are contiguous: two neighbouring taps form one 2 * mma_n-row operand)
      p.dx = 1;
      p.acc_stride = (2 * p.mma_n + 31) & ~31;
      set_tiles(1);
    }
    // Outputs go straight from registers to global memory (two full 32-byte sectors per lane and
    // 32-channel group). The alternative - a swizzled staging tile drained by TMA stores,
    // CAL_CONV_DIRECT=0 in a -DCAL_HALO_STAGED build - costs 32 KB of shared-memory traffic per tile next to the MMA operand
    // reads, which are what bounds these small-N layers (an N = 48 MMA takes 32 + N/4 cycles of
    // operand reads against N/2 of math): measured 217 -> 193 us with a residual, equal without.
    static const bool direct = [] { const char* e = getenv("CAL_CONV_DIRECT"); return !H_STAGED_BUILD || !(e && e[0] == '0'); }();
    p.out_bufs = direct ? 0 : 2;
    p.b_stages = 0;
    p.G = 1;
    p.b_stage_bytes = p.b_slice_bytes;
    p.a_stages = (budget_res - w_all) / p.a_stage_bytes;
    // more than four patches in flight measured slower with a residual input (175 -> 189 us at C = 48)
    { static const int cap = [] { const char* e = getenv("CAL_A_STAGES_RES"); return e ? atoi(e) : 4; }();
      if (p.a_stages > cap) p.a_stages = cap; }
  } else {
    p.w_resident = 0;
    p.out_bufs = 0;
    // Experiment (CAL_CONV_NTILE_MAX=128): N tiles of at most 128 columns, so that T = 2 keeps a
    // double-buffered accumulator ring (C = 192 -> 2 x 96, C = 384 -> 3 x 128) and the weight slices
    // are read half as often per output. Measured on B200: C=192 34x60 97 us vs 86 us unsplit (the
    // N = 96 MMAs run at 86 % of the N = 192 rate and the input patch is read twice), C=384 equal -
    // so the default keeps the widest N tile.
    static const int ntile_max = [] { const char* e = getenv("CAL_CONV_NTILE_MAX"); return e ? atoi(e) : 256; }();
    if (a->Cout_rows == a->Cout_pad && p.N_tile > ntile_max) {
      for (int nt = n_tiles; nt <= a->Cout_pad / 32; ++nt)
        if (a->Cout_pad % nt == 0 && (a->Cout_pad / nt) % 32 == 0 && a->Cout_pad / nt <= ntile_max) { set_ntiles(nt); break; }
    }
    const bool fits2s = 4 * p.acc_stride <= H_TMEM_COLS && a->Hout > p.R;
    for (int T = 4; T >= 1; T >>= 1) {
      if (T > max_tiles && T > 1) continue;
      if (T > 1 && (T * p.acc_stride > H_TMEM_COLS || a->Hout <= p.R * (T / 2))) continue;
      if (T > 1 && !force_tiles && !fits2s) continue;
      set_tiles(T);
      if (T == 1 || (budget_ring - 2 * p.a_stage_bytes) / p.b_slice_bytes >= 4) break;
    }
    // G slices per ring stage: as many as leave three stages next to two input patches
    static const int force_g = [] { const char* e = getenv("CAL_CONV_G"); return e ? atoi(e) : 0; }();
    p.a_stages = 2;
    p.G = 1;
    for (int G = 3; G >= 1; --G) {
      if (force_g && G != force_g) continue;
      const int st = (budget_ring - p.a_stages * p.a_stage_bytes) / (G * p.b_slice_bytes);
      if (st >= 3 || G == 1 || force_g) { p.G = G; break; }
    }
    p.b_stage_bytes = p.G * p.b_slice_bytes;
    // a third / fourth input patch while the ring keeps four stages
    static const int want_a = [] { const char* e = getenv("CAL_A_STAGES"); return e ? atoi(e) : 4; }();
    while (p.a_stages < want_a && p.a_stages < H_MAX_A_STAGES &&
           (budget_ring - (p.a_stages + 1) * p.a_stage_bytes) / p.b_stage_bytes >= 4)
      ++p.a_stages;
    p.b_stages = (budget_ring - p.a_stages * p.a_stage_bytes) / p.b_stage_bytes;
    if (p.b_stages > H_MAX_B_STAGES) p.b_stages = H_MAX_B_STAGES;
    if (p.b_stages < 2) return CAL_E_UNSUPPORTED;
  }
  if (p.a_stages > H_MAX_A_STAGES) p.a_stages = H_MAX_A_STAGES;
  if (p.a_stages < 2) return CAL_E_UNSUPPORTED;
  // streamed weights: multicast every slice over a cluster of 4 (or 2) CTAs - each loads a
  // quarter (half) of the rows - so the L2 -> SM weight traffic per tile drops by the same factor
  p.cluster = 1; p.n_slowest = 0; p.part_rows = p.mma_n;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    CAL_CHECK_CUDA(cudaGetDevice(&dev));
    CAL_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<true, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<true, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<true, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CAL_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  if (!p.w_resident) {
    static const int max_cluster = [] { const char* e = getenv("CAL_CONV_CLUSTER"); return e ? atoi(e) : 1; }();   // measured slower than unicast at 4 (lock-step stage recycling): opt-in
    const int m_items = a->B * p.tiles_x * p.tiles_y;
    for (int cl = 4; cl >= 2; cl >>= 1) {
      if (cl > max_cluster) continue;
      if (grid % cl || p.total_tiles % cl || (p.mma_n / cl) % 8 || p.mma_n % cl) continue;
      if (n_tiles > 1 && m_items % cl) continue;
      p.cluster = cl; p.part_rows = p.mma_n / cl; p.n_slowest = n_tiles > 1 ? 1 : 0;
      break;
    }
  }
  // two MMA-issuing warps (see the kernel): alternate items for the resident T = 1 variants, one tile each of
  // the T = 2 items; the streamed T = 1 layers (N = 192: the MMAs are long enough for one thread) keep one
  static const int dual_mask = [] { const char* e = getenv("CAL_CONV_DUAL"); return e ? atoi(e) : 3; }();
  p.dual = 0;
  if (p.cluster == 1) {
    // (alternating items: the second issuer's waits run ncc stages ahead of the first's; a parity wait is only
    // meaningful within one pass of the ring)
    if (p.T == 1 && p.w_resident && 2 * p.ncc <= p.a_stages && p.n_acc >= 2 && (dual_mask & 1)) p.dual = 1;
    if (p.T == 2 && (dual_mask & 2)) p.dual = 2;
  }
  const int w_slots = p.w_resident ? 9 * p.ncc : p.b_stages * p.G;
  const size_t smem = 1024 + static_cast<size_t>(p.a_stages) * p.a_stage_bytes +
                      static_cast<size_t>(w_slots) * p.b_slice_bytes + static_cast<size_t>(p.out_bufs) * p.nblk * H_STAGE_BLOCK + tail;

  CUtensorMap tmA, tmB, tmBp, tmY;
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin_pad, (uint64_t)a->Win, (uint64_t)a->Hin, (uint64_t)a->B};
    const uint64_t strides[3] = {(uint64_t)a->Cin_pad * 2, (uint64_t)a->Win * a->Cin_pad * 2,
                                 (uint64_t)a->Hin * a->Win * a->Cin_pad * 2};
    const uint32_t box[4] = {64, (uint32_t)p.TWp, (uint32_t)(p.RI + 2), 1};
    int rc = encode_tmap_f16(&tmA, a->x, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  {
    const uint64_t ktot = 9ull * a->Cin_pad;
    // (rows, K) K-major matrix, or slice-major: (slices * rows, 64)
    const uint64_t dims[2] = {a->w_slices ? 64ull : ktot,
                              a->w_slices ? (ktot / 64) * (uint64_t)a->Cout_rows : (uint64_t)a->Cout_rows};
    const uint64_t strides[1] = {a->w_slices ? 128ull : ktot * 2};
    const uint32_t box[2] = {64, (uint32_t)p.mma_n};
    int rc = encode_tmap_f16(&tmB, a->w, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
    const uint32_t boxp[2] = {64, (uint32_t)p.part_rows};
    rc = encode_tmap_f16(&tmBp, a->w, 2, dims, strides, boxp, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)a->Cout_pad, (uint64_t)a->Wout, (uint64_t)a->Hout, (uint64_t)a->B};
    const uint64_t strides[3] = {(uint64_t)a->Cout_pad * 2, (uint64_t)a->Wout * a->Cout_pad * 2,
                                 (uint64_t)a->Hout * a->Wout * a->Cout_pad * 2};
    const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.R, 1};
    int rc = encode_tmap_f16(&tmY, a->y, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != CAL_OK) return rc;
  }
  {
    static const bool show = getenv("CAL_DEBUG_CONFIG") != nullptr;
    if (show)
      fprintf(stderr, "halo conv %dx%d Cin_pad %d Cout_pad %d: N_tile %d G %d TWp %d R %d T %d resident %d a_stages %d b_stages %d "
                      "slice %d B n_acc %d cluster %d smem %zu grid %d items %d dx %d dual %d\n",
              a->Hout, a->Wout, a->Cin_pad, a->Cout_pad, p.N_tile, p.G, p.TWp, p.R, p.T, p.w_resident, p.a_stages, p.b_stages,
              p.b_slice_bytes, p.n_acc, p.cluster, smem, grid, p.total_tiles, p.dx, p.dual);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(H_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (p.w_resident && p.dx && p.mma_n == 32) CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<true, 1, 2>, tmA, tmB, tmBp, tmY, p));
  else if (p.w_resident && p.dx && p.mma_n == 48) CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<true, 1, 3>, tmA, tmB, tmBp, tmY, p));
  else if (p.w_resident && p.dx) CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<true, 1, 4>, tmA, tmB, tmBp, tmY, p));
  else if (p.w_resident && p.T == 2) CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<true, 2>, tmA, tmB, tmBp, tmY, p));
  else if (p.w_resident) CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<true, 1>, tmA, tmB, tmBp, tmY, p));
  else if (p.T == 4) CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<false, 4>, tmA, tmB, tmBp, tmY, p));
  else if (p.T == 2) CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<false, 2>, tmA, tmB, tmBp, tmY, p));
  else CAL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<false, 1>, tmA, tmB, tmBp, tmY, p));
  return CAL_OK;
}

}  // namespace cal
