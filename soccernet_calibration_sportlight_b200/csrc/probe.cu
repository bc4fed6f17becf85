// probe.cu - hardware experiments behind the kernel design (not on the product path).
//
// cal_debug_mma_pattern: cycles per "tile" for the tcgen05.mma trains the 3x3 kernels issue, measured in
// isolation (one warp issuing, operands resident in shared memory, optional epilogue-like TMEM readers):
// which part of a tile's issue time is the instruction mix itself.
#define CAL_TU "probe.cu"
#include "common.cuh"

namespace cal {
namespace {

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

// pattern (n = output channels, TWp = 32 pixels per patch row, nk K = 16 steps per 64-channel chunk):
//   0  filter-row grouping as conv3x3_halo_kernel<.,.,DX>: per dy  nk x N = 2n at row dy*32, nk x N = n at row dy*32 + 1
//   1  the same with every A operand at the same address
//   2  nine taps, N = n each (9 * nk MMAs), shifted views
//   3  per dy nk x N = 3n (three taps side by side)
//   4  pattern 0 with N = 2n for every MMA
//   5  one N = 256 MMA per (dy, k) (reference point: large N)
//   6  pattern 0, the N = 2n trains of all dy first, then the N = n trains
__global__ void __launch_bounds__(224, 1)
mma_pattern_probe_kernel(int pattern, int n, int nk, int iters, int flags, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                        // 2 stages x 256 rows x 128 B
  uint8_t* sB = smem + 2 * 256 * 128;        // up to 1152 rows x 128 B
  __shared__ uint64_t mbar, mbar2, mbar3;
  __shared__ uint32_t tslot[2];
  __shared__ volatile uint32_t stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total16 = (2 * 256 + 1152) * 128 / 16;
  uint32_t seed = 12345u + threadIdx.x * 977u + blockIdx.x * 131071u;
  for (int i = threadIdx.x; i < total16; i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (flags & 1) {
      // small fp16 values: sign + exponent 0x2c..0x2f (0.06 .. 0.5), random mantissa
      auto h = [&]() { const uint32_t r = lcg(seed); return (r & 0x83FFu) | 0x2C00u | ((r >> 16) & 0x0300u); };
      v.x = h() | (h() << 16); v.y = h() | (h() << 16); v.z = h() | (h() << 16); v.w = h() | (h() << 16);
    }
    reinterpret_cast<uint4*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_init(&mbar2, 1u << 20); mbar_init(&mbar3, 1); fence_barrier_init(); stop = 0u; }
  if (warp == 0) tmem_alloc(&tslot[0], 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tslot[0];
  const int n_acc = ((flags >> 8) & 0xF) ? ((flags >> 8) & 0xF) : 1;
  const uint32_t acc_cols = 512 / n_acc;
  const bool dual = (flags & 64) != 0;
  if (warp == 1 || (dual && warp == 6)) {
    const int iw = warp == 1 ? 0 : 1;            // issuer index: dual issuers take alternate tiles, own accumulators / barriers
    uint64_t* done = iw ? &mbar3 : &mbar;
    const bool issuer = elect_one();
    const uint32_t id1 = make_idesc_f16(128, n), id2 = make_idesc_f16(128, 2 * n), id3 = make_idesc_f16(128, 3 * n <= 256 ? 3 * n : 256);
    const uint32_t id256 = make_idesc_f16(128, 256);
    const uint64_t desc0 = make_smem_desc(0, 128, 2);
    const uint32_t dhi = static_cast<uint32_t>(desc0 >> 32), dlo = static_cast<uint32_t>(desc0);
    const uint32_t a0 = dlo + ((smem_u32(sA) & 0x3FFFF) >> 4), b0 = dlo + ((smem_u32(sB) & 0x3FFFF) >> 4);
    const uint32_t a_stage = (256 * 128) >> 4;
    const uint32_t w_tap = static_cast<uint32_t>(n * 128) >> 4;
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
      tap_off[tap] = (pattern == 1) ? 0u : static_cast<uint32_t>(((tap / 3) * 32 + (tap % 3)) * 128) >> 4;
    long long t0 = 0, t1 = 0;
    if (issuer) {
      for (int i = 0; i < 16; ++i) umma_f16_lo(tmem_base + iw * 256, a0 + 2 * (i & 3), b0 + 2 * (i & 3), dhi, id1, i != 0);
      umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    if (issuer) {
      t0 = clock64();
      int as = 0, sa = iw;
      const int my_iters = dual ? iters / 2 : iters;
      for (int it = 0; it < my_iters; ++it) {
        const uint32_t d = tmem_base + (dual ? iw * 256 + as * (256 / n_acc) : as * acc_cols);
        const uint32_t a_lo = a0 + sa * a_stage;
        if (pattern == 0 || pattern == 1 || pattern == 4) {
          const uint32_t idb = pattern == 4 ? id2 : id1;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t at = a_lo + tap_off[dy * 3], at1 = a_lo + tap_off[dy * 3 + 1];
            const uint32_t bl = b0 + dy * 3 * w_tap, bl2 = bl + 2 * w_tap;
            umma_f16_lo(d, at, bl, dhi, id2, dy != 0);
            if (nk > 1) umma_f16_lo(d, at + 2, bl + 2, dhi, id2, 1);
            if (nk > 2) umma_f16_lo(d, at + 4, bl + 4, dhi, id2, 1);
            if (nk > 3) umma_f16_lo(d, at + 6, bl + 6, dhi, id2, 1);
            umma_f16_lo(d, at1, bl2, dhi, idb, 1);
            if (nk > 1) umma_f16_lo(d, at1 + 2, bl2 + 2, dhi, idb, 1);
            if (nk > 2) umma_f16_lo(d, at1 + 4, bl2 + 4, dhi, idb, 1);
            if (nk > 3) umma_f16_lo(d, at1 + 6, bl2 + 6, dhi, idb, 1);
            if (dy == 1 && (flags & 8)) {          // the real loop's look-ahead waits (already completed phases)
              mbar_wait(done, 0);
              mbar_wait(done, 0);
              if (!(flags & 16)) tc_fence_after();
            }
          }
        } else if (pattern == 6) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t at = a_lo + tap_off[dy * 3];
            const uint32_t bl = b0 + dy * 3 * w_tap;
            umma_f16_lo(d, at, bl, dhi, id2, dy != 0);
            if (nk > 1) umma_f16_lo(d, at + 2, bl + 2, dhi, id2, 1);
            if (nk > 2) umma_f16_lo(d, at + 4, bl + 4, dhi, id2, 1);
            if (nk > 3) umma_f16_lo(d, at + 6, bl + 6, dhi, id2, 1);
          }
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t at1 = a_lo + tap_off[dy * 3 + 1];
            const uint32_t bl2 = b0 + dy * 3 * w_tap + 2 * w_tap;
            umma_f16_lo(d, at1, bl2, dhi, id1, 1);
            if (nk > 1) umma_f16_lo(d, at1 + 2, bl2 + 2, dhi, id1, 1);
            if (nk > 2) umma_f16_lo(d, at1 + 4, bl2 + 4, dhi, id1, 1);
            if (nk > 3) umma_f16_lo(d, at1 + 6, bl2 + 6, dhi, id1, 1);
          }
        } else if (pattern == 2) {
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) {
            const uint32_t at = a_lo + tap_off[t9], bl = b0 + t9 * w_tap;
            umma_f16_lo(d, at, bl, dhi, id1, t9 != 0);
            if (nk > 1) umma_f16_lo(d, at + 2, bl + 2, dhi, id1, 1);
            if (nk > 2) umma_f16_lo(d, at + 4, bl + 4, dhi, id1, 1);
            if (nk > 3) umma_f16_lo(d, at + 6, bl + 6, dhi, id1, 1);
          }
        } else if (pattern == 3 || pattern == 5) {
          const uint32_t idx = pattern == 3 ? id3 : id256;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t at = a_lo + tap_off[dy * 3], bl = b0 + dy * 3 * w_tap;
            umma_f16_lo(d, at, bl, dhi, idx, dy != 0);
            if (nk > 1) umma_f16_lo(d, at + 2, bl + 2, dhi, idx, 1);
            if (nk > 2) umma_f16_lo(d, at + 4, bl + 4, dhi, idx, 1);
            if (nk > 3) umma_f16_lo(d, at + 6, bl + 6, dhi, idx, 1);
          }
        }
        if (flags & 2) { umma_commit(&mbar2); if (flags & 32) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&tslot[1])) : "memory"); } umma_commit(&mbar2); }
        if (++as == n_acc) as = 0;
        sa = (sa + 1) & 1;
      }
      umma_commit(done);
    }
    mbar_wait(done, 1);
    if (issuer) {
      t1 = clock64();
      if (dual) atomicMax(reinterpret_cast<unsigned long long*>(out + blockIdx.x), static_cast<unsigned long long>(t1 - t0));
      else out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
    if (lane == 0) atomicAdd(const_cast<uint32_t*>(&stop), 1u);
  } else if (warp >= 2 && warp < 6 && (flags & 4)) {
    // epilogue-like TMEM readers: warps 2..5 read 32 columns at a time from their lane quarter
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc[32], sink = 0;
    int c = 0;
    while (stop < (dual ? 2u : 1u)) {
      tmem_ld32(taddr + c, acc);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) sink ^= acc[j];
      c = (c + 32) & 511;
    }
    if (sink == 0x12345678u) out[blockIdx.x] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace cal

extern "C" int cal_debug_mma_pattern(int pattern, int n, int nk, int iters, int flags, int ctas, long long* out_cycles,
                                     void* stream) {
  using namespace cal;
  CAL_REQUIRE(pattern >= 0 && pattern <= 6 && n >= 16 && n <= 128 && n % 16 == 0 && nk >= 1 && nk <= 4 && iters >= 1 &&
                  ctas >= 1 && out_cycles,
              CAL_E_INVALID, "cal_debug_mma_pattern: bad args");
  const size_t smem = (2 * 256 + 1152) * 128 + 1024;
  CAL_CHECK_CUDA(cudaFuncSetAttribute(mma_pattern_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  mma_pattern_probe_kernel<<<ctas, 224, smem, static_cast<cudaStream_t>(stream)>>>(pattern, n, nk, iters, flags, out_cycles);
  CAL_CHECK_CUDA(cudaGetLastError());
  return CAL_OK;
}
