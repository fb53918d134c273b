// K2/K3: causal Conv1D(128->128, k=6) + LeakyReLU, and the IGLOO value projection (y @ w_v, MaxPool1D(8)),
// as ONE persistent tcgen05 kernel template (conv_t_kernel<false> / conv_t_kernel<true>).
//
// Reference semantics:
//   Conv1D x2 + LeakyReLU(0.1)      genomad/neural_network/igloo.py:64-72
//         y'[t,:] = lrelu(b + sum_{j=0..5, t-5+j>=0} y[t-5+j,:] @ W[j])      W: [6][128 in][128 out]
//   y_proj = y @ w_v, MaxPool1D(8)  genomad/neural_network/igloo.py:208-210
//         q[g,:] = max_{r<8} (y[8g+r,:] @ Wv)     g < 749 (positions 5992..5996 are dropped)
//
// Arithmetic: fp32-equivalent "3-pass split".  Every fp32 operand x is carried as two fp16 numbers
// hi = fp16(x), lo = fp16(x - hi) (|x - hi - lo| <~ 2^-22 |x|).  A product A*B is evaluated on the
// tensor cores as Ahi*Bhi + Alo*Bhi + Ahi*Blo with fp32 accumulation in TMEM (the dropped Alo*Blo
// term is ~2^-22).  tools/precision_study.py (profiles/r01_precision_study.md) shows why a single
// TF32/fp16 pass is not enough for the 1e-4 parity bar (1.4e-4 worst case for the convs, 7e-4 for w_v)
// while this recipe gives ~7e-6.
//
// Data layout.  Activations are [n][5997][256] fp16 (128 "hi" | 128 "lo" halves per 512-byte row).
// One work unit = 256 consecutive positions of one window (24 units per window).  For each
// (plane, K-half) two TMA boxes of 136 rows x 64 channels bring rows t0-5 .. t0+266 of the window into
// a 272-row SWIZZLE_128B slab ONCE; conv tap j is the same slab read j rows further down -- only the
// UMMA descriptor start address changes (row j is position t0-5+j; TMA zero-fills rows with t < 0 or
// t >= 5997, which is exactly Keras' causal padding).  Measured on B200 (profiles/r01_bringup.md): the
// 128B swizzle is a function of the absolute shared-memory address, so a descriptor may start at any
// 128-byte row of a 1024B-aligned slab with base_offset = 0.
//
// Operand roles ("transposed" formulation).  Measured (tools/mma_microbench.cu,
// profiles/r01_mma_microbench.md): a cta_group::1 M=128 x N=128 x K=16 UMMA needs 128 B/cycle of
// shared-memory operand bandwidth -- all of it -- and in the real kernel (TMA fills and epilogue
// traffic share the banks) ran at ~109 instead of 64 cycles; N=256 needs 96 B/cycle.  The model has
// only 128 output channels, so N=256 comes from swapping the roles:
//
//     D^T[cout (M=128 TMEM lanes)][position (N=256 TMEM columns)] += W_tap^T[cout][cin] * Y[position + tap][cin]
//
//   A operand = one 16 KB weight stage [128 cout][64 cin] fp16 K-major SWIZZLE_128B (streamed by TMA through
//               a 4-stage ring, host-packed in consumption order), shared by both tiles of the unit;
//   B operand = 256 consecutive rows of the activation slab.
//
// Schedule.  Stages are ordered K-half-major (all taps of channels 0..63, then 64..127) so each half of
// the slab is free after 12 stages and is reloaded for the NEXT unit while the other half is in use;
// activations and weights have separate producer threads; 2 accumulator sets x 256 TMEM columns let the
// epilogue of unit u overlap the MMAs of unit u+1.  The MMA warp stays converged and issues through
// elect.sync so consecutive tcgen05.mma stay on the uniform datapath (the first version issued from a
// divergent single lane and spent ~140 cycles per MMA in the compiler's waterfall loop).
//
// Warp roles (256 threads, 1 CTA per SM, persistent over units):
//   warp 0 lane 0 : weight producer (TMA)          warp 3 lane 0 : activation producer (TMA)
//   warp 1        : tcgen05.mma issuer             warp 2        : TMEM allocator
//   warps 4..7    : epilogue.  A thread owns one output CHANNEL (TMEM lane) and reads 32 positions at a time.
//       conv : bias + LeakyReLU + fp16 hi/lo split, transposed through a 16 KB shared staging tile
//              ([32 positions][256 halves]) and written back as full 512-byte activation rows;
//       w_v  : the max over 8 consecutive positions is a max over 8 registers (no shuffles); a warp writes
//              128 contiguous bytes of q[g][:] per pooled row.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace gnm {

constexpr int kTileM       = 128;
constexpr int kUnitsPerWin = (kTok + 2 * kTileM - 1) / (2 * kTileM);   // 24
constexpr int kSlabRows    = 136;                                // rows per TMA box
constexpr int kARegion     = kSlabRows * 128;                    // bytes per box: rows x 128 B (64 fp16)  = 17408
constexpr int kA2Region    = 2 * kARegion;                       // 272-row slab region                    = 34816
constexpr int kA2Bytes     = 4 * kA2Region;                      // hi.k0 hi.k1 lo.k0 lo.k1                = 139264
constexpr int kBStage      = 128 * 128;                          // one weight stage: 128 rows x 64 fp16   = 16384
constexpr int kConvThreads = 256;
constexpr int kConvStages  = 24;                                 // conv: (K-half, tap, weight hi/lo)
constexpr int kWvStages    = 4;                                  // w_v : (K-half, weight hi/lo)

struct ConvTcParams {
  const float* bias;        // [128] (conv) or nullptr (w_v)
  __half* y_out;            // [n][5997][256] (conv) or nullptr
  float* q_out;             // [n][749][128] (w_v) or nullptr
  int n_tiles;              // number of work units = n_windows * 24
  int experiment;           // timing experiments only (results become wrong): 2 = no epilogue global stores
  long long* dbg;           // optional [gridDim.x][8] cycle counters (nullptr = off)
  DeviceStatus* status;
};

constexpr int kTWStages   = 4;                                   // weight ring depth (16 KB each)
constexpr int kTStageTile = 32 * kRowHalfs * 2;                  // 16 KB epilogue staging tile
constexpr int kConvTSmem  = kA2Bytes + kTWStages * kBStage + kTStageTile + 2048;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool kWvMode>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_t_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
              const ConvTcParams p) {
  constexpr int kSPK = kWvMode ? 2 : 12;                          // stages per K-half
  constexpr int kStagesU = 2 * kSPK;                              // stages per unit
  constexpr uint32_t kIdesc = umma_idesc_f16(128, 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;                                   // activation slab: 4 regions x 272 rows x 128 B
  uint8_t* s_w = smem + kA2Bytes;                        // weight ring
  uint8_t* s_stage = s_w + kTWStages * kBStage;          // epilogue staging tile (conv mode)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + kTStageTile);
  uint64_t* a_full = bars;            // [2]  per K-half
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* w_full = bars + 4;        // [4]
  uint64_t* w_empty = bars + 8;       // [4]
  uint64_t* acc_full = bars + 12;     // [2]
  uint64_t* acc_empty = bars + 14;    // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_tiles;      // n_windows * 24

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_act);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kTWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(s_tmem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 3 && lane == 0) {
    // ===================================================================== activation producer
    int it = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int w = unit / kUnitsPerWin;
      const int t0 = (unit - w * kUnitsPerWin) * (2 * kTileM);
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        mbar_wait(&a_empty[kh], ph ^ 1, p.status, 100 + kh);
        mbar_arrive_expect_tx(&a_full[kh], 2 * kA2Region);
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          uint8_t* dst = s_a + (plane * 2 + kh) * kA2Region;
          const int c0 = plane * kC + kh * 64;
          tma_load_3d(dst, &tm_act, &a_full[kh], c0, t0 - 5, w);
          tma_load_3d(dst + kARegion, &tm_act, &a_full[kh], c0, t0 - 5 + kSlabRows, w);
        }
      }
    }
  } else if (warp == 0 && lane == 0) {
    // ===================================================================== weight producer
    uint32_t wcount = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      for (int q = 0; q < kStagesU; ++q, ++wcount) {
        const int s = wcount % kTWStages;
        const uint32_t wphase = (wcount / kTWStages) & 1;
        mbar_wait(&w_empty[s], wphase ^ 1, p.status, 110 + s);
        mbar_arrive_expect_tx(&w_full[s], kBStage);
        tma_load_2d(s_w + s * kBStage, &tm_w, &w_full[s], 0, q * 128);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (converged warp, elected lane issues)
    uint32_t wcount = 0;
    int it = 0;
    const uint64_t desc0 = umma_desc_sw128(0);
    const uint32_t a_base = smem_u32(s_a);
    const uint32_t w_base = smem_u32(s_w);
    long long w_acc = 0, w_a = 0, w_w = 0, tq;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t accphase = (it >> 1) & 1;
      const uint32_t aph = it & 1;
      const uint32_t acc = tmem_base + as * 256;
      tq = clock64();
      mbar_wait(&acc_empty[as], accphase ^ 1, p.status, 200 + as);
      w_acc += clock64() - tq;
      for (int q = 0; q < kStagesU; ++q, ++wcount) {
        const int kh = q / kSPK, r = q - kh * kSPK;
        const int tap = kWvMode ? 5 : (r >> 1), w_lo = r & 1;
        if (r == 0) { tq = clock64(); mbar_wait(&a_full[kh], aph, p.status, 210 + kh); w_a += clock64() - tq; }
        const int s = wcount % kTWStages;
        const uint32_t wphase = (wcount / kTWStages) & 1;
        tq = clock64();
        mbar_wait(&w_full[s], wphase, p.status, 220 + s);
        w_w += clock64() - tq;
        tc_fence_after();
        if (elect_one()) {
          const uint64_t wdesc = desc0 + ((w_base + s * kBStage) >> 4);                              // A: weights
          const uint64_t yhi = desc0 + ((a_base + kh * kA2Region + tap * 128) >> 4);                 // B: activations
          const uint64_t ylo = desc0 + ((a_base + (2 + kh) * kA2Region + tap * 128) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_f16(acc, wdesc + kk * 2, yhi + kk * 2, kIdesc, (q == 0 && kk == 0) ? 0u : 1u);
            if (!w_lo) umma_f16(acc, wdesc + kk * 2, ylo + kk * 2, kIdesc, 1u);
          }
          umma_commit(&w_empty[s]);
          if (r == kSPK - 1) umma_commit(&a_empty[kh]);      // this K-half of the slab is no longer needed
          if (q == kStagesU - 1) umma_commit(&acc_full[as]);
        }
        __syncwarp();
      }
    }
    if (p.dbg && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 8;
      d[0] = clock64() - t_begin; d[1] = w_acc; d[2] = w_a; d[3] = w_w; d[4] = it;
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int wq = warp - 4;                                   // TMEM lane quarter = channels 32*wq .. 32*wq+31
    const int ch = wq * 32 + lane;
    const float bias = kWvMode ? 0.f : p.bias[ch];
    const int te = threadIdx.x - 128;                          // 0..127 within the epilogue group
    int it = 0;
    long long w_full_c = 0, tq;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t accphase = (it >> 1) & 1;
      const int w = unit / kUnitsPerWin;
      const int t0 = (unit - w * kUnitsPerWin) * (2 * kTileM);
      tq = clock64();
      mbar_wait(&acc_full[as], accphase, p.status, 300 + as);
      w_full_c += clock64() - tq;
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * 256;
#pragma unroll 1
      for (int c32 = 0; c32 < 8; ++c32) {
        const int p0 = t0 + c32 * 32;                          // first position of this chunk
        if (p0 >= kTok) break;                                 // uniform: the rest of the unit is past the window end
        uint32_t r[32];
        tmem_ld_32x32(lane_addr + c32 * 32, r);
        tmem_wait_ld();
        if (kWvMode) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float m = __uint_as_float(r[8 * g]);
#pragma unroll
            for (int k = 1; k < 8; ++k) m = fmaxf(m, __uint_as_float(r[8 * g + k]));
            const int gg = (p0 >> 3) + g;
            if (gg < kPooled) p.q_out[(static_cast<size_t>(w) * kPooled + gg) * kC + ch] = m;
          }
        } else {
          __half* st_hi = reinterpret_cast<__half*>(s_stage) + ch;
          __half* st_lo = st_hi + kC;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = lrelu(__uint_as_float(r[i]) + bias);
            __half h, l;
            split_f16(v, h, l);
            st_hi[i * kRowHalfs] = h;
            st_lo[i * kRowHalfs] = l;
          }
          named_bar_sync(1, 128);                              // staging tile complete
          // each warp writes back 8 full rows (512 B = 32 lanes x 16 B): perfectly coalesced
          uint4 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            v[k] = *reinterpret_cast<const uint4*>(s_stage + (wq * 8 + k) * (kRowHalfs * 2) + lane * 16);
          named_bar_sync(2, 128);                              // staging tile may be overwritten
          if (!(p.experiment & 2)) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int t = p0 + wq * 8 + k;
              if (t < kTok)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.y_out) +
                                          (static_cast<size_t>(w) * kTok + t) * (kRowHalfs * 2) + lane * 16) = v[k];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
    if (p.dbg && warp == 4 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 8;
      d[5] = clock64() - t_begin; d[6] = w_full_c;
    }
    (void)te;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace gnm
