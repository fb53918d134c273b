// K2/K3: causal Conv1D(128->128, k=6) + LeakyReLU, and the IGLOO value projection (y @ w_v, MaxPool1D(8)),
// as ONE persistent tcgen05 kernel template (conv_t_kernel<false> / conv_t_kernel<true>).
//
// Reference semantics:
//   Conv1D x2 + LeakyReLU(0.1)      genomad/neural_network/igloo.py:64-72
//         y'[t,:] = lrelu(b + sum_{j=0..5, t-5+j>=0} y[t-5+j,:] @ W[j])      W: [6][128 in][128 out]
//   y_proj = y @ w_v, MaxPool1D(8)  genomad/neural_network/igloo.py:208-210
//         q[g,:] = max_{r<8} (y[8g+r,:] @ Wv)     g < 749 (positions 5992..5996 are dropped)
//
// Arithmetic: fp32-equivalent split.  Every operand x is carried as hi = fp16(x) plus a correction
// lo = x - hi, and A*B is evaluated on the tensor cores as Ahi*Bhi + Alo*Bhi + Ahi*Blo with fp32
// accumulation in ONE TMEM accumulator (the dropped Alo*Blo term is ~2^-22).  tools/precision_study.py
// (profiles/r01_precision_study.md) shows why a single TF32/fp16 pass is not enough for the 1e-4 parity
// bar (1.4e-4 worst case for the convs, 7e-4 for w_v).
//   * w_v (feeds max-pool -> mean -> BatchNorm with a 31x gain): all three passes in fp16 (kind::f16), ~7e-6.
//   * convs: the main pass Ahi*Bhi in fp16; the two correction passes only need ~8 bits, so they run in
//     e4m3 (kind::f8f6f4, K = 32 per instruction: twice the rate) on pre-scaled operands.  All passes are
//     scaled to a common 2^S so they add up in the same accumulator, and the epilogue multiplies by 2^-S:
//         main  : (32*Ahi)            x fp16(Whi * 2^d)          S = 5 + d          (d = 16 for |W| < 0.78)
//         corr1 : e4m3(Alo * 2^12)    x e4m3(Whi * 2^(S-12))
//         corr2 : e4m3(Ahi * 2^7)     x e4m3(Wlo * 2^(S-7))
//     CPU emulation of exactly this recipe: max |dp| 6.5e-6 over 256 worst-case-family windows.
//
// Data layout.  Activation rows are 768 bytes: hi16 | lo16 | e4m3 pairs (lo8[c], hi8[c]) (common.cuh), tensor [n][5997][768 B].
// One work unit = 256 consecutive positions of one window (24 units per window).  The unit's operands
// are four slab REGIONS of 272 rows x 128 bytes (conv: hi16 ch 0-63, hi16 ch 64-127, e4m3 pairs of ch 0-63, of ch 64-127;
// w_v: hi16 and lo16 halves); two TMA boxes of 136 rows bring rows t0-5 .. t0+266 of the window into
// each SWIZZLE_128B region ONCE; conv tap j is the same slab read j rows further down -- only the
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
//   A operand = one 16 KB weight stage, [128 cout][64 cin] fp16 or [128 cout][128 cin] e4m3, K-major
//               SWIZZLE_128B (streamed by TMA through a 5-stage ring, host-packed in consumption order);
//   B operand = 256 consecutive rows of one slab region.
//
// Schedule.  Weight stages are ordered region-major (all 6 taps against region 0, then region 1, ...) so a
// region is free after its 6 stages and is reloaded for the NEXT unit while the other regions are in use;
// activations and weights have separate producer threads; 2 accumulator sets x 256 TMEM columns let the
// epilogue of unit u overlap the MMAs of unit u+1.  The MMA warp stays converged and issues through
// elect.sync so consecutive tcgen05.mma stay on the uniform datapath (the first version issued from a
// divergent single lane and spent ~140 cycles per MMA in the compiler's waterfall loop).
//
// Warp roles (384 threads, 1 CTA per SM, persistent over units):
//   warp 0 lane 0 : weight producer (TMA)          warp 3 lane 0 : activation producer (TMA)
//   warp 1        : tcgen05.mma issuer             warp 2        : TMEM allocator
//   warps 4..11   : epilogue (two warps per TMEM lane quarter, alternating 32-position chunks).  A thread owns one output CHANNEL (TMEM lane) and reads 32 positions at a time.
//       conv : 2^-S scale + bias + LeakyReLU, then the planes the consumer needs (conv2 -> hi16, lo8, hi8 for
//              conv3; conv3 -> hi16, lo16 for w_v / gather) stored straight to global memory: for a fixed position
//              a warp's 32 lanes are 32 consecutive channels = one contiguous 64-byte (fp8: 32-byte) piece of the row
//              (an earlier version transposed through a 16 KB shared staging tile; shared-memory bandwidth is the
//              scarce resource of this kernel, and dropping the tile also paid for a 5th weight stage; a lane-pair
//              exchange that halves the store count was measured and brought nothing);
//       w_v  : the max over 8 consecutive positions is a max over 8 registers (no shuffles); a warp writes
//              128 contiguous bytes of q[g][:] per pooled row.
//
// Round 2: (i) the two e4m3 planes are interleaved as (lo8, hi8) pairs per channel (common.cuh), so the correction passes are one
// K = 256 contraction per tap against weights interleaved the same way and conv2's epilogue needs one 2-byte store per position
// instead of two 1-byte stores (conv2 2.30 -> 2.21 ms, profiles/r02_ab_pairs_interleaved.log); (ii) the epilogue tracks the largest
// |Y| it produced and raises DeviceStatus::act_overflow beyond the operand formats' range; (iii) launching the persistent grid as
// thread-block clusters of 2 / 4 (hoping L2 would merge the CTAs' identical weight-stage requests) was measured and is slower
// (2.29 -> 2.40 / 2.43-2.49 ms, option conv_cluster, profiles/r02_ab_conv_cluster.log); (iv) the steady state of the whole step is
// power-bound (1 kW cap, ~1.45-1.5 GHz during this kernel): tensor pipe 63-68 % active at that clock.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace gnm {

constexpr int kTileM       = 128;
constexpr int kUnitsPerWin = (kTok + 2 * kTileM - 1) / (2 * kTileM);   // 24
constexpr int kSlabRows    = 136;                                // rows per TMA box
constexpr int kARegion     = kSlabRows * 128;                    // bytes per box: rows x 128 B            = 17408
constexpr int kA2Region    = 2 * kARegion;                       // 272-row slab region                    = 34816
constexpr int kNumRegions  = 4;
constexpr int kA2Bytes     = kNumRegions * kA2Region;            //                                        = 139264
constexpr int kBStage      = 128 * 128;                          // one weight stage: 128 rows x 128 B     = 16384
constexpr int kConvThreads = 384;                                // 4 control warps + 8 epilogue warps
constexpr int kConvStages  = 24;                                 // conv: (region, tap)
constexpr int kWvStages    = 4;                                  // w_v : (K-half, weight hi/lo)
constexpr int kTWStages    = 5;                                  // weight ring depth (16 KB each)
constexpr int kConvTSmem   = kA2Bytes + kTWStages * kBStage + 2048;

struct ConvTcParams {
  const float* bias;        // [128] (conv) or nullptr (w_v)
  uint8_t* y_out;           // [n][5997][768 B] (conv) or nullptr
  float* q_out;             // [n][749][128] (w_v) or nullptr
  float out_scale;          // conv: 2^-S (undoes the common operand scaling); w_v: 1/32 (activation scale)
  int out_fp8;              // conv: 1 = write hi16 + lo8 + hi8 (consumer is a conv), 0 = write hi16 + lo16
  int n_tiles;              // number of work units = n_windows * 24
  int experiment;           // timing experiments only (results become wrong): 2 = no epilogue global stores
  long long* dbg;           // optional [gridDim.x][8] cycle counters (nullptr = off)
  DeviceStatus* status;
};

// byte offset inside the 768-byte activation row of the data that fills slab region r
template <bool kWvMode> __device__ __forceinline__ constexpr int region_src(int r) {
  return kWvMode ? (r == 0 ? kOffHi16 : r == 1 ? kOffHi16 + 128 : r == 2 ? kOffLo16 : kOffLo16 + 128)
                 : (r == 0 ? kOffHi16 : r == 1 ? kOffHi16 + 128 : r == 2 ? kOffP8 : kOffP8 + 128);
}

template <bool kWvMode>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_t_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
              const ConvTcParams p) {
  constexpr int kStagesU = kWvMode ? kWvStages : kConvStages;     // stages per unit
  constexpr uint32_t kIdescFull = umma_idesc_f16(128, 256);
  // The last unit of a window holds only 5997 - 23 * 256 = 109 positions: N = 112 columns instead of 256 saves 56 % of that
  // unit's tensor time (2.3 % of the kernel).  Columns 112.. of the accumulator then keep stale values; the epilogue never
  // stores positions >= 5997 (conv) / pooled rows >= 749 (w_v), which is everything from column 109 on.
  constexpr int kTailCols = ((kTok - (kUnitsPerWin - 1) * 2 * kTileM) + 15) / 16 * 16;          // 112
  constexpr uint32_t kIdescTail = umma_idesc_f16(128, kTailCols);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;                                   // activation slab: 4 regions x 272 rows x 128 B
  uint8_t* s_w = smem + kA2Bytes;                        // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + kTWStages * kBStage);
  uint64_t* a_full = bars;            // [4]  per region
  uint64_t* a_empty = bars + 4;       // [4]
  uint64_t* w_full = bars + 8;        // [5]
  uint64_t* w_empty = bars + 13;      // [5]
  uint64_t* acc_full = bars + 18;     // [2]
  uint64_t* acc_empty = bars + 20;    // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_tiles;      // n_windows * 24

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_act);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < kNumRegions; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kTWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
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
    const uint64_t pol = l2_policy_evict_first();          // activations stream through L2 once
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int w = unit / kUnitsPerWin;
      const int t0 = (unit - w * kUnitsPerWin) * (2 * kTileM);
#pragma unroll
      for (int k = 0; k < kNumRegions; ++k) {
        const int r = kWvMode ? (k == 0 ? 0 : k == 1 ? 2 : k == 2 ? 1 : 3) : k;      // order in which the MMAs need them
        mbar_wait(&a_empty[r], ph ^ 1, p.status, 100 + r);
        mbar_arrive_expect_tx(&a_full[r], kA2Region);
        uint8_t* dst = s_a + r * kA2Region;
        tma_load_3d_hint(dst, &tm_act, &a_full[r], region_src<kWvMode>(r), t0 - 5, w, pol);
        tma_load_3d_hint(dst + kARegion, &tm_act, &a_full[r], region_src<kWvMode>(r), t0 - 5 + kSlabRows, w, pol);
      }
    }
  } else if (warp == 0 && lane == 0) {
    // ===================================================================== weight producer
    uint32_t wcount = 0;
    const uint64_t pol = l2_policy_evict_last();           // the 384 KB of weights are re-read by every CTA for every unit
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      for (int q = 0; q < kStagesU; ++q, ++wcount) {
        const int s = wcount % kTWStages;
        const uint32_t wphase = (wcount / kTWStages) & 1;
        mbar_wait(&w_empty[s], wphase ^ 1, p.status, 110 + s);
        mbar_arrive_expect_tx(&w_full[s], kBStage);
        tma_load_2d_hint(s_w + s * kBStage, &tm_w, &w_full[s], 0, q * 128, pol);
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
      const uint32_t kIdesc = (unit % kUnitsPerWin == kUnitsPerWin - 1) ? kIdescTail : kIdescFull;
      tq = clock64();
      mbar_wait(&acc_empty[as], accphase ^ 1, p.status, 200 + as);
      w_acc += clock64() - tq;
      for (int q = 0; q < kStagesU; ++q, ++wcount) {
        // conv: stage q = (region q/6, tap q%6); regions 0,1 are fp16 K-halves, 2 / 3 = e4m3 pairs (lo8, hi8) of channels 0-63 / 64-127
        //       against weights interleaved the same way (Whi8, Wlo8)
        // w_v : stage q = (K-half q/2, weight hi/lo q%2); hi-weight stages multiply both the hi16 and the lo16 region
        const int reg = kWvMode ? (q >> 1) : (q / 6);
        const int tap = kWvMode ? 5 : (q - reg * 6);
        const bool first_use = kWvMode ? ((q & 1) == 0) : (tap == 0);
        if (first_use) {
          tq = clock64();
          mbar_wait(&a_full[reg], aph, p.status, 210 + reg);
          if (kWvMode) mbar_wait(&a_full[2 + reg], aph, p.status, 214 + reg);
          w_a += clock64() - tq;
        }
        const int s = wcount % kTWStages;
        const uint32_t wphase = (wcount / kTWStages) & 1;
        tq = clock64();
        mbar_wait(&w_full[s], wphase, p.status, 220 + s);
        w_w += clock64() - tq;
        tc_fence_after();
        if (elect_one()) {
          const uint64_t wdesc = desc0 + ((w_base + s * kBStage) >> 4);                              // A: weights
          const uint64_t y0 = desc0 + ((a_base + reg * kA2Region + tap * 128) >> 4);                 // B: activations
          if (kWvMode) {
            const uint64_t y1 = desc0 + ((a_base + (2 + reg) * kA2Region + tap * 128) >> 4);        // lo16 region
            const bool w_lo = (q & 1) != 0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16(acc, wdesc + kk * 2, y0 + kk * 2, kIdesc, (q == 0 && kk == 0) ? 0u : 1u);
              if (!w_lo) umma_f16(acc, wdesc + kk * 2, y1 + kk * 2, kIdesc, 1u);
            }
            umma_commit(&w_empty[s]);
            if (q == 0) umma_commit(&a_empty[2]);                  // lo16.k0 is only used by stage 0
            if (q == 1) umma_commit(&a_empty[0]);
            if (q == 2) umma_commit(&a_empty[3]);
            if (q == 3) umma_commit(&a_empty[1]);
          } else {
            if (reg < 2) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16(acc, wdesc + kk * 2, y0 + kk * 2, kIdesc, (q == 0 && kk == 0) ? 0u : 1u);
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f8(acc, wdesc + kk * 2, y0 + kk * 2, kIdesc, 1u);
            }
            umma_commit(&w_empty[s]);
            if (tap == 5) umma_commit(&a_empty[reg]);              // this region of the slab is no longer needed
          }
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
    // 8 epilogue warps: warp w may touch TMEM lanes 32*(w%4)..+31 (= 32 output channels); the two warps that share a
    // lane quarter split the unit's eight 32-position chunks between them (even / odd).
    const int wq = warp & 3;                                   // TMEM lane quarter = channels 32*wq .. 32*wq+31
    const int grp = (warp - 4) >> 2;                           // 0: chunks 0,2,4,6   1: chunks 1,3,5,7
    const int ch = wq * 32 + lane;
    const float bias = kWvMode ? 0.f : p.bias[ch];
    const float oscale = p.out_scale;
    int it = 0;
    float amax = 0.f;                                          // largest |Y| this thread produced (range check, common.cuh)
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
      for (int c32 = grp; c32 < 8; c32 += 2) {
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
            if (gg < kPooled) p.q_out[(static_cast<size_t>(w) * kPooled + gg) * kC + ch] = m * oscale;
          }
        } else {
          // Direct global stores: for a fixed position the 32 lanes of a warp hold 32 consecutive channels, so one
          // 2-byte store per lane is a contiguous 64-byte piece of the row (32 bytes for an fp8 plane).  No shared-memory
          // staging: the MMAs already use ~75 % of the shared-memory bandwidth for their operands and TMA fills ~23 %.
          uint8_t* rowp = p.y_out + (static_cast<size_t>(w) * kTok + p0) * kRowBytes;
          const bool store = !(p.experiment & 2);
          if (p.out_fp8) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float y0 = kActScale * lrelu(fmaf(__uint_as_float(r[i]), oscale, bias));
              const float y1 = kActScale * lrelu(fmaf(__uint_as_float(r[i + 1]), oscale, bias));
              amax = fmaxf(amax, fmaxf(fabsf(y0), fabsf(y1)));
              const __half2 h = __floats2half2_rn(y0, y1);
              const float2 f = __half22float2(h);
              // e4m3 pair (lo8, hi8) of this channel at each of the two positions: one 2-byte store per position
              const uint16_t e0 = pack_e4m3x2((y0 - f.x) * kLo8Scale, f.x * kHi8Scale);
              const uint16_t e1 = pack_e4m3x2((y1 - f.y) * kLo8Scale, f.y * kHi8Scale);
              if (store && p0 + i < kTok) {
                uint8_t* q = rowp + i * kRowBytes;
                reinterpret_cast<__half*>(q + kOffHi16)[ch] = __low2half(h);
                reinterpret_cast<uint16_t*>(q + kOffP8)[ch] = e0;
              }
              if (store && p0 + i + 1 < kTok) {
                uint8_t* q = rowp + (i + 1) * kRowBytes;
                reinterpret_cast<__half*>(q + kOffHi16)[ch] = __high2half(h);
                reinterpret_cast<uint16_t*>(q + kOffP8)[ch] = e1;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float y0 = kActScale * lrelu(fmaf(__uint_as_float(r[i]), oscale, bias));
              const float y1 = kActScale * lrelu(fmaf(__uint_as_float(r[i + 1]), oscale, bias));
              amax = fmaxf(amax, fmaxf(fabsf(y0), fabsf(y1)));
              __half2 h, l;
              split2_f16(y0, y1, h, l);
              if (store && p0 + i < kTok) {
                uint8_t* q = rowp + i * kRowBytes;
                reinterpret_cast<__half*>(q + kOffHi16)[ch] = __low2half(h);
                reinterpret_cast<__half*>(q + kOffLo16)[ch] = __low2half(l);
              }
              if (store && p0 + i + 1 < kTok) {
                uint8_t* q = rowp + (i + 1) * kRowBytes;
                reinterpret_cast<__half*>(q + kOffHi16)[ch] = __high2half(h);
                reinterpret_cast<__half*>(q + kOffLo16)[ch] = __high2half(l);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
    if (!kWvMode) flag_act_overflow(p.status, amax, p.out_fp8 ? kHi8Limit : kF16Limit, p.out_fp8 ? 2 : 3);
    if (p.dbg && warp == 4 && lane == 0) {   // (group 0's view)
      long long* d = p.dbg + blockIdx.x * 8;
      d[5] = clock64() - t_begin; d[6] = w_full_c;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace gnm
