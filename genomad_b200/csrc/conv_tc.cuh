// K2/K3: causal Conv1D(128->128, k=6) + LeakyReLU, and the IGLOO value projection y @ w_v with
// MaxPool1D(8), as ONE persistent tcgen05 kernel template.
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
// term is ~2^-22).  tools/precision_study.py shows why a single TF32/fp16 pass is not enough for the
// 1e-4 parity bar (1.4e-4 worst case for the convs, 7e-4 for w_v) while this recipe gives ~7e-6.
//
// Tiling: one tile = 128 consecutive positions of one window (47 tiles per window) x all 128 output
// channels; K = 6 taps x 128 input channels.  The activation tensor is [n][5997][256] fp16 (hi | lo
// per row).  One TMA box of 136 rows x 64 channels per (plane, K-half) brings the whole 133-row halo
// slab of a tile into shared memory ONCE (SWIZZLE_128B); tap j is the same slab read 'j' rows further
// down, i.e. only the UMMA descriptor start address changes (row j of the slab is position t0-5+j,
// and TMA zero-fills rows with t < 0 or t >= 5997, which is exactly Keras' causal padding).
// Measured on B200 (profiles/r01_bringup.md): the 128B swizzle is a function of the absolute shared-
// memory address, so a descriptor whose start address is 'j' rows (j*128 B) into a 1024B-aligned
// slab reads the rows TMA wrote, with descriptor base_offset = 0 -- no per-tap reload is needed.
// Weights stream through a 4-stage ring of 16 KB stages (one K-half of one [128 out][128 in] fp16
// matrix, K-major), re-packed on the host in consumption order.
//
// Warp roles (256 threads, 1 CTA per SM, persistent over tiles):
//   warp 0 lane 0 : TMA producer (activation slabs, weight stages)
//   warp 1 lane 0 : tcgen05.mma issuer; accumulators live in TMEM (2 sets x {Y,Z} x 128 columns)
//   warp 2        : TMEM allocator
//   warps 4..7    : epilogue (tcgen05.ld -> bias + LeakyReLU -> fp16 hi/lo split -> global;
//                   for the w_v accumulator: 8-row max via warp shuffles -> global)
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace gnm {

constexpr int kTileM       = 128;
constexpr int kTilesPerWin = (kTok + kTileM - 1) / kTileM;      // 47
constexpr int kSlabRows    = 136;                                // 128 + 5 halo rows, rounded to 8
constexpr int kARegion     = kSlabRows * 128;                    // bytes: rows x 128 B (64 fp16)   = 17408
constexpr int kABuf        = 4 * kARegion;                       // hi.k0 hi.k1 lo.k0 lo.k1          = 69632
constexpr int kBStage      = 128 * 128;                          // 128 out-rows x 64 fp16          = 16384
constexpr int kNumBStages  = 4;
constexpr int kConvThreads = 256;
constexpr int kConvSmem    = 2 * kABuf + kNumBStages * kBStage + 2048;   // + bias/barriers + align slack

struct ConvTcParams {
  const float* bias;        // [128] or nullptr
  __half* y_out;            // [n][5997][256] or nullptr
  float* q_out;             // [n][749][128] or nullptr
  int n_tiles;              // n_windows * 47
  int experiment;           // timing experiments only (results become wrong): 1 = no tap row shift, 2 = no epilogue stores
  long long* dbg;           // optional [gridDim.x][8] cycle counters (nullptr = off)
  DeviceStatus* status;
};

template <int kNTaps, bool kWv>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
               const ConvTcParams p) {
  constexpr int kConvStages = kNTaps * 4;
  constexpr int kStages = kConvStages + (kWv ? 4 : 0);
  constexpr uint32_t kIdesc = umma_idesc_f16(kTileM, kC);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;                                   // 2 x kABuf
  uint8_t* s_b = smem + 2 * kABuf;                       // kNumBStages x kBStage
  float* s_bias = reinterpret_cast<float*>(s_b + kNumBStages * kBStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + kC);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* b_full = bars + 4;        // [4]
  uint64_t* b_empty = bars + 8;       // [4]
  uint64_t* acc_full = bars + 12;     // [2]
  uint64_t* acc_empty = bars + 14;    // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x < kC) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_act);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kNumBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
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

  if (warp == 0 && lane == 0) {
    // ===================================================================== TMA producer
    uint32_t bcount = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int w = tile / kTilesPerWin;
      const int t0 = (tile - w * kTilesPerWin) * kTileM;
      mbar_wait(&a_empty[ab], aphase ^ 1, p.status, 100 + ab);
      mbar_arrive_expect_tx(&a_full[ab], kABuf);
#pragma unroll
      for (int r = 0; r < 4; ++r)
        tma_load_3d(s_a + ab * kABuf + r * kARegion, &tm_act, &a_full[ab], r * 64, t0 - 5, w);
      for (int q = 0; q < kStages; ++q, ++bcount) {
        const int s = bcount % kNumBStages;
        const uint32_t bphase = (bcount / kNumBStages) & 1;
        mbar_wait(&b_empty[s], bphase ^ 1, p.status, 110 + s);
        mbar_arrive_expect_tx(&b_full[s], kBStage);
        tma_load_2d(s_b + s * kBStage, &tm_w, &b_full[s], 0, q * 128);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp converged, one elected lane issues)
    uint32_t bcount = 0;
    int it = 0;
    const uint64_t desc0 = umma_desc_sw128(0);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const uint32_t acc_y = tmem_base + ab * 256;
      const uint32_t acc_z = tmem_base + ab * 256 + 128;
      mbar_wait(&acc_empty[ab], aphase ^ 1, p.status, 200 + ab);
      mbar_wait(&a_full[ab], aphase, p.status, 210 + ab);
      tc_fence_after();
      const uint32_t a_base = smem_u32(s_a + ab * kABuf);
      for (int q = 0; q < kStages; ++q, ++bcount) {
        const int s = bcount % kNumBStages;
        const uint32_t bphase = (bcount / kNumBStages) & 1;
        mbar_wait(&b_full[s], bphase, p.status, 220 + s);
        tc_fence_after();
        int arow, w_lo, kh;
        uint32_t acc;
        bool first;
        if (q < kConvStages) {
          arow = q >> 2; w_lo = (q >> 1) & 1; kh = q & 1; acc = acc_y; first = (q == 0);
        } else {
          const int qq = q - kConvStages;
          arow = 5; w_lo = qq >> 1; kh = qq & 1; acc = acc_z; first = (qq == 0);
        }
        if (elect_one()) {
          const uint64_t bdesc = desc0 + (smem_u32(s_b + s * kBStage) >> 4);
          const uint64_t ahi = desc0 + ((a_base + kh * kARegion + arow * 128) >> 4);
          const uint64_t alo = desc0 + ((a_base + (2 + kh) * kARegion + arow * 128) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_f16(acc, ahi + kk * 2, bdesc + kk * 2, kIdesc, (first && kk == 0) ? 0u : 1u);
            if (!w_lo) umma_f16(acc, alo + kk * 2, bdesc + kk * 2, kIdesc, 1u);
          }
          umma_commit(&b_empty[s]);          // stage is free once these MMAs have read it
          if (q == kStages - 1) {
            umma_commit(&a_empty[ab]);       // slab is free once every MMA of the tile is done
            umma_commit(&acc_full[ab]);      // ... and the accumulators are complete
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int wq = warp - 4;               // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int w = tile / kTilesPerWin;
      const int t0 = (tile - w * kTilesPerWin) * kTileM;
      const int t = t0 + wq * 32 + lane;
      mbar_wait(&acc_full[ab], aphase, p.status, 300 + ab);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + ab * 256;
      if (kNTaps > 0) {
        __half* row = p.y_out + (static_cast<size_t>(w) * kTok + (t < kTok ? t : 0)) * kRowHalfs;
#pragma unroll
        for (int c32 = 0; c32 < 4; ++c32) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + c32 * 32, r);
          tmem_wait_ld();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v0 = lrelu(__uint_as_float(r[2 * i]) + s_bias[c32 * 32 + 2 * i]);
            const float v1 = lrelu(__uint_as_float(r[2 * i + 1]) + s_bias[c32 * 32 + 2 * i + 1]);
            __half h0, l0, h1, l1;
            split_f16(v0, h0, l0);
            split_f16(v1, h1, l1);
            hi[i] = pack_h2(h0, h1);
            lo[i] = pack_h2(l0, l1);
          }
          if (t < kTok) {
            uint4* dh = reinterpret_cast<uint4*>(row + c32 * 32);
            uint4* dl = reinterpret_cast<uint4*>(row + kC + c32 * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              dh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              dl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
          }
        }
      }
      if (kWv) {
        const int g = t >> 3;                                  // pooled row; same for 8 adjacent lanes
        const int sub = lane & 7;
        float* qrow = p.q_out + (static_cast<size_t>(w) * kPooled + (g < kPooled ? g : 0)) * kC;
#pragma unroll
        for (int c32 = 0; c32 < 4; ++c32) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + 128 + c32 * 32, r);
          tmem_wait_ld();
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = __uint_as_float(r[i]);
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
            if ((i >> 2) == sub) {
              if ((i & 3) == 0) o.x = v; else if ((i & 3) == 1) o.y = v; else if ((i & 3) == 2) o.z = v; else o.w = v;
            }
          }
          if (g < kPooled) *reinterpret_cast<float4*>(qrow + c32 * 32 + sub * 4) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =================================================================================================
// conv2t: the production Conv1D kernel.  Same math as conv_tc_kernel<6,false>, different schedule:
// one work unit = TWO adjacent 128-position tiles of a window (24 units per window), so every 16 KB
// weight stage fetched from L2 feeds 2x the MMAs (measured on the 1-tile kernel: tensor pipe 36 %
// busy, stalled on the weight ring; profiles/r01_conv_1tile_ncu.md).
//
//   * activation slab: 2 x 136-row TMA boxes per (plane, K-half) = rows t0-5 .. t0+266; tile b's
//     taps start 128 rows further down the same slab.
//   * the slab is "double-buffered by K-half": stages are ordered K-half-major (all taps of
//     channels 0..63, then all taps of channels 64..127), so the k0 regions are free again after
//     the first 12 stages and are reloaded for the NEXT unit while the k1 stages run, and vice versa.
//   * separate producer threads for activations (warp 3) and weights (warp 0) so neither blocks the other;
//     weight ring = 5 stages.
//   * accumulators: 2 sets x {tile a, tile b} x 128 TMEM columns = all 512 columns, so the epilogue
//     of unit u overlaps the MMAs of unit u+1.
// =================================================================================================
constexpr int kUnitsPerWin  = (kTilesPerWin + 1) / 2;             // 24
constexpr int kSlab2Rows    = 2 * kSlabRows;                      // 272
constexpr int kA2Region     = kSlab2Rows * 128;                   // 34816 B
constexpr int kA2Bytes      = 4 * kA2Region;                      // 139264 B (hi.k0 hi.k1 lo.k0 lo.k1)
constexpr int kNumB2Stages  = 5;
constexpr int kConv2tSmem   = kA2Bytes + kNumB2Stages * kBStage + 2048;
constexpr int kConv2tStages = 24;                                 // (K-half, tap, weight hi/lo)

__global__ void __launch_bounds__(kConvThreads, 1)
conv2t_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
              const ConvTcParams p) {
  constexpr uint32_t kIdesc = umma_idesc_f16(kTileM, kC);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;                                   // 4 regions x 272 rows x 128 B
  uint8_t* s_b = smem + kA2Bytes;                        // 5 x 16 KB
  float* s_bias = reinterpret_cast<float*>(s_b + kNumB2Stages * kBStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + kC);
  uint64_t* a_full = bars;            // [2]  per K-half
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* b_full = bars + 4;        // [5]
  uint64_t* b_empty = bars + 9;       // [5]
  uint64_t* acc_full = bars + 14;     // [2]
  uint64_t* acc_empty = bars + 16;    // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_tiles;      // for this kernel n_tiles carries n_windows * 24

  if (threadIdx.x < kC) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_act);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kNumB2Stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
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
    uint32_t bcount = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      for (int q = 0; q < kConv2tStages; ++q, ++bcount) {
        const int s = bcount % kNumB2Stages;
        const uint32_t bphase = (bcount / kNumB2Stages) & 1;
        mbar_wait(&b_empty[s], bphase ^ 1, p.status, 110 + s);
        mbar_arrive_expect_tx(&b_full[s], kBStage);
        tma_load_2d(s_b + s * kBStage, &tm_w, &b_full[s], 0, q * 128);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // The whole warp runs this loop converged; one elected lane issues.  Descriptors are built once
    // and advanced by adding to their low word (start address >> 4), so only a couple of uniform
    // integer ops separate consecutive tcgen05.mma instructions.
    uint32_t bcount = 0;
    int it = 0;
    const uint64_t desc0 = umma_desc_sw128(0);
    const uint32_t a_base = smem_u32(s_a);
    const uint32_t b_base = smem_u32(s_b);
    long long w_acc = 0, w_a = 0, w_b = 0, tq;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t accphase = (it >> 1) & 1;
      const uint32_t aph = it & 1;
      const uint32_t acc0 = tmem_base + as * 256;          // tile a; tile b = +128
      tq = clock64();
      mbar_wait(&acc_empty[as], accphase ^ 1, p.status, 200 + as);
      w_acc += clock64() - tq;
      for (int q = 0; q < kConv2tStages; ++q, ++bcount) {
        const int kh = q / 12, r = q - kh * 12, tap = r >> 1, w_lo = r & 1;
        if (r == 0) { tq = clock64(); mbar_wait(&a_full[kh], aph, p.status, 210 + kh); w_a += clock64() - tq; }
        const int s = bcount % kNumB2Stages;
        const uint32_t bphase = (bcount / kNumB2Stages) & 1;
        tq = clock64();
        mbar_wait(&b_full[s], bphase, p.status, 220 + s);
        w_b += clock64() - tq;
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bdesc = desc0 + ((b_base + s * kBStage) >> 4);
          const int arow = (p.experiment & 1) ? 0 : tap;
          const uint64_t ahi = desc0 + ((a_base + kh * kA2Region + arow * 128) >> 4);
          const uint64_t alo = (p.experiment & 8) ? ahi : desc0 + ((a_base + (2 + kh) * kA2Region + arow * 128) >> 4);
          constexpr uint32_t kTileStep = (kTileM * 128) >> 4;      // tile b starts 128 rows further down
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            const uint32_t acc = acc0 + tile * 128;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16(acc, ahi + tile * kTileStep + kk * 2, bdesc + kk * 2, kIdesc, (q == 0 && kk == 0) ? 0u : 1u);
              if (!w_lo) umma_f16(acc, alo + tile * kTileStep + kk * 2, bdesc + kk * 2, kIdesc, 1u);
            }
          }
          umma_commit(&b_empty[s]);
          if (r == 11) umma_commit(&a_empty[kh]);          // this K-half of the slab is no longer needed
          if (q == kConv2tStages - 1) umma_commit(&acc_full[as]);
        }
        __syncwarp();
      }
    }
    if (p.dbg && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 8;
      d[0] = clock64() - t_begin; d[1] = w_acc; d[2] = w_a; d[3] = w_b; d[4] = it;
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int wq = warp - 4;
    int it = 0;
    long long w_full = 0, tq;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t accphase = (it >> 1) & 1;
      const int w = unit / kUnitsPerWin;
      const int t0 = (unit - w * kUnitsPerWin) * (2 * kTileM);
      tq = clock64();
      mbar_wait(&acc_full[as], accphase, p.status, 300 + as);
      w_full += clock64() - tq;
      tc_fence_after();
#pragma unroll 1
      for (int tile = 0; tile < 2; ++tile) {
        const int t = t0 + tile * kTileM + wq * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * 256 + tile * 128;
        __half* row = p.y_out + (static_cast<size_t>(w) * kTok + (t < kTok ? t : 0)) * kRowHalfs;
#pragma unroll
        for (int c32 = 0; c32 < 4; ++c32) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + c32 * 32, r);
          tmem_wait_ld();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v0 = lrelu(__uint_as_float(r[2 * i]) + s_bias[c32 * 32 + 2 * i]);
            const float v1 = lrelu(__uint_as_float(r[2 * i + 1]) + s_bias[c32 * 32 + 2 * i + 1]);
            __half h0, l0, h1, l1;
            split_f16(v0, h0, l0);
            split_f16(v1, h1, l1);
            hi[i] = pack_h2(h0, h1);
            lo[i] = pack_h2(l0, l1);
          }
          if (t < kTok && !(p.experiment & 2)) {
            uint4* dh = reinterpret_cast<uint4*>(row + c32 * 32);
            uint4* dl = reinterpret_cast<uint4*>(row + kC + c32 * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              dh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              dl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
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
      d[5] = clock64() - t_begin; d[6] = w_full;
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
