// K3 + K4 fused: the IGLOO value projection q = maxpool8(y @ w_v) on tcgen05 AND the IGLOO patch gather
// mpi[p] = sum_k y[P[p,k],:] . Wf[p,k,:], in ONE pass over the activations (round 2; replaces conv_t_kernel<true> +
// patch_stream_kernel, which each streamed the same hi16/lo16 planes from HBM: 3.1 + 2.4 GB per 1024 windows).
//
// Reference semantics (genomad/neural_network/igloo.py:190-214):
//   mpi   = gather_nd(transpose(y), patches) * w_mult, reshaped, @ w_summer + w_bias          (lines 192-206)
//   y_proj = y @ w_v, MaxPool1D(8)                                                              (lines 208-210)
//
// Work decomposition: POSITION BANDS.  A unit is one band of 32 consecutive positions of 8 consecutive windows
// (256 activation rows = one N = 256 tensor-core tile; 188 bands x ceil(n/8) window groups).  Units are numbered
// band-major and every CTA owns a contiguous range of them, so a CTA stays on one band (at most three) for the whole
// launch while the grid as a whole sweeps the windows front to back.  That is what makes the gather cheap here: the
// ~45 (patch, slot) entries whose position falls into the CTA's band use the same 23 KB of folded weights for every unit
// (L1 / L2 resident), instead of every window re-reading all 4.3 MB.
//
//   * value projection: as conv_t_kernel<true> (3 fp16 passes Ahi*Whi + Alo*Whi + Ahi*Wlo into one TMEM accumulator,
//     operands swapped so that D^T[cout][row]), but the four slab regions [8 windows][32 rows][128 B] come from ONE 3-D TMA
//     box each, and the 64 KB of w_v weight stages stay resident in shared memory for the whole launch (the old kernel
//     re-streamed them for every unit: a third of its L2 -> SM traffic).
//   * patch gather: the 20 consumer warps that drain the accumulators first run the band's entries (<= 3 per warp) against the
//     unit's 8 windows, reading the rows from the SLAB IN SHARED MEMORY that the TMA engine filled for the MMAs.  The gather
//     follows the MMAs' K-half order so slab regions are still recycled one K-half at a time: pass 0 takes channels 0..63 from
//     the hi16.k0 / lo16.k0 regions, pass 1 channels 64..127 from hi16.k1 / lo16.k1; a region goes back to the producer when
//     both the tensor core (tcgen05.commit) and every consumer warp have arrived on its "empty" barrier (count 21).
//     One LDS.128 per lane covers two windows of one entry: lanes 0-7 / 8-15 read the 128-byte hi / lo row of window 2i
//     (16-byte chunk j of a row sits at chunk j ^ (row & 7): TMA's 128-byte swizzle), lanes 16-31 the same for window 2i+1;
//     each quarter-warp reads one whole row, so the access is bank-conflict free.  A row that the warp's previous entry
//     already pulled out is not read again (entries are sorted by position; 46 % share their row with a neighbour).  The four
//     partial sums per lane are reduced by a 5-shuffle transposing butterfly (fixed order -> deterministic); lanes
//     0,4,..,28 end up with the 8 windows' values and write part_t[slot][window] after pass 1 (the pass-0 halves wait in
//     registers; bands with more than 60 entries take a generic path that parks them in part_t).
//     patch_finish_t_kernel adds a patch's four slots in fixed order k = 0..3 plus the bias, as before.
//   * what bounds it (profiles/r02_wv_gather_ncu.md): not HBM (48 % of peak) and not the gather's FMAs -- reading
//     the rows WITHOUT any arithmetic costs the same -- but the shared-memory port that UMMA operand reads (288 KB per unit),
//     TMA fills (131 KB) and the gather's LDS (~150 KB) share: ~650 cycles per (entry, K-half) in a consumer warp.
//
// Warp roles (768 threads, 1 CTA per SM):  warp 0 lane 0: weight loader (once) | warp 1: tcgen05.mma issuer |
// warp 2: TMEM allocator | warp 3 lane 0: activation producer (TMA) | warps 4..23: patch gather, then epilogue
// (max over 8 positions = max over 8 registers; a warp writes 128 contiguous bytes of q[g][:] per pooled row).
#pragma once
#include <cuda.h>
#include <type_traits>
#include "common.cuh"
#include "conv_t.cuh"
#include "igloo.cuh"

namespace gnm {

constexpr int kBandRows   = 32;                                   // positions per band (multiple of the pool size 8)
constexpr int kBandWins   = 8;                                    // windows per unit
constexpr int kNumBands   = (kTok + kBandRows - 1) / kBandRows;   // 188 (the last band holds 13 valid rows)
constexpr int kWgRegion   = kBandRows * kBandWins * 128;          // bytes per slab region (256 rows x 128 B) = 32768
constexpr int kWgBufs     = 4;                                    // region buffers.  The code is a ring (region k of unit `it` lives in buffer
                                                                  // (4 it + k) % kWgBufs): with 5 buffers the TMA can run one region ahead of
                                                                  // the consumers' releases -- measured, no gain (1.15 / 1.26 ms vs 1.05 / 1.14),
                                                                  // so the slab stays at exactly one unit
constexpr int kWgSlab     = kWgBufs * kWgRegion;                  // hi16.k0 | lo16.k0 | hi16.k1 | lo16.k1        = 131072
constexpr int kWgWeights  = kWvStages * kBStage;                  // resident w_v stages                         =  65536
constexpr int kWgWarps    = 20;                                   // consumer warps: patch gather + accumulator epilogue
constexpr int kWgThreads  = (4 + kWgWarps) * 32;                  // 768
constexpr int kWgWarpCap  = 3;                                    // entries per warp on the fast path (60 per band; mean 45)
constexpr int kWgSmem     = kWgSlab + kWgWeights + 2048;          //                                                   = 198656
static_assert(kWgSmem <= 232448, "wv_gather_kernel exceeds the 227 KB of shared memory a CTA may use");
static_assert(kBandRows * kBandWins == 256, "one unit = one N = 256 tile");

struct WvGatherParams {
  float* q_out;                // [n][749][128]
  float out_scale;             // 1/32 (activation scale)
  const int32_t* ent_pos;      // [kGsSlots] position of each entry slot, sorted ascending
  const float* ent_w;          // [kGsSlots][128] folded weights / 32
  const int32_t* band_start;   // [kNumBands + 1] first entry slot of every band
  float* part_t;               // [kGsSlots][n_pad] per-entry dot products (slot-major: a unit's 8 windows are contiguous)
  int n_windows, n_pad;        // n_pad = n rounded up to a multiple of 8
  int groups;                  // window groups per band = n_pad / 8
  int n_units;                 // kNumBands * groups
  const int32_t* cta_split;    // [gridDim.x + 1] unit range of every CTA (balanced by the bands' entry counts)
  int experiment;              // timing experiments only (results become wrong): 32 = no gather work, 64 = no part_t stores, 128 = no q stores, 256 = gather reads its rows but does no arithmetic
  const uint32_t* wv_t16;      // [2 hi/lo][128 cout][64] packed fp16 pairs of w_v^T (A operand from tensor memory, ts_mode)
  int ts_mode;                 // 1 = w_v weights live in TMEM (columns 256..383), one accumulator; 0 = weights in shared memory, two accumulators
  long long* dbg;              // optional [gridDim.x][8] cycle counters (nullptr = off), see tools/ab_stages.py --wvg-cycles
  DeviceStatus* status;
};

__device__ __forceinline__ float4 ldg_weights(const float* p) {      // folded weights: keep them in L1 across units
  float4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float dot8_h(const uint4& v, const float4& a, const float4& b) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  float s = __low2float(h[0]) * a.x;
  s = fmaf(__high2float(h[0]), a.y, s);
  s = fmaf(__low2float(h[1]), a.z, s); s = fmaf(__high2float(h[1]), a.w, s);
  s = fmaf(__low2float(h[2]), b.x, s); s = fmaf(__high2float(h[2]), b.y, s);
  s = fmaf(__low2float(h[3]), b.z, s); s = fmaf(__high2float(h[3]), b.w, s);
  return s;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wv_gather_kernel(const __grid_constant__ CUtensorMap tm_band, const __grid_constant__ CUtensorMap tm_w,
                 const WvGatherParams p) {
  constexpr uint32_t kIdesc = umma_idesc_f16(128, 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;                                   // ring of kWgBufs region buffers, each [8 windows][32 rows] x 128 B
  uint8_t* s_w = smem + kWgSlab;                         // 4 resident weight stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + kWgWeights);
  uint64_t* a_full = bars;            // [kWgBufs <= 5]  per buffer
  uint64_t* a_empty = bars + 5;       // [kWgBufs <= 5]
  uint64_t* w_full = bars + 10;       // [1]
  uint64_t* acc_full = bars + 11;     // [2]
  uint64_t* acc_empty = bars + 13;    // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 15);
  // Region k of a unit (need order: 0 hi16.k0, 1 lo16.k0, 2 hi16.k1, 3 lo16.k1) is load number g = 4 * it + k of this CTA and
  // lives in buffer g % kWgBufs; every role walks the buffers in the same order, so each keeps the first buffer of the
  // current unit (b0, advanced by 4 mod kWgBufs per unit) and one phase bit per buffer that it flips after each use.

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous range of band-major unit numbers; the split points weigh a unit by its band's entry count (host side)
  const int u_begin = p.cta_split[blockIdx.x], u_end = p.cta_split[blockIdx.x + 1];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_band);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < kWgBufs; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1 + kWgWarps); }    // tcgen05.commit + the consumer warps
    mbar_init(w_full, p.ts_mode ? 4 : 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kWgWarps); }
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

  if (p.ts_mode && warp >= 4 && warp < 8) {
    // ===================================================================== weights -> tensor memory (once): thread = output
    // channel (TMEM lane), columns 256..319 = fp16(w_v^T) hi as 64 packed pairs along k, 320..383 = lo
    const int cout = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int part = 0; part < 4; ++part) {                  // (hi, lo) x (k 0..63, k 64..127): 32 columns each
      const uint32_t* src = p.wv_t16 + (static_cast<size_t>(part >> 1) * kC + cout) * 64 + (part & 1) * 32;
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + i);
        r[i] = v.x; r[i + 1] = v.y; r[i + 2] = v.z; r[i + 3] = v.w;
      }
      tmem_st_32x32(lane_addr + 256 + part * 32, r);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(w_full);
  }
  if (warp == 0 && lane == 0) {
    // ===================================================================== weights: loaded once, resident
    if (!p.ts_mode) {
      const uint64_t pol = l2_policy_evict_last();
      mbar_arrive_expect_tx(w_full, kWgWeights);
      for (int q = 0; q < kWvStages; ++q) tma_load_2d_hint(s_w + q * kBStage, &tm_w, w_full, 0, q * 128, pol);
    }
  } else if (warp == 3 && lane == 0) {
    // ===================================================================== activation producer
    const uint64_t pol = l2_policy_evict_first();
    uint32_t phases = 0;                                   // bit b: parity of buffer b's next "empty" wait is phase ^ 1
    int b0 = 0;
    for (int unit = u_begin; unit < u_end; ++unit) {
      const int band = unit / p.groups;
      const int w0 = (unit - band * p.groups) * kBandWins;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int b = b0 + k; if (b >= kWgBufs) b -= kWgBufs;
        mbar_wait(&a_empty[b], ((phases >> b) & 1) ^ 1, p.status, 500 + b);
        phases ^= 1u << b;
        mbar_arrive_expect_tx(&a_full[b], kWgRegion);
        // k: 0 hi16 channels 0-63, 1 lo16 channels 0-63, 2 hi16 channels 64-127, 3 lo16 channels 64-127 (byte offset in the row)
        const int src = (k & 1 ? kOffLo16 : kOffHi16) + (k >> 1) * 128;
        tma_load_3d_hint(s_a + b * kWgRegion, &tm_band, &a_full[b], src, band * kBandRows, w0, pol);
      }
      b0 += 4; if (b0 >= kWgBufs) b0 -= kWgBufs;
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (converged warp, elected lane issues)
    const uint64_t desc0 = umma_desc_sw128(0);
    const uint32_t a_base = smem_u32(s_a);
    const uint32_t w_base = smem_u32(s_w);
    mbar_wait(w_full, 0, p.status, 510);
    tc_fence_after();
    const bool ts = p.ts_mode != 0;
    int it = 0;
    uint32_t phases = 0;
    int b0 = 0;
    long long m_wait_full = 0, m_wait_acc = 0, tq = 0;
    for (int unit = u_begin; unit < u_end; ++unit, ++it) {
      const int as = ts ? 0 : (it & 1);                      // ts_mode: the second accumulator's columns hold the weights
      const uint32_t accphase = ts ? (it & 1) : ((it >> 1) & 1);
      const uint32_t acc = tmem_base + as * 256;
      if (p.dbg) tq = clock64();
      mbar_wait(&acc_empty[as], accphase ^ 1, p.status, 520 + as);
      if (p.dbg) m_wait_acc += clock64() - tq;
#pragma unroll
      for (int q = 0; q < kWvStages; ++q) {
        // stage q = (K-half q/2, weight hi/lo q%2); hi-weight stages multiply both the hi16 and the lo16 region
        const int kh = q >> 1;
        int bh = b0 + 2 * kh; if (bh >= kWgBufs) bh -= kWgBufs;               // hi16 region of this K-half
        int bl = b0 + 2 * kh + 1; if (bl >= kWgBufs) bl -= kWgBufs;           // lo16 region
        if ((q & 1) == 0) {
          if (p.dbg) tq = clock64();
          mbar_wait(&a_full[bh], (phases >> bh) & 1, p.status, 530 + bh);
          mbar_wait(&a_full[bl], (phases >> bl) & 1, p.status, 536 + bl);
          if (p.dbg) m_wait_full += clock64() - tq;
          phases ^= (1u << bh) | (1u << bl);
        }
        tc_fence_after();
        if (elect_one()) {
          const uint64_t wdesc = desc0 + ((w_base + q * kBStage) >> 4);                      // A: weights
          const uint64_t y0 = desc0 + ((a_base + bh * kWgRegion) >> 4);                      // B: hi16 rows
          const uint64_t y1 = desc0 + ((a_base + bl * kWgRegion) >> 4);                      // B: lo16 rows
          const bool w_lo = (q & 1) != 0;
          if (ts) {
            const uint32_t wt = tmem_base + 256 + (w_lo ? 64 : 0) + kh * 32;                 // 16 k-elements = 8 columns per MMA
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16_ts(acc, wt + kk * 8, y0 + kk * 2, kIdesc, (q == 0 && kk == 0) ? 0u : 1u);
              if (!w_lo) umma_f16_ts(acc, wt + kk * 8, y1 + kk * 2, kIdesc, 1u);
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16(acc, wdesc + kk * 2, y0 + kk * 2, kIdesc, (q == 0 && kk == 0) ? 0u : 1u);
              if (!w_lo) umma_f16(acc, wdesc + kk * 2, y1 + kk * 2, kIdesc, 1u);
            }
          }
          if (!w_lo) umma_commit(&a_empty[bl]);                  // the lo16 region is only used by the hi-weight stage
          else umma_commit(&a_empty[bh]);
          if (q == 3) umma_commit(&acc_full[as]);
        }
        __syncwarp();
      }
      b0 += 4; if (b0 >= kWgBufs) b0 -= kWgBufs;
    }
    if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 8 + 6] = m_wait_full; p.dbg[blockIdx.x * 8 + 7] = m_wait_acc; }
  } else if (warp >= 4) {
    // ===================================================================== patch gather, then accumulator epilogue
    const int gw = warp - 4;                                   // 0..19
    const int wq = warp & 3;                                   // TMEM lane quarter = channels 32*wq .. 32*wq+31
    const int grp = gw >> 2;                                   // epilogue: windows grp and grp+5 of the unit
    const int ch = wq * 32 + lane;
    const float oscale = p.out_scale;
    // gather lane roles: quarter-warp (lane >> 3): 0 = hi row of the even window, 1 = lo row of the even window,
    // 2 / 3 = the same for the odd window of the pair; chunk j = lane & 7 = channels 8j .. 8j+7 of the current K-half
    const int jch = lane & 7;
    const int plane = (lane >> 3) & 1;                         // 0 hi16, 1 lo16
    const int wodd = lane >> 4;                                // window 2i + wodd of pair i
    const bool up8 = lane & 8, up4 = lane & 4;
    const int wi_out = ((lane >> 2) & 1) * 4 + ((lane >> 3) & 1) * 2 + (lane >> 4);      // window this lane ends up holding
    const uint32_t slab = smem_u32(s_a);
    // 4 partial sums (window pairs 0..3) of one entry -> the total of window wi_out in the lanes with (lane & 3) == 0:
    // transposing butterfly over the 16 lanes that share a window, fixed order (deterministic)
    auto reduce4 = [&](const float (&a)[4]) -> float {
      float b0, b1, c;
      { const float g = up8 ? a[0] : a[1], k = up8 ? a[1] : a[0]; b0 = k + __shfl_xor_sync(0xffffffffu, g, 8); }
      { const float g = up8 ? a[2] : a[3], k = up8 ? a[3] : a[2]; b1 = k + __shfl_xor_sync(0xffffffffu, g, 8); }
      { const float g = up4 ? b0 : b1, k = up4 ? b1 : b0; c = k + __shfl_xor_sync(0xffffffffu, g, 4); }
      c += __shfl_xor_sync(0xffffffffu, c, 2);
      c += __shfl_xor_sync(0xffffffffu, c, 1);
      return c;
    };
    int it = 0;
    uint32_t phases = 0;
    int b0 = 0;
    long long c_wait_full = 0, c_gather = 0, c_wait_acc = 0, c_epi = 0, tq = 0;
    const long long t_begin = clock64();
    int cur_band = -1, e_begin = 0, cnt = 0;
    int rr[kWgWarpCap] = {0, 0, 0};                         // rows (inside the band) of this warp's entries, fast path
    bool fast = true;
    for (int unit = u_begin; unit < u_end; ++unit, ++it) {
      const int band = unit / p.groups;
      const int w0 = (unit - band * p.groups) * kBandWins;
      if (band != cur_band) {                                  // this warp's contiguous run of the band's entries
        cur_band = band;
        const int s0 = p.band_start[band], b1 = p.band_start[band + 1];
        const int per = (b1 - s0 + kWgWarps - 1) / kWgWarps;
        e_begin = min(b1, s0 + gw * per);
        cnt = min(b1, e_begin + per) - e_begin;
        if (p.experiment & 32) cnt = 0;
        fast = per <= kWgWarpCap;
#pragma unroll
        for (int i = 0; i < kWgWarpCap; ++i) rr[i] = (fast && i < cnt) ? p.ent_pos[e_begin + i] - band * kBandRows : 0;
      }
      // ---------------- gather: this warp's entries x the unit's 8 windows, one K-half per pass.  Few entries per warp
      // (20 warps share the band's ~45) keep the time a K-half's regions are held short: the slab is single-buffered, so
      // the next unit's loads start only when every consumer has released a region.
      float c0[kWgWarpCap];                                    // pass-0 halves of the fast path (lanes with (lane & 3) == 0)
#pragma unroll 1
      for (int kh = 0; kh < 2; ++kh) {
        int bh = b0 + 2 * kh; if (bh >= kWgBufs) bh -= kWgBufs;               // buffer of this K-half's hi16 region
        int bl = b0 + 2 * kh + 1; if (bl >= kWgBufs) bl -= kWgBufs;           // ... and of its lo16 region
        const uint32_t ph_h = (phases >> bh) & 1, ph_l = (phases >> bl) & 1;
        phases ^= (1u << bh) | (1u << bl);
        const uint32_t reg_base = slab + (plane ? bl : bh) * kWgRegion + wodd * (kBandRows * 128) + (jch << 4);
        if (fast) {
          if (p.dbg) tq = clock64();
          mbar_wait(&a_full[bh], ph_h, p.status, 550 + bh);          // hi16 K-half kh
          mbar_wait(&a_full[bl], ph_l, p.status, 556 + bl);          // lo16 K-half kh
          if (p.dbg) { const long long t = clock64(); c_wait_full += t - tq; tq = t; }
          // Entries are sorted by position and 46 % of them share their row with a neighbour (8,400 entries hit ~4,500
          // distinct positions): a row that the previous entry of this warp already pulled out of the slab is not read again.
          // (Timing experiment 256 -- rows read, no arithmetic -- costs as much as the full gather: the LDS traffic next to the
          // MMAs' operand reads and the TMA fills is what the gather costs, not its FMAs.)
          uint4 v[4];
#pragma unroll
          for (int i = 0; i < kWgWarpCap; ++i)
            if (i < cnt) {
              if (i == 0 || rr[i] != rr[i - 1]) {
                const uint32_t ad = (reg_base + rr[i] * 128) ^ ((rr[i] & 7) << 4);   // 128-byte swizzle: chunk j -> j ^ (row & 7)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = lds128(ad + j * (2 * kBandRows * 128));
              }
              // Folded weights of this entry and K-half through L1 / L2 (evict_last).  Staging them in shared memory instead
              // (30 KB, private slots per warp) was measured and is slower (1.11 / 1.30 ms): the kernel's limit is the
              // shared-memory port (UMMA operand reads 288 KB + TMA fills 131 KB + these LDS per unit ~ 95 of 128 B/cycle;
              // cycle counters: ~650 cycles per (entry, K-half) in a consumer warp with everything in shared memory), so
              // weight reads are better off on the L1 / L2 path.
              const float* wp = p.ent_w + static_cast<size_t>(e_begin + i) * kC + kh * 64 + jch * 8;
              const float4 wa = ldg_weights(wp), wb = ldg_weights(wp + 4);
              float a[4];
              if (p.experiment & 256) {                        // timing experiment: rows are read, no arithmetic
                asm volatile("" :: "r"(v[0].x | v[1].x | v[2].x | v[3].x));
                c0[i] = 0.f;
                continue;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) a[j] = dot8_h(v[j], wa, wb);
              const float c = reduce4(a);
              if (kh == 0) c0[i] = c;
              else if ((lane & 3) == 0 && !(p.experiment & 64)) p.part_t[static_cast<size_t>(e_begin + i) * p.n_pad + w0 + wi_out] = c0[i] + c;
            }
        } else {
          // generic path (band with more than 64 entries): pass-0 halves parked in part_t itself
          mbar_wait(&a_full[bh], ph_h, p.status, 550 + bh);
          mbar_wait(&a_full[bl], ph_l, p.status, 556 + bl);
#pragma unroll 1
          for (int i = 0; i < cnt; ++i) {
            const int e = e_begin + i;
            const int r = p.ent_pos[e] - band * kBandRows;
            const float* wp = p.ent_w + static_cast<size_t>(e) * kC + kh * 64 + jch * 8;
            const float4 wa = ldg_weights(wp), wb = ldg_weights(wp + 4);
            const uint32_t ad = (reg_base + r * 128) ^ ((r & 7) << 4);
            float a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = dot8_h(lds128(ad + j * (2 * kBandRows * 128)), wa, wb);
            const float c = reduce4(a);
            if ((lane & 3) == 0) {
              float* gp = p.part_t + static_cast<size_t>(e) * p.n_pad + w0 + wi_out;
              *gp = kh == 0 ? c : *gp + c;
            }
          }
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(&a_empty[bh]); mbar_arrive(&a_empty[bl]); }           // this warp is done with the K-half's regions
        if (p.dbg && fast) c_gather += clock64() - tq;
      }
      b0 += 4; if (b0 >= kWgBufs) b0 -= kWgBufs;
      // ---------------- epilogue: q[w][band*4 + g][ch] = max over the 8 positions of pool group g
      const int as = p.ts_mode ? 0 : (it & 1);
      const uint32_t accphase = p.ts_mode ? (it & 1) : ((it >> 1) & 1);
      if (p.dbg) tq = clock64();
      mbar_wait(&acc_full[as], accphase, p.status, 540 + as);
      if (p.dbg) { const long long t = clock64(); c_wait_acc += t - tq; tq = t; }
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * 256;
#pragma unroll 1
      for (int c32 = grp; c32 < kBandWins; c32 += kWgWarps / 4) {
        const int w = w0 + c32;
        if (w >= p.n_windows) break;                           // uniform: the rest of the unit is past the batch end
        uint32_t r[32];
        tmem_ld_32x32(lane_addr + c32 * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float m = __uint_as_float(r[8 * g]);
#pragma unroll
          for (int k = 1; k < 8; ++k) m = fmaxf(m, __uint_as_float(r[8 * g + k]));
          const int gg = band * (kBandRows / kPool) + g;
          if (gg < kPooled && !(p.experiment & 128)) p.q_out[(static_cast<size_t>(w) * kPooled + gg) * kC + ch] = m * oscale;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (p.dbg) c_epi += clock64() - tq;
    }
    if (p.dbg && warp == 4 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 8;
      d[0] = clock64() - t_begin; d[1] = c_wait_full; d[2] = c_gather; d[3] = c_wait_acc; d[4] = c_epi; d[5] = it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// mpi[w][p] = ((part_t[s0][w] + part_t[s1][w]) + part_t[s2][w]) + part_t[s3][w] + bias[p],  s_k = slot_of[4p + k];
// also the value's two TF32 halves for the tensor-core logits GEMM (logits_tc.cuh).  32 patches x 32 windows per CTA;
// reads are contiguous along the window axis, the transposed tile makes the writes contiguous along the patch axis.
__global__ void __launch_bounds__(256)
patch_finish_t_kernel(const float* __restrict__ part_t, const int32_t* __restrict__ slot_of, const float* __restrict__ w_bias,
                      float* __restrict__ mpi, float* __restrict__ mpi_hi, float* __restrict__ mpi_lo, int n_windows, int n_pad) {
  __shared__ float tile[32][33];
  const int p0 = blockIdx.x * 32, w0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 8 rows of 32 threads
  for (int pp = ty; pp < 32; pp += 8) {
    const int pch = p0 + pp, w = w0 + tx;
    float v = 0.f;
    if (pch < kPatches && w < n_windows) {
      const int4 s4 = *reinterpret_cast<const int4*>(slot_of + pch * 4);
      v = (((part_t[static_cast<size_t>(s4.x) * n_pad + w] + part_t[static_cast<size_t>(s4.y) * n_pad + w]) +
            part_t[static_cast<size_t>(s4.z) * n_pad + w]) + part_t[static_cast<size_t>(s4.w) * n_pad + w]) + w_bias[pch];
    }
    tile[pp][tx] = v;
  }
  __syncthreads();
  for (int ww = ty; ww < 32; ww += 8) {
    const int pch = p0 + tx, w = w0 + ww;
    if (pch < kPatches && w < n_windows) {
      const float v = tile[tx][ww];
      const size_t o = static_cast<size_t>(w) * kPatches + pch;
      mpi[o] = v;
      const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      mpi_hi[o] = hi;
      mpi_lo[o] = __uint_as_float(__float_as_uint(v - hi) & 0xffffe000u);
    }
  }
}

}  // namespace gnm
