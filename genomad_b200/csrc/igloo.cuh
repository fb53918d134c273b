// K4: IGLOO patch gather, K6: attention softmax + weighted sum of the pooled projections.
//
// Reference semantics (genomad/neural_network/igloo.py:190-217):
//   M   = gather_nd(transpose(y), patches)                 -> y[b, P[p,k], c]
//   mpi = sum_k sum_c y[b,P[p,k],c] * w_mult[p,k,c] * w_summer[128k+c] + w_bias[p]      (lines 192-206)
//   alpha = softmax(mpi @ w_qk)   (lines 211-212; the GEMM is K5 in dense.cuh)
//   out = alpha @ maxpool8(y @ w_v)                        (lines 213-214)
// The patch weights are folded once at load time: Wf[p,k,c] = w_mult[p,k,c] * w_summer[128k+c]
// (fp32 product, so where the reference's two-step product underflows to 0, so does this one).
#pragma once
#include "common.cuh"

namespace gnm {

// ------------------------------------------------------------------------------------------
// Patch gather, position-ordered streaming form.
//
// The 8400 (patch, slot) entries of an IGLOO kernel touch only ~4500 distinct positions, so reading one
// row per entry moves every popular row ~1.9x through L2 -> SM; the first version (one CTA per 32-patch
// group, random rows) was bound by exactly that (5 TB/s L2->SM for 2.4 GB of DRAM reads,
// profiles/r01_small_kernels_ncu.md).  Here the entries are sorted by position on the host and dealt to
// 444 CTAs x 4 warps x 5 entries (= 3 CTAs per SM, one resident wave, no shared memory): a warp's rows
// are neighbours in memory, repeats of a row are served by L1, and the whole grid sweeps each window
// front to back.  Every lane keeps its 5 x 8 folded weights in registers and a 4-window deep ring of
// row fragments in flight (one 512-byte row = ONE 16-byte load per lane: lanes 0-15 the fp16 "hi"
// half-row, lanes 16-31 the "lo" half-row; lanes l and l+16 use the same weights, so hi*w and lo*w meet
// in the warp-shuffle reduction).  Each entry has exactly one writer (part[w][slot]); patch_finish_kernel
// then adds the four slots of a patch in fixed order k = 0..3 plus the bias, so results are deterministic.
// Measured (profiles/r01_small_kernels_ncu.md): 0.74 ms per 1024 windows = 3.2 TB/s of DRAM reads, 40 % of peak.
// A variant that staged the rows with 512-byte cp.async.bulk copies (6 windows deep, 184 KB in flight per SM)
// was slower (0.88 ms), i.e. the limit is not outstanding-load capacity but DRAM efficiency on 512-byte pieces
// requested by many CTAs at once; reading each CTA's position range as one contiguous block is the next step.
// HBM-bound: 2.4 MB of distinct activation rows per window per IGLOO kernel (algorithmic 8400 x 512 B = 4.3 MB).
// ------------------------------------------------------------------------------------------
constexpr int kGsGroups = 444;                                   // CTAs (3 per SM on 148 SMs)
constexpr int kGsPerWarp = 5;                                    // entries per warp
constexpr int kGsThreads = 128;
constexpr int kGsPerCta = 4 * kGsPerWarp;                        // 20
constexpr int kGsSlots = kGsGroups * kGsPerCta;                  // 8880 >= 8400 (padded with zero-weight entries)
constexpr int kGsDepth = 4;                                      // windows in flight per warp
static_assert(kGsSlots >= kPatches * kPatchLen, "not enough entry slots");

__global__ void __launch_bounds__(kGsThreads, 3)
patch_stream_kernel(const uint8_t* __restrict__ y,         // [n][5997][768 B]; reads the hi16 / lo16 planes (scaled by 32)
                    const int32_t* __restrict__ ent_pos,   // [8880] position of each slot (0 for padding)
                    const float* __restrict__ ent_w,       // [8880][128] folded weights / 32 of each slot (0 for padding)
                    float* __restrict__ part,              // [n][8880]
                    int n_windows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int e0 = blockIdx.x * kGsPerCta + warp * kGsPerWarp;
  int rowoff[kGsPerWarp];
  float wt[kGsPerWarp][8];
#pragma unroll
  for (int i = 0; i < kGsPerWarp; ++i) {
    rowoff[i] = ent_pos[e0 + i] * kRowBytes + (half ? kOffLo16 : kOffHi16) + l16 * 16;
    const float4 a = *reinterpret_cast<const float4*>(ent_w + static_cast<size_t>(e0 + i) * kC + l16 * 8);
    const float4 b = *reinterpret_cast<const float4*>(ent_w + static_cast<size_t>(e0 + i) * kC + l16 * 8 + 4);
    wt[i][0] = a.x; wt[i][1] = a.y; wt[i][2] = a.z; wt[i][3] = a.w;
    wt[i][4] = b.x; wt[i][5] = b.y; wt[i][6] = b.z; wt[i][7] = b.w;
  }
  const uint8_t* ybase = y;
  constexpr size_t kWinBytes = static_cast<size_t>(kTok) * kRowBytes;
  uint4 ring[kGsDepth][kGsPerWarp];
#pragma unroll
  for (int d = 0; d < kGsDepth; ++d)
    if (d < n_windows) {
#pragma unroll
      for (int i = 0; i < kGsPerWarp; ++i)
        ring[d][i] = __ldg(reinterpret_cast<const uint4*>(ybase + d * kWinBytes + rowoff[i]));
    }
  for (int w0 = 0; w0 < n_windows; w0 += kGsDepth) {
#pragma unroll
    for (int d = 0; d < kGsDepth; ++d) {
      const int w = w0 + d;
      if (w < n_windows) {
        float acc[kGsPerWarp];
#pragma unroll
        for (int i = 0; i < kGsPerWarp; ++i) {
          const __half2* h = reinterpret_cast<const __half2*>(&ring[d][i]);
          float a = __low2float(h[0]) * wt[i][0];
          a = fmaf(__high2float(h[0]), wt[i][1], a);
          a = fmaf(__low2float(h[1]), wt[i][2], a); a = fmaf(__high2float(h[1]), wt[i][3], a);
          a = fmaf(__low2float(h[2]), wt[i][4], a); a = fmaf(__high2float(h[2]), wt[i][5], a);
          a = fmaf(__low2float(h[3]), wt[i][6], a); a = fmaf(__high2float(h[3]), wt[i][7], a);
          acc[i] = a;
        }
        if (w + kGsDepth < n_windows) {
#pragma unroll
          for (int i = 0; i < kGsPerWarp; ++i)
            ring[d][i] = __ldg(reinterpret_cast<const uint4*>(ybase + (w + kGsDepth) * kWinBytes + rowoff[i]));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
          for (int i = 0; i < kGsPerWarp; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
        if (lane < kGsPerWarp) {
          const float r = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : lane == 3 ? acc[3] : acc[4];
          part[static_cast<size_t>(w) * kGsSlots + e0 + lane] = r;
        }
      }
    }
  }
}

// mpi[w][p] = ((part[s0] + part[s1]) + part[s2]) + part[s3] + bias[p],  s_k = slot_of[4p + k]
__global__ void __launch_bounds__(256)
patch_finish_kernel(const float* __restrict__ part, const int32_t* __restrict__ slot_of, const float* __restrict__ w_bias,
                    float* __restrict__ mpi, int n_windows) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = blockIdx.y;
  if (p >= kPatches) return;
  const float* pw = part + static_cast<size_t>(w) * kGsSlots;
  const int4 s4 = *reinterpret_cast<const int4*>(slot_of + p * 4);
  mpi[static_cast<size_t>(w) * kPatches + p] = (((pw[s4.x] + pw[s4.y]) + pw[s4.z]) + pw[s4.w]) + w_bias[p];
}

// ------------------------------------------------------------------------------------------
// Attention tail: one CTA (128 threads = channels) per (window, igloo).  Row softmax over the
// 749 logits (max-subtracted, expf, fp32), then out[c] = sum_g alpha[g] * q[g][c].
// ------------------------------------------------------------------------------------------
constexpr int kLogitsLd = 752;     // logits row stride (749 rounded up to a multiple of 4)

__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ logits,   // [n][752]
                 const float* __restrict__ q,        // [n][749][128]
                 float* __restrict__ h0,             // [n][256]
                 int col_offset) {                   // 0 for IGLOO#0, 128 for IGLOO#1
  __shared__ float s_alpha[kPooled];
  __shared__ float s_red[4];
  const int w = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* lrow = logits + static_cast<size_t>(w) * kLogitsLd;
  float m = -INFINITY;
  for (int g = tid; g < kPooled; g += 128) { const float v = lrow[g]; s_alpha[g] = v; m = fmaxf(m, v); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int g = tid; g < kPooled; g += 128) { const float e = expf(s_alpha[g] - m); s_alpha[g] = e; sum += e; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
  const float inv = 1.f / sum;
  const float* qw = q + static_cast<size_t>(w) * kPooled * kC + tid;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int g = 0;
  for (; g + 4 <= kPooled; g += 4) {
    a0 = fmaf(s_alpha[g] * inv, qw[static_cast<size_t>(g) * kC], a0);
    a1 = fmaf(s_alpha[g + 1] * inv, qw[static_cast<size_t>(g + 1) * kC], a1);
    a2 = fmaf(s_alpha[g + 2] * inv, qw[static_cast<size_t>(g + 2) * kC], a2);
    a3 = fmaf(s_alpha[g + 3] * inv, qw[static_cast<size_t>(g + 3) * kC], a3);
  }
  for (; g < kPooled; ++g) a0 = fmaf(s_alpha[g] * inv, qw[static_cast<size_t>(g) * kC], a0);
  h0[static_cast<size_t>(w) * 256 + col_offset + tid] = (a0 + a1) + (a2 + a3);
}

}  // namespace gnm
