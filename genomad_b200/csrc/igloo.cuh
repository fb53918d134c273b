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
// Patch gather.  Grid = (148 patch groups of <=15 patches) x (window chunks), sized by the host to
// exactly one resident wave (CTAs of 128 threads, several per SM).  A CTA stages its group's folded weights
// (<= 16 x 4 x 128 fp32 = 32 KB) in shared memory once and reuses them for every window of its chunk.
// Each warp owns 4 patches = 16 activation rows per window: it issues all 16 row loads back to back
// (one 512-byte row = ONE 16-byte load per lane: lanes 0-15 fetch the fp16 "hi" half-row, lanes 16-31
// the "lo" half-row) and keeps the NEXT window's 16 rows in flight while it multiplies the current ones by the staged weights (lanes l and l+16 use the same 8 weights, so
// hi*w and lo*w are summed by the final warp-shuffle reduction) and writes 4 scores.
// All groups walk the windows in the same order, so the ~1.9x re-use of popular rows hits in L2.
// HBM-bound: 8400 rows x 512 B = 4.3 MB of activation rows per window per IGLOO kernel (2.4 MB distinct).
// ------------------------------------------------------------------------------------------
constexpr int kGatherGroups = 148;
constexpr int kGatherPB = 16;                                    // patch slots per CTA (4 warps x 4)
constexpr int kGatherPPG = (kPatches + kGatherGroups - 1) / kGatherGroups;   // 15 patches per group
constexpr int kGatherThreads = 128;
constexpr int kGatherSmem = kGatherPB * kPatchLen * kC * 4 + kGatherPB * kPatchLen * 4 + kGatherPB * 4;
static_assert(kGatherPPG <= kGatherPB, "patch group does not fit the CTA");

__global__ void __launch_bounds__(kGatherThreads, 2)
patch_gather_kernel(const __half* __restrict__ y,         // [n][5997][256]
                    const float* __restrict__ wf,         // [2100][4][128] folded
                    const int32_t* __restrict__ patches,  // [2100][4]
                    const float* __restrict__ w_bias,     // [2100]
                    float* __restrict__ mpi,              // [n][2100]
                    int n_windows, int windows_per_cta) {
  extern __shared__ __align__(16) uint8_t s_g[];
  float* s_wf = reinterpret_cast<float*>(s_g);                                    // [16][4][128]
  int* s_idx = reinterpret_cast<int*>(s_wf + kGatherPB * kPatchLen * kC);          // [16][4]
  float* s_bias = reinterpret_cast<float*>(s_idx + kGatherPB * kPatchLen);         // [16]
  const int p0 = blockIdx.x * kGatherPPG;
  const int np = max(0, min(kGatherPPG, kPatches - p0));
  for (int i = threadIdx.x; i < kGatherPB * kPatchLen * kC / 4; i += kGatherThreads) {
    const int pp = i / (kPatchLen * kC / 4);
    reinterpret_cast<float4*>(s_wf)[i] = pp < np
        ? reinterpret_cast<const float4*>(wf + static_cast<size_t>(p0) * kPatchLen * kC)[i]
        : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < kGatherPB * kPatchLen; i += kGatherThreads)
    s_idx[i] = (i / kPatchLen) < np ? patches[p0 * kPatchLen + i] : 0;
  for (int i = threadIdx.x; i < kGatherPB; i += kGatherThreads) s_bias[i] = i < np ? w_bias[p0 + i] : 0.f;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;               // half 0: "hi" plane, 1: "lo" plane
  int rowoff[16];                                            // byte offset of this lane's 16 B inside each of the 16 rows
#pragma unroll
  for (int j = 0; j < 16; ++j)
    rowoff[j] = s_idx[warp * 16 + j] * (kRowHalfs * 2) + half * (kC * 2) + l16 * 16;
  const float* wl = s_wf + warp * 16 * kC + l16 * 8;          // this lane's 8 weights of row j: wl[j*128 .. +7]
  const int w_begin = blockIdx.y * windows_per_cta;
  const int w_end = min(n_windows, w_begin + windows_per_cta);
  const uint8_t* ybase = reinterpret_cast<const uint8_t*>(y);
  constexpr size_t kWinBytes = static_cast<size_t>(kTok) * kRowHalfs * 2;
  uint4 nv[16];                                              // rows of the NEXT window, in flight while this one is reduced
  if (w_begin < w_end) {
#pragma unroll
    for (int j = 0; j < 16; ++j) nv[j] = __ldg(reinterpret_cast<const uint4*>(ybase + w_begin * kWinBytes + rowoff[j]));
  }
  for (int w = w_begin; w < w_end; ++w) {
    uint4 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = nv[j];
    if (w + 1 < w_end) {
#pragma unroll
      for (int j = 0; j < 16; ++j) nv[j] = __ldg(reinterpret_cast<const uint4*>(ybase + (w + 1) * kWinBytes + rowoff[j]));
    }
    float acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = i * 4 + k;
        const float4 w0 = *reinterpret_cast<const float4*>(wl + j * kC);
        const float4 w1 = *reinterpret_cast<const float4*>(wl + j * kC + 4);
        const __half2* h = reinterpret_cast<const __half2*>(&v[j]);
        a = fmaf(__low2float(h[0]), w0.x, a); a = fmaf(__high2float(h[0]), w0.y, a);
        a = fmaf(__low2float(h[1]), w0.z, a); a = fmaf(__high2float(h[1]), w0.w, a);
        a = fmaf(__low2float(h[2]), w1.x, a); a = fmaf(__high2float(h[2]), w1.y, a);
        a = fmaf(__low2float(h[3]), w1.z, a); a = fmaf(__high2float(h[3]), w1.w, a);
      }
      acc[i] = a;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
    if (lane < 4) {
      const int pl = warp * 4 + lane;
      if (pl < np) {
        const float r = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        mpi[static_cast<size_t>(w) * kPatches + p0 + pl] = r + s_bias[pl];
      }
    }
  }
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
