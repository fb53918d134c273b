// K0: ASCII -> 4-mer tokens, and K0+K1 fused: ASCII/tokens -> first conv layer output.
//
// Reference semantics:
//   tokenize_dna            genomad/sequence.py:170-193  (closed form: tok = 0 if any of the 4 bytes
//                           is not one of 'A','C','G','T' (65,67,71,84), else 1 + base-4 value)
//   one-hot + Conv1D#1      genomad/neural_network/model.py:11, igloo.py:45-48
//                           y1[t] = lrelu(b + sum_{j=0..5, t-5+j>=0} W1[j][tok[t-5+j]][:])
//                           (causal zero padding adds nothing -- it is NOT token 0)
#pragma once
#include "common.cuh"

namespace gnm {

// 2-bit code of an upper-case nucleotide byte, or 4 for anything else
__device__ __forceinline__ uint32_t base_code(uint8_t b) {
  return b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 4u;
}
__device__ __forceinline__ uint16_t kmer_token(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  const uint32_t bad = (c0 | c1 | c2 | c3) & 4u;
  const uint32_t v = 1u + 64u * c0 + 16u * c1 + 4u * c2 + c3;
  return bad ? uint16_t(0) : uint16_t(v);
}

// ------------------------------------------------------------------------------------------
// K0 (stand-alone): one CTA per (window, 2048-token segment).  Bytes are staged through shared
// memory with 16-byte coalesced loads; every thread then emits 8 consecutive tokens as one
// 16-byte store.  Pure byte/integer work, HBM-bound: 6000 B in + 11994 B out per window.
// ------------------------------------------------------------------------------------------
constexpr int kEncSeg = 2048;                         // tokens per CTA
constexpr int kEncThreads = kEncSeg / 8;              // 256

__global__ void __launch_bounds__(kEncThreads)
encode_tokens_kernel(const uint8_t* __restrict__ ascii, uint16_t* __restrict__ tokens, int n_windows) {
  __shared__ __align__(16) uint8_t s_b[kEncSeg + 16];
  const int w = blockIdx.y;
  const int t0 = blockIdx.x * kEncSeg;
  const uint8_t* src = ascii + static_cast<size_t>(w) * kWindow;
  // window rows are 6000 B apart: 16-byte aligned (6000 = 375 * 16) as long as the base is.
  for (int i = threadIdx.x; i < (kEncSeg + 16) / 16; i += blockDim.x) {
    const int off = t0 + i * 16;
    uint4 v = make_uint4(0x4E4E4E4Eu, 0x4E4E4E4Eu, 0x4E4E4E4Eu, 0x4E4E4E4Eu);   // 'N'
    if (off + 16 <= kWindow) {
      v = *reinterpret_cast<const uint4*>(src + off);
    } else if (off < kWindow) {
      uint8_t tmp[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) tmp[k] = (off + k < kWindow) ? src[off + k] : uint8_t('N');
      v = *reinterpret_cast<uint4*>(tmp);
    }
    *reinterpret_cast<uint4*>(s_b + i * 16) = v;
  }
  __syncthreads();
  const int lt = threadIdx.x * 8;            // first token (within the segment) of this thread
  uint32_t c[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) c[k] = base_code(s_b[lt + k]);
  uint16_t out[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) out[k] = kmer_token(c[k], c[k + 1], c[k + 2], c[k + 3]);
  uint16_t* dst = tokens + static_cast<size_t>(w) * kTok;
  const int t = t0 + lt;
  // token rows are 5997*2 B apart -> not 16-byte aligned in general: scalar stores
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (t + k < kTok) dst[t + k] = out[k];
}

// ------------------------------------------------------------------------------------------
// K0+K1(+K4 of IGLOO#0) fused: tokens -> layer-1 activation rows -> patch-gather partials of the FIRST IGLOO kernel.
//
//   y1[t] = lrelu( (A + B) + bias ),   A = (W1[0][k0] + W1[1][k1]) + W1[2][k2],  B = (W1[3][k3] + W1[4][k4]) + W1[5][k5]
//
// with k_j = tok[t-5+j] and padded taps contributing exactly 0.  Tokens t-5, t-4, t-3 are the 4-mers of 6 consecutive
// bases, so A has only 4^6 = 4096 possible values when those bases are all ACGT, and likewise B: two 2 MB "triple"
// tables (built on the host with the same fp32 operation order, so a table hit is bit-identical to the three-row sum)
// replace six 512-byte row reads by two.  Positions next to the window start, or touching a non-ACGT base, fall back to
// the single-row table.  (The first version, six reads per position, was L1-bandwidth bound: l1tex 97 % busy.)
//
// Work split: a CTA owns a 64-position segment for a whole chunk of windows (grid = 94 segments x window chunks, one
// resident wave), so the folded patch weights of the entries that fall into its segment (they are contiguous in the
// position-sorted slot order, ~90 x 512 B) sit in shared memory for the lifetime of the CTA.  A warp computes one
// activation row in registers (4 channels per lane), writes its four planes (768 B, layout in common.cuh) and, while the
// row is still in registers, takes its dot product with every patch entry at that position -- so the first IGLOO
// kernel's gather never re-reads y1 from HBM (2.4 GB per 1024 windows).  One writer per entry slot; patch_finish_kernel
// adds the four slots of a patch in fixed order.
// ------------------------------------------------------------------------------------------
constexpr int kL1Seg = 64;                            // positions per CTA
constexpr int kL1Segs = (kTok + kL1Seg - 1) / kL1Seg; // 94
constexpr int kL1Threads = 256;
constexpr int kL1Win = 4;                             // windows tokenised per block-wide barrier
constexpr int kL1MaxEnt = 112;                        // patch entries staged in shared memory (56 KB); the rest stream from L2
constexpr int kTriple = 4096;
constexpr int kL1Smem = kL1MaxEnt * kC * 4;

template <bool kFromAscii>
__global__ void __launch_bounds__(kL1Threads)
layer1_gather_kernel(const uint8_t* __restrict__ ascii, const uint16_t* __restrict__ tokens_in,
                     const float* __restrict__ table,    // [6][257][128]
                     const float* __restrict__ triple,   // [2][4096][128]: A-table, B-table
                     const float* __restrict__ bias,     // [128]
                     const float* __restrict__ ent_w,    // [slots][128] folded patch weights / 32, position-sorted slots
                     const int32_t* __restrict__ pos_slot,   // [5998] first slot of each position (CSR over positions)
                     uint8_t* __restrict__ y_out,        // [n][5997][768 B] activation rows (hi16 | lo16 | lo8 | hi8)
                     float* __restrict__ part,           // [n][part_ld] per-entry partial dot products of IGLOO#0
                     int part_ld, int n_windows, int windows_per_cta) {
  extern __shared__ __align__(16) float s_wf[];             // [<= kL1MaxEnt][128]
  __shared__ int16_t s_tok[kL1Win][kL1Seg + 8];             // token at position p0 - 5 + i, or -1 (causal pad)
  __shared__ int s_ptr[kL1Seg + 1];                         // local entry offsets of the segment's positions
  const int p0 = blockIdx.x * kL1Seg;
  const int npos = min(kL1Seg, kTok - p0);
  const int s_lo = pos_slot[p0], s_hi = pos_slot[p0 + npos];
  for (int i = threadIdx.x; i <= kL1Seg; i += kL1Threads) s_ptr[i] = pos_slot[min(p0 + i, p0 + npos)] - s_lo;
  const int n_stage = min(s_hi - s_lo, kL1MaxEnt);
  for (int i = threadIdx.x; i < n_stage * (kC / 4); i += kL1Threads)
    reinterpret_cast<float4*>(s_wf)[i] = reinterpret_cast<const float4*>(ent_w + static_cast<size_t>(s_lo) * kC)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 b4 = reinterpret_cast<const float4*>(bias)[lane];
  const float4* tab4 = reinterpret_cast<const float4*>(table);
  const float4* tri4 = reinterpret_cast<const float4*>(triple);
  auto row = [&](int j, int tk) -> float4 {
    return tk >= 0 ? __ldg(tab4 + (static_cast<size_t>(j) * kVocab + tk) * (kC / 4) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto add4 = [](float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); };
  const int w_begin = blockIdx.y * windows_per_cta;
  const int w_end = min(n_windows, w_begin + windows_per_cta);

  for (int w0 = w_begin; w0 < w_end; w0 += kL1Win) {
    const int nw = min(kL1Win, w_end - w0);
    __syncthreads();                                        // previous group's tokens are no longer read (also covers the staging above)
    for (int idx = threadIdx.x; idx < nw * (kL1Seg + 5); idx += kL1Threads) {
      const int k = idx / (kL1Seg + 5), i = idx - k * (kL1Seg + 5);
      const int p = p0 - 5 + i;
      int16_t tk = -1;
      if (p >= 0 && p < kTok) {
        if (kFromAscii) {
          const uint8_t* src = ascii + static_cast<size_t>(w0 + k) * kWindow + p;
          tk = static_cast<int16_t>(kmer_token(base_code(src[0]), base_code(src[1]), base_code(src[2]), base_code(src[3])));
        } else {
          tk = static_cast<int16_t>(tokens_in[static_cast<size_t>(w0 + k) * kTok + p]);
        }
      }
      s_tok[k][i] = tk;
    }
    __syncthreads();
    for (int item = warp; item < nw * npos; item += kL1Threads / 32) {
      const int k = item / npos, i = item - k * npos;
      const int w = w0 + k, t = p0 + i;
      int tk[kTaps];
#pragma unroll
      for (int j = 0; j < kTaps; ++j) tk[j] = s_tok[k][i + j];        // token at position t - 5 + j
      float4 A, B;
      if (tk[0] > 0 && tk[2] > 0) {                                   // bases t-5 .. t all ACGT (implies tk[1] > 0)
        const int code = ((tk[0] - 1) << 4) | ((tk[2] - 1) & 15);
        A = __ldg(tri4 + static_cast<size_t>(code) * (kC / 4) + lane);
      } else {
        A = add4(add4(row(0, tk[0]), row(1, tk[1])), row(2, tk[2]));
      }
      if (tk[3] > 0 && tk[5] > 0) {                                   // bases t-2 .. t+3 all ACGT (implies tk[4] > 0)
        const int code = ((tk[3] - 1) << 4) | ((tk[5] - 1) & 15);
        B = __ldg(tri4 + (static_cast<size_t>(kTriple) + code) * (kC / 4) + lane);
      } else {
        B = add4(add4(row(3, tk[3]), row(4, tk[4])), row(5, tk[5]));
      }
      float4 a = add4(A, B);
      // Y = 32 * y1; planes: hi16, lo16 (w_v, gather), lo8 / hi8 (conv2 correction passes)
      a.x = kActScale * lrelu(a.x + b4.x); a.y = kActScale * lrelu(a.y + b4.y);
      a.z = kActScale * lrelu(a.z + b4.z); a.w = kActScale * lrelu(a.w + b4.w);
      __half2 h01, h23, l01, l23;
      split2_f16(a.x, a.y, h01, l01);
      split2_f16(a.z, a.w, h23, l23);
      const float2 fa = __half22float2(h01), fb = __half22float2(h23);
      uint8_t* rowp = y_out + (static_cast<size_t>(w) * kTok + t) * kRowBytes;
      *reinterpret_cast<uint2*>(rowp + kOffHi16 + lane * 8) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
      *reinterpret_cast<uint2*>(rowp + kOffLo16 + lane * 8) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      *reinterpret_cast<uint32_t*>(rowp + kOffLo8 + lane * 4) =
          static_cast<uint32_t>(pack_e4m3x2((a.x - fa.x) * kLo8Scale, (a.y - fa.y) * kLo8Scale)) |
          (static_cast<uint32_t>(pack_e4m3x2((a.z - fb.x) * kLo8Scale, (a.w - fb.y) * kLo8Scale)) << 16);
      *reinterpret_cast<uint32_t*>(rowp + kOffHi8 + lane * 4) =
          static_cast<uint32_t>(pack_e4m3x2(fa.x * kHi8Scale, fa.y * kHi8Scale)) |
          (static_cast<uint32_t>(pack_e4m3x2(fb.x * kHi8Scale, fb.y * kHi8Scale)) << 16);
      // patch entries at this position: dot(Y, Wf/32) = dot(y1, Wf), the row still being in registers
      for (int e = s_ptr[i]; e < s_ptr[i + 1]; ++e) {
        const float4 wv = e < kL1MaxEnt ? *reinterpret_cast<const float4*>(s_wf + e * kC + lane * 4)
                                        : __ldg(reinterpret_cast<const float4*>(ent_w + static_cast<size_t>(s_lo + e) * kC) + lane);
        float acc = a.x * wv.x;
        acc = fmaf(a.y, wv.y, acc); acc = fmaf(a.z, wv.z, acc); acc = fmaf(a.w, wv.w, acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) part[static_cast<size_t>(w) * part_ld + s_lo + e] = acc;
      }
    }
  }
}

}  // namespace gnm
