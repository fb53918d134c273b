// tcgen05.mma issue-rate microbenchmark (sm_100a).  One CTA (or one CTA pair) per SM issues a long
// chain of MMAs on operands already resident in shared memory (contents irrelevant) and reports
// cycles per instruction.  Used to pick the UMMA shape / operand roles for the conv kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_microbench tools/mma_microbench.cu && ./mma_microbench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../genomad_b200/csrc/common.cuh"

using namespace gnm;

__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct Result { long long cycles; int n; };

// cta_group::1, M=128, N=n_dim.  mode bits: 1 = alternate accumulator every MMA (else same accumulator),
// 2 = walk A over taps/K-halves like the conv kernel, 4 = a second warp streams tcgen05.ld concurrently,
// 8 = other warps keep writing shared memory (stand-in for TMA fill traffic).
// 16 = tcgen05.commit after every 8 MMAs, 32 = leave/re-enter the elected region (fence + elect + syncwarp) every
// 8 MMAs like the kernel's per-stage loop, 64 = a warp streams cp.async.bulk (TMA engine) 16 KB copies into smem.
__global__ void __launch_bounds__(256, 1) bench_cg1(Result* out, int n_dim, int mode, int iters, const uint8_t* gsrc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t ring[8];
  __shared__ uint64_t tbar;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&tbar, 1); for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1); fence_barrier_init(); s_done = 0; }
  fence_proxy_async_smem();
  if (warp == 1) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = s_tmem;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_f16(128, n_dim);
    const uint64_t d0 = umma_desc_sw128(0);
    const uint32_t a_addr = smem_u32(smem);                       // A: 4 regions x 272 rows x 128 B = 136 KB
    const uint32_t b_addr = smem_u32(smem) + 136 * 1024;          // B: 256 rows x 128 B (N=256) = 32 KB, only first used
    long long t0 = clock64();
    if (mode & 32) {
      for (int i = 0; i < iters; i += 2) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t tap = (mode & 2) ? ((i + j) % 6) : 0, reg = (mode & 2) ? (((i + j) / 6) & 3) : 0;
            const uint64_t ad = d0 + ((a_addr + reg * 34816 + tap * 128) >> 4);
            const uint64_t bd = d0 + (b_addr >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_f16(tb, ad + kk * 2, bd + kk * 2, idesc, 1u);
          }
          if (mode & 16) umma_commit(&ring[(i >> 1) & 7]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
    } else if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        const uint32_t tap = (mode & 2) ? (i % 6) : 0, reg = (mode & 2) ? ((i / 6) & 3) : 0;
        const uint64_t ad = d0 + ((a_addr + reg * 34816 + tap * 128) >> 4);
        const uint64_t bd = d0 + (b_addr >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t acc = (mode & 1) ? tb + ((kk & 1) * n_dim) : tb;
          umma_f16(acc, ad + kk * 2, bd + kk * 2, idesc, 1u);
        }
        if ((mode & 16) && (i & 1)) umma_commit(&ring[(i >> 1) & 7]);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0, nullptr, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out->cycles = t1 - t0; out->n = iters * 4; }
    if (lane == 0) s_done = 1;
  } else if (warp >= 4 && (mode & 4)) {
    uint32_t r[32];
    uint32_t sink = 0;
    while (!s_done) {
      tmem_ld_32x32(tb + (static_cast<uint32_t>((warp - 4) * 32) << 16) + 256, r);
      tmem_wait_ld();
      sink += r[0] + r[31];
    }
    if (sink == 0x12345678u) out->n = -1;
  } else if (warp == 3 && (mode & 64)) {
    if (lane == 0) {
      uint32_t ph = 0, k = 0;
      while (!s_done) {
        mbar_arrive_expect_tx(&tbar, 16384);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + 168 * 1024 + (k & 1) * 16384)), "l"(gsrc + (size_t)((blockIdx.x * 64 + (k & 63)) * 16384)), "r"(16384), "r"(smem_u32(&tbar)) : "memory");
        mbar_wait(&tbar, ph, nullptr, 0);
        ph ^= 1; ++k;
      }
    }
  } else if (warp >= 2 && warp < 4 && (mode & 8)) {
    uint4* dst = reinterpret_cast<uint4*>(smem + 168 * 1024);    // 32 KB scratch beyond the operands
    uint32_t k = 0;
    while (!s_done) {
      dst[(k * 64 + (warp - 2) * 32 + lane) & 2047] = make_uint4(k, k, k, k);
      ++k;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

// cta_group::2: M=256 (128 rows per CTA), N=n_dim (each CTA holds n_dim/2 rows of B).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench_cg2(Result* out, int n_dim, int iters, int n_acc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tb = s_tmem;
  if (rank == 0 && warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(256, n_dim);
    const uint32_t a_addr = smem_u32(smem);
    const uint32_t b_addr = smem_u32(smem) + 40 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_f16_cg2(tb + (i % n_acc) * n_dim, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), idesc, 1u);
    }
    umma_commit_cg2_mc(&bar, 3);
    mbar_wait(&bar, 0, nullptr, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) { out->cycles = t1 - t0; out->n = iters * 4; }
  } else if (warp == 0 && lane == 0) {
    mbar_wait(&bar, 0, nullptr, 0);       // peer CTA: its barrier is signalled by the multicast commit
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
  }
}

int main() {
  Result* d; Result h;
  cudaMalloc(&d, sizeof(Result));
  const int smem = 204 * 1024;
  cudaFuncSetAttribute(bench_cg1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench_cg2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  uint8_t* gsrc; cudaMalloc(&gsrc, (size_t)nsm * 64 * 16384); cudaMemset(gsrc, 0, (size_t)nsm * 64 * 16384);
  struct { int n, mode; } cfg1[] = {{128, 0}, {128, 2}, {128, 16}, {128, 32}, {128, 48}, {128, 50}, {128, 64}, {128, 66}, {128, 64 + 48 + 2 + 4}, {256, 0}, {256, 48}, {256, 64 + 48 + 2 + 4}};
  for (auto c : cfg1) {
    bench_cg1<<<nsm, 256, smem>>>(d, c.n, c.mode, iters, gsrc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("cg1 M=128 N=%3d mode=%2d : %7.1f cycles/MMA (K=16)  -> %5.1f %% of 8192 flop/cyc/SM   [%s]\n", c.n, c.mode,
           double(h.cycles) / h.n, 100.0 * (2.0 * 128 * c.n * 16 / (double(h.cycles) / h.n)) / 8192.0, cudaGetErrorString(e));
  }
  struct { int n, nacc; } cfg2[] = {{128, 1}, {128, 2}, {256, 1}, {256, 2}};
  for (auto c : cfg2) {
    for (int grid : {2, nsm}) {
      bench_cg2<<<grid, 128, smem>>>(d, c.n, iters, c.nacc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
      printf("cg2 M=256 N=%3d acc=%d grid=%3d : %7.1f cycles/MMA (K=16)  -> %5.1f %% of 8192 flop/cyc/SM (per SM) [%s]\n", c.n, c.nacc, grid,
             double(h.cycles) / h.n, 100.0 * (2.0 * 128 * c.n * 16 / (double(h.cycles) / h.n)) / 8192.0, cudaGetErrorString(e));
    }
  }
  return 0;
}
