// Microbenchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16) from one thread, as a function of N and of
// how the descriptors are produced.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate.bin mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;               // 64 KB of A
  uint8_t* sb = smem + 65536;       // 64 KB of B
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t a_lo = ((smem_u32(sa) >> 4) & 0x3FFFu) | 0x10000u;
      const uint32_t b_lo = ((smem_u32(sb) >> 4) & 0x3FFFu) | 0x10000u;
      const long long t0 = clock64();
      if (MODE == 0) {            // constant descriptors
        const uint64_t a = ((uint64_t)hi << 32) | a_lo, b = ((uint64_t)hi << 32) | b_lo;
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) umma(tbase, a, b, idesc, 1u);
        }
      } else if (MODE == 1) {     // descriptor low words move like a K loop (4 k16 steps) over 9 taps
        for (int i = 0; i < iters; i += 8) {
          const uint32_t tap = (uint32_t)(i >> 3) % 9u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma(tbase + (uint32_t)((j >> 2) * N), ((uint64_t)hi << 32) | (a_lo + tap * 8u + (j >> 2) * 64u + 2u * (j & 3)),
                 ((uint64_t)hi << 32) | (b_lo + tap * 256u + 2u * (j & 3)), idesc, 1u);
        }
      } else if (MODE == 3 || MODE == 4) {
        // the conv kernel's A operand: 8-pixel row groups of a 16x8 tile inside an 18-pixel-wide halo box (SBO = 18 rows of
        // 128 B = 2304 B), tap (ky, kx) shifts the start by (ky * 18 + kx) rows (MODE 3) / no tap shift (MODE 4)
        const uint32_t hi_halo = ((2304u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
        for (int i = 0; i < iters; i += 8) {
          const uint32_t tap = (uint32_t)(i >> 3) % 9u;
          const uint32_t shift = (MODE == 3) ? ((tap / 3u) * 18u + (tap % 3u)) * 8u : 0u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma(tbase + (uint32_t)((j >> 2) * N), ((uint64_t)hi_halo << 32) | (a_lo + shift + (j >> 2) * 64u + 2u * (j & 3)),
                 ((uint64_t)hi << 32) | (b_lo + tap * 256u + 2u * (j & 3)), idesc, 1u);
        }
      } else if (MODE >= 5) {
        // the conv kernel's per-item structure: groups of G MMAs (G = 72: one 64 -> 64 item; 36: one 32 -> 32 item), the
        // first MMAs of a group overwrite their accumulators, accumulators alternate between groups, and every group ends
        // with tcgen05.commit to a barrier (MODE 5: one commit; MODE 6: two, as a_empty + tfull; MODE 7: none)
        const int G = (N == 32) ? 36 : 72;
        const uint64_t a = ((uint64_t)hi << 32) | a_lo, b = ((uint64_t)hi << 32) | b_lo;
        for (int i = 0; i < iters; i += G) {
          const uint32_t d = tbase + (uint32_t)(((i / G) & 1) * 2 * N);
          for (int j = 0; j < G; ++j) umma(d + (uint32_t)(((j >> 2) & 1) * N), a, b, idesc, j < 8 ? 0u : 1u);
          if (MODE == 5 || MODE == 6) commit(&bar2[0]);
          if (MODE == 6) commit(&bar2[1]);
        }
      } else {                    // two alternating accumulators, constant descriptors
        const uint64_t a = ((uint64_t)hi << 32) | a_lo, b = ((uint64_t)hi << 32) | b_lo;
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) umma(tbase + (uint32_t)((j & 1) * N), a, b, idesc, 1u);
        }
      }
      const long long t1 = clock64();
      commit(&bar);
      while (!try_wait(&bar, 0)) {}
      const long long t2 = clock64();
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------
// cta_group::2: a CTA pair (cluster of 2) executes ONE MMA of M = 256: each CTA supplies its own 128 A rows and HALF of the
// B rows (N / 2) from its own shared memory, the leader CTA's thread issues.  Per SM and MMA the operand fetch is
// 4096 + 16 N bytes instead of 4096 + 32 N.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {      // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;
  uint8_t* sb = smem + 65536;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  if (warp == 1 && rank == 0) {
    if (elect_one()) {
      // M = 256 (field M >> 4 = 16), N columns
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t a_lo = ((smem_u32(sa) >> 4) & 0x3FFFu) | 0x10000u;
      const uint32_t b_lo = ((smem_u32(sb) >> 4) & 0x3FFFu) | 0x10000u;
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
        const uint32_t tap = (uint32_t)(i >> 3) % 9u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma2(tbase + (uint32_t)((j >> 2) * N), ((uint64_t)hi << 32) | (a_lo + tap * 8u + (j >> 2) * 64u + 2u * (j & 3)),
                ((uint64_t)hi << 32) | (b_lo + tap * 128u + 2u * (j & 3)), idesc, 1u);
      }
      const long long t1 = clock64();
      commit2(&bar);
      while (!try_wait(&bar, 0)) {}
      const long long t2 = clock64();
      out[(blockIdx.x >> 1) * 2] = t1 - t0;
      out[(blockIdx.x >> 1) * 2 + 1] = t2 - t0;
    }
  } else if (warp == 1 && rank == 1) {
    while (!try_wait(&bar, 0)) {}        // the multicast commit also completes the peer's barrier
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
  }
}

static void run2(int N, int grid, int iters) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 131072 + 1024;
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) rate2_kernel<<<grid, 128, smem>>>(N, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cta_group::2 N %d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  const int pairs = grid / 2;
  long long* h = new long long[pairs * 2];
  cudaMemcpy(h, d, pairs * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < pairs; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  printf("cta_group::2 M=256 N=%3d grid=%3d : issue %.1f clk/MMA   complete %.1f clk/MMA   (math floor %d, operand bytes/SM %d -> %d clk)\n",
         N, grid, issue / pairs / iters, total / pairs / iters, N / 2, 4096 + 16 * N, (4096 + 16 * N) / 128);
  delete[] h; cudaFree(d);
}

template <int MODE>
static void run(int N, int grid, int iters) {
  long long* d;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  const int smem = 131072 + 1024;
  cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<MODE><<<grid, 128, smem>>>(N, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d N %d: %s\n", MODE, N, cudaGetErrorString(e)); exit(1); }
  long long* h = new long long[grid * 2];
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  printf("mode %d  N=%3d grid=%3d : issue %.1f clk/MMA   complete %.1f clk/MMA   (floor %d)\n", MODE, N, grid,
         issue / grid / iters, total / grid / iters, N / 2);
  delete[] h; cudaFree(d);
}

int main(int argc, char** argv) {
  const int iters = 4096;
  if (argc > 1 && argv[1][0] == '2') {          // `mma_rate.bin 2`: only the cta_group::2 table
    for (int g : {2, 148}) for (int N : {32, 64, 128, 256}) run2(N, g, iters);
    return 0;
  }
  const int Ns[5] = {32, 64, 128, 192, 256};
  for (int g : {1, 148}) {
    for (int N : Ns) run<0>(N, g, iters);
    for (int N : Ns) if (2 * N <= 512) run<1>(N, g, iters);
    for (int N : Ns) if (2 * N <= 512) run<2>(N, g, iters);
    for (int N : Ns) if (2 * N <= 512) run<3>(N, g, iters);
    for (int N : Ns) if (2 * N <= 512) run<4>(N, g, iters);
    for (int N : {32, 64, 128}) { run<5>(N, g, 4032); run<6>(N, g, 4032); run<7>(N, g, 4032); }
  }
  return 0;
}
