// tcgen05.mma issue/execute rate with BOTH operands in shared memory (SS mode, SWIZZLE_128B, K-major, cta_group::1):
// the floor of the correlation kernel's tile loop.  One CTA per SM; one thread issues `n` MMAs of M=128 x N x K=16 (fp16,
// fp32 accumulate) back to back over a 4-slab operand block (the kernel's access pattern: 4 MMAs per 128-byte slab), then
// one tcgen05.commit; cycles from the first issue to the commit's arrival.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Isd_animation_optical_flow_b200/csrc tools/microbench/mma_rate.cu -o tools/microbench/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tc_ptx.cuh"

using namespace sdof;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* cycles, int n_mma, int N, int a_bytes_per_slab, int same_slab, int mode) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar_slab[2];   // [0]: per-slab commits land here; [1]: already complete, polled per slab
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = base;               // 4 slabs x 256 rows x 128 B
  const uint32_t smem_a = base + 4 * 32768;   // 4 slabs x 128 rows x 128 B
  // operands: zeros (no denormal / NaN side effects)
  for (uint32_t i = threadIdx.x; i < (4 * 32768 + 4 * 16384) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar_slab[0]), 1u << 20);
    mbar_init(smem_u32(&bar_slab[1]), 1);
    mbar_arrive(smem_u32(&bar_slab[1]));   // phase 0 complete: a wait on parity 0 returns at once
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_fmt(0, 128, (uint32_t)N);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      // mode 1: a tcgen05.commit after every slab (4 MMAs); mode 2: plus the kernel's per-slab mbarrier wait + fence;
      // mode 3: mode 2 with the clock-free fast path of the wait
      if ((i & 3) == 0 && mode >= 2) {
        if (mode == 2) {
          mbar_wait(smem_u32(&bar_slab[1]), 0);
        } else {
          uint32_t done;
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                       : "=r"(done) : "r"(smem_u32(&bar_slab[1])), "r"(0) : "memory");
          if (!done) mbar_wait(smem_u32(&bar_slab[1]), 0);
        }
        tc_fence_after();
      }
      const int slab = same_slab ? 0 : (i >> 2) & 3, j = i & 3;
      const uint64_t adesc = make_smem_desc_sw128(smem_a + slab * a_bytes_per_slab) + 2 * j;
      const uint64_t bdesc = make_smem_desc_sw128(smem_b + slab * 32768) + 2 * j;
      tc_mma<true>(tmem_base + ((i >> 4) & 1) * 256, adesc, bdesc, idesc, (i & 15) ? 1u : 0u);
      if ((i & 3) == 3 && mode >= 1) tc_commit(smem_u32(&bar_slab[0]));
    }
    const long long t1 = clock64();
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    cycles[2 * blockIdx.x] = t1 - t0;
    cycles[2 * blockIdx.x + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 2 * 148 * sizeof(long long));
  const int smem = 4 * 32768 + 4 * 16384 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {148})
    for (int N : {256})
      for (int mode : {0, 1, 2, 3}) {
        const int same = 0;
        const int n = 2048;
        mma_rate_kernel<<<grid, 128, smem>>>(d, n, N, 16384, same, mode);
        mma_rate_kernel<<<grid, 128, smem>>>(d, n, N, 16384, same, mode);
        cudaDeviceSynchronize();
        long long h[2 * 148];
        cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
        double issue = 0, done = 0;
        for (int i = 0; i < grid; ++i) {
          issue += (double)h[2 * i] / grid;
          done += (double)h[2 * i + 1] / grid;
        }
        printf("CTAs %3d  M=128 N=%3d K=16  %s : issue %6.1f cycles per MMA, completion %6.1f cycles per MMA (nominal %d)  %s\n", grid, N,
               mode == 0 ? "back to back              " : mode == 1 ? "+ commit per slab         " : mode == 2 ? "+ mbarrier wait and fence " : "+ clock-free wait         ", issue / n, done / n, N / 2, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
