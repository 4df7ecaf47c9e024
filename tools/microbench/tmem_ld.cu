// TMEM -> register read rate on this part (tcgen05.ld.32x32b), the limiter of the correlation kernel's epilogue:
// one CTA per SM, W warps (warp w reads lane quarter w % 4), every warp loops over the 512 allocated columns with K loads of
// 32 columns in flight before each tcgen05.wait::ld.  Reports bytes per SM clock.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Isd_animation_optical_flow_b200/csrc tools/microbench/tmem_ld.cu -o tools/microbench/tmem_ld
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tc_ptx.cuh"

using namespace sdof;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32], int o) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]),
                 "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])
               : "r"(taddr)
               : "memory");
}

// K = x32 loads in flight per warp; HALF: issue the 32 columns as two x16 instructions
template <int K, bool HALF>
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(long long* cycles, uint32_t* sink, int reps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[K][32];
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    // every warp walks all 512 columns; warps sharing a lane quarter start at different columns
    for (int c = 0; c < 512; c += 32 * K) {
      const int col = (c + (warp >> 2) * 128) & 511;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (HALF) {
          tmem_ld16(lane_base + ((col + 32 * k) & 511), r[k], 0);
          tmem_ld16(lane_base + ((col + 32 * k + 16) & 511), r[k], 16);
        } else {
          tmem_ld32(lane_base + ((col + 32 * k) & 511), r[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) tmem_ld_wait(r[k]);
#pragma unroll
      for (int k = 0; k < K; ++k) acc ^= r[k][0] ^ r[k][31];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int K, bool HALF>
static void run(int warps, long long* d_cycles, uint32_t* d_sink) {
  const int reps = 64;
  tmem_ld_kernel<K, HALF><<<148, warps * 32>>>(d_cycles, d_sink, reps);
  tmem_ld_kernel<K, HALF><<<148, warps * 32>>>(d_cycles, d_sink, reps);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += (double)h[i] / 148;
  const double bytes = (double)reps * warps * 512 * 32 * 4;   // per SM
  printf("warps %2d  in flight %d x %s : %8.0f cycles  %6.1f B/clk per SM  (%s)\n", warps, K, HALF ? "2 x16" : "x32", mean, bytes / mean,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  for (int warps : {4, 8, 16}) {
    run<1, false>(warps, d_cycles, d_sink);
    run<2, false>(warps, d_cycles, d_sink);
    run<1, true>(warps, d_cycles, d_sink);
  }
  return 0;
}
