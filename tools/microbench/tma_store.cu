// Ceiling of a TMA-store epilogue for the fp16 correlation pyramid: 148 persistent CTAs write [maps][h][w] halves with
// cp.async.bulk.tensor stores of one {32 x, 4 y, 64 maps} box (16 KB, 64-byte inner rows -- the kernel's 32x4 patch times a
// column quarter) from shared memory, tiles in the kernel's order.  `inflight` = bulk groups allowed in flight before the
// issuing thread waits for the shared-memory reads (1 = single staging buffer, 2 = double buffered, 0 = never wait).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/microbench/tma_store.cu -lcuda -o tools/microbench/tma_store
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(128, 1) tma_store_kernel(const __grid_constant__ CUtensorMap map, int rows, int h, int w, int inflight,
                                                          int issuers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 4 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int txt = w / 32, tyt = h / 4, per_block = txt * tyt;
  const int total = (rows / 256) * per_block;
  const int t_begin = (int)((long long)total * blockIdx.x / gridDim.x), t_end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // `issuers` warps share the column quarters of every tile (1: one thread issues all four boxes)
  if (warp < issuers && lane == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem);
    for (int t = t_begin; t < t_end; ++t) {
      const int mt = t / per_block, r = t % per_block;
      const int ty = r / txt, tx = r % txt;
      for (int cq = warp; cq < 4; cq += issuers) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&map), "r"(src + cq * 16384),
                     "r"(tx * 32), "r"(ty * 4), "r"(mt * 256 + cq * 64)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (inflight == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (inflight == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  struct Case { const char* name; int rows, h, w; } cases[] = {{"768x512 level 0 (75 MB)", 6144, 96, 64}, {"720x1280 level 0 (415 MB)", 14336, 88, 160}};
  for (const Case& c : cases) {
    const long long bytes = (long long)c.rows * c.h * c.w * 2;
    void* buf;
    cudaMalloc(&buf, bytes);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)c.w, (cuuint64_t)c.h, (cuuint64_t)c.rows};
    cuuint64_t strides[2] = {(cuuint64_t)c.w * 2, (cuuint64_t)c.w * c.h * 2};
    cuuint32_t box[3] = {32, 4, 64};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
      return 1;
    }
    cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384);
    for (int issuers : {1, 4})
      for (int inflight : {0, 2, 1}) {
        cudaEvent_t s, e;
        cudaEventCreate(&s);
        cudaEventCreate(&e);
        for (int i = 0; i < 3; ++i) tma_store_kernel<<<148, 128, 4 * 16384>>>(map, c.rows, c.h, c.w, inflight, issuers);
        const int n = 20;
        cudaEventRecord(s);
        for (int i = 0; i < n; ++i) tma_store_kernel<<<148, 128, 4 * 16384>>>(map, c.rows, c.h, c.w, inflight, issuers);
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms;
        cudaEventElapsedTime(&ms, s, e);
        const double us = ms / n * 1e3;
        printf("%-28s issuing threads %d, groups in flight %s : %7.1f us  %5.0f GB/s  (%s)\n", c.name, issuers,
               inflight == 0 ? "unbounded" : inflight == 2 ? "2        " : "1        ", us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
      }
    cudaFree(buf);
  }
  return 0;
}
