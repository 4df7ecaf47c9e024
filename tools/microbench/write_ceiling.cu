// Write-only ceilings of the B200 memory system, for the roofline of the correlation-pyramid epilogue (a pure store stream):
//   fill     : grid-stride 16-byte stores over consecutive lines (the friendliest write stream there is)
//   memset   : cudaMemsetAsync of the same buffer
//   pyr16    : the fp16-stored pyramid's address pattern -- 148 persistent CTAs x 16 warps; a warp instruction writes two
//              64-byte runs (x..x+31 of one row of the maps of source pixels j and j+1), 32 instructions walk 64 source
//              pixels (one map = h*w*2 bytes apart); tiles in the order the kernel visits them
// at buffer sizes below and above the 126 MB L2 (a buffer that fits is rewritten in L2; a larger one must drain to HBM).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/write_ceiling.cu -o /tmp/write_ceiling && /tmp/write_ceiling
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(256) fill_kernel(uint4* out, long long n16) {
  const uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n16; i += (long long)gridDim.x * 256) out[i] = v;
}

// rows = source pixels (maps), each map h x w halves; patch = 32 x 4 target pixels, 256 maps per tile
__global__ void __launch_bounds__(512, 1) pyr16_kernel(__half* out, int rows, int h, int w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cq = warp >> 2, q = warp & 3;
  const bool odd = lane & 1;
  const long long pitch = (long long)h * w;
  const int txt = w / 32, tyt = h / 4, per_block = txt * tyt;
  const int total = (rows / 256) * per_block;
  const int t_begin = (int)((long long)total * blockIdx.x / gridDim.x), t_end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
  const __half2 v = __floats2half2_rn((float)lane, 1.f);
  for (int t = t_begin; t < t_end; ++t) {
    const int mt = t / per_block, r = t % per_block;
    const int ty = r / txt, tx = r % txt;
    const int y = ty * 4 + q, xe = (tx * 32 + lane) & ~1;
    __half2* p = reinterpret_cast<__half2*>(out + ((long long)mt * 256 + cq * 64 + (odd ? 1 : 0)) * pitch + (long long)y * w + xe);
#pragma unroll 16
    for (int j = 0; j < 32; ++j) {
      *p = v;
      p += pitch;   // two maps further, in half2 units
    }
  }
}

// patch-blocked layout [source block][patch ty][patch tx][map (256)][py (4)][px (32)]: a tile is one contiguous 64 KB run
__global__ void __launch_bounds__(512, 1) pyr16_blocked_kernel(__half* out, int rows, int h, int w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cq = warp >> 2, q = warp & 3;
  const bool odd = lane & 1;
  const int txt = w / 32, tyt = h / 4, per_block = txt * tyt;
  const int total = (rows / 256) * per_block;
  const int t_begin = (int)((long long)total * blockIdx.x / gridDim.x), t_end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
  const __half2 v = __floats2half2_rn((float)lane, 1.f);
  for (int t = t_begin; t < t_end; ++t) {
    // tile t = 32768 halves; map m of the tile at m * 128, row q at + q * 32, x at + (lane & ~1)
    __half2* p = reinterpret_cast<__half2*>(out + (long long)t * 32768 + (cq * 64 + (odd ? 1 : 0)) * 128 + q * 32 + (lane & ~1));
#pragma unroll 16
    for (int j = 0; j < 32; ++j) {
      *p = v;
      p += 128;   // two maps further (2 * 128 halves), in half2 units
    }
  }
}

static float time_us(void (*launch)(void*, long long, int, int, int), void* buf, long long bytes, int rows, int h, int w) {
  cudaEvent_t s, e;
  cudaEventCreate(&s);
  cudaEventCreate(&e);
  for (int i = 0; i < 3; ++i) launch(buf, bytes, rows, h, w);
  const int n = 20;
  cudaEventRecord(s);
  for (int i = 0; i < n; ++i) launch(buf, bytes, rows, h, w);
  cudaEventRecord(e);
  cudaEventSynchronize(e);
  float ms;
  cudaEventElapsedTime(&ms, s, e);
  return ms / n * 1e3f;
}

int main() {
  struct Case { const char* name; int rows, h, w; } cases[] = {
      {"768x512 level 0 (75 MB)", 6144, 96, 64}, {"720x1280 level 0 (415 MB)", 14336, 88, 160}, {"2 GB", 65536, 128, 128}};
  for (const Case& c : cases) {
    const long long bytes = (long long)c.rows * c.h * c.w * 2;
    void* buf;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return 1;
    const float t_fill = time_us([](void* b, long long n, int, int, int) { fill_kernel<<<148 * 8, 256>>>((uint4*)b, n / 16); }, buf, bytes, 0, 0, 0);
    const float t_set = time_us([](void* b, long long n, int, int, int) { cudaMemsetAsync(b, 1, n); }, buf, bytes, 0, 0, 0);
    const float t_pyr = time_us([](void* b, long long, int rows, int h, int w) { pyr16_kernel<<<148, 512>>>((__half*)b, rows, h, w); }, buf, bytes,
                                c.rows, c.h, c.w);
    const float t_blk = time_us([](void* b, long long, int rows, int h, int w) { pyr16_blocked_kernel<<<148, 512>>>((__half*)b, rows, h, w); }, buf, bytes,
                                c.rows, c.h, c.w);
    printf("%-28s fill %7.1f us %5.0f GB/s | memset %7.1f us %5.0f GB/s | pyramid pattern %7.1f us %5.0f GB/s | patch-blocked layout %7.1f us %5.0f GB/s\n",
           c.name, t_fill, bytes / t_fill / 1e3, t_set, bytes / t_set / 1e3, t_pyr, bytes / t_pyr / 1e3, t_blk, bytes / t_blk / 1e3);
    cudaFree(buf);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
