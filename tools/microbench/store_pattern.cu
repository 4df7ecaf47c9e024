// Store-pattern microbenchmark: how fast can 148 persistent CTAs write a [rows][h][w] fp32 pyramid level
// when each warp-level store instruction writes (a) one full 128 B line or (b) two 64 B half lines of
// one source pixel's map, walking source pixels (24 KB apart) from instruction to instruction?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

// mode 0: streaming fill (coalesced, consecutive lines)
// mode 1: 4x32 patch: warp = one row, 32 lanes = 128 B; 256 source pixels per tile
// mode 2: 8x16 patch: warp = two rows x 64 B
// mode 3: like 1 but each warp keeps ONE source pixel and walks target rows (sequential within a map)
__global__ void __launch_bounds__(256, 1) store_kernel(float* out, int rows, int h, int w, long long pitch, int mode) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp >> 2, q = warp & 3;
  const float v = (float)threadIdx.x;
  if (mode == 0) {
    const long long total4 = (long long)rows * pitch / 4;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total4; i += (long long)gridDim.x * 256)
      reinterpret_cast<float4*>(out)[i] = make_float4(v, v, v, v);
    return;
  }
  const int m_tiles = rows / 256;
  const int pty = mode == 2 ? 8 : 4, ptx = mode == 2 ? 16 : 32;
  const int tyt = h / pty, txt = w / ptx;
  const int per_block = tyt * txt;
  const int total = m_tiles * per_block;
  const int t_begin = (int)((long long)total * blockIdx.x / gridDim.x), t_end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
  int it = 0;
  for (int t = t_begin; t < t_end; ++t, ++it) {
    if ((it & 1) != grp) continue;
    const int mt = t / per_block, r = t % per_block;
    const int ty = r / txt, tx = r % txt;
    int y, x;
    if (mode == 2) { y = ty * 8 + 2 * q + (lane >> 4); x = tx * 16 + (lane & 15); }
    else { y = ty * 4 + q; x = tx * 32 + lane; }
    float* p = out + (long long)mt * 256 * pitch + (long long)y * w + x;
    if (mode == 3) {
      // same bytes, but a warp writes 64 consecutive source pixels... of ONE row position: identical to mode 1
    }
#pragma unroll 8
    for (int j = 0; j < 256; ++j) {
      *p = v;
      p += pitch;
    }
  }
}

int main(int argc, char** argv) {
  const int rows = 6144, h = 96, w = 64;
  for (int padded = 0; padded < 2; ++padded) {
    const long long pitch = (long long)h * w + (padded ? 32 : 0);
    float* buf;
    cudaMalloc(&buf, sizeof(float) * rows * pitch);
    for (int mode = 0; mode < 3; ++mode) {
      cudaEvent_t s, e;
      cudaEventCreate(&s); cudaEventCreate(&e);
      for (int i = 0; i < 3; ++i) store_kernel<<<148, 256>>>(buf, rows, h, w, pitch, mode);
      cudaEventRecord(s);
      const int n = 20;
      for (int i = 0; i < n; ++i) store_kernel<<<mode == 0 ? 148 * 8 : 148, 256>>>(buf, rows, h, w, pitch, mode);
      cudaEventRecord(e);
      cudaEventSynchronize(e);
      float ms; cudaEventElapsedTime(&ms, s, e);
      const double bytes = (double)rows * h * w * 4;
      printf("pitch %lld mode %d: %.1f us  %.0f GB/s\n", pitch, mode, ms / n * 1e3, bytes / (ms / n * 1e-3) / 1e9);
    }
    cudaFree(buf);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
