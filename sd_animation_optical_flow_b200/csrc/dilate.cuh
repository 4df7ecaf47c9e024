// Shared-memory tile dilation with OpenCV's MORPH_ELLIPSE footprint (used by mask.cu
// and fused.cu).  A CTA owns a TW x TH tile; the tile plus an r-pixel halo is staged
// in shared memory as bytes by a caller-supplied loader, then each thread produces 4
// consecutive pixels with packed byte max (vmaxu4) over unaligned 32-bit smem reads.
#pragma once

#include "sdof_common.cuh"

namespace sdof {

constexpr int kDilTW = 64;    // tile width  (multiple of 4)
constexpr int kDilTH = 16;    // tile height
constexpr int kDilMaxK = 31;  // largest structuring element
constexpr int kDilThreads = (kDilTW / 4) * kDilTH;  // 256

struct EllipseRows {
  int ksize;
  signed char half[kDilMaxK];  // row half-widths (cv2.getStructuringElement)
};

inline int make_ellipse(int ksize, EllipseRows* e) {
  if (ksize < 1 || !(ksize & 1) || ksize > kDilMaxK) return SDOF_ERR_INVALID;
  int32_t hw[kDilMaxK];
  ellipse_half_widths_host(ksize, hw);
  e->ksize = ksize;
  for (int i = 0; i < kDilMaxK; ++i) e->half[i] = (signed char)(i < ksize ? hw[i] : -1);
  return SDOF_OK;
}

// smem geometry for radius r: tile pixel (0,0) sits at column r4 = roundup(r,4), row r.
__host__ __device__ inline int dil_r4(int r) { return (r + 3) & ~3; }
__host__ __device__ inline int dil_pitch(int r) { return kDilTW + 2 * dil_r4(r); }
__host__ __device__ inline int dil_rows(int r) { return kDilTH + 2 * r; }
inline size_t dil_smem_bytes(int r) { return (size_t)dil_pitch(r) * dil_rows(r); }

// Stage the tile + halo.  loader(gy, gx) -> byte for an in-image pixel; pixels outside the
// image read 0 (cv2.dilate's default border never wins the max).
template <typename Loader>
__device__ __forceinline__ void dil_stage(unsigned char* tile, int r, int ty0, int tx0, int H, int W, Loader loader) {
  const int r4 = dil_r4(r), pitch = dil_pitch(r), rows = dil_rows(r);
  const int cols = kDilTW + 2 * r;  // columns actually needed
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    const int ly = i / cols, lx = i - ly * cols;
    const int gy = ty0 - r + ly, gx = tx0 - r + lx;
    unsigned char v = 0;
    if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) v = loader(gy, gx);
    tile[ly * pitch + (r4 - r) + lx] = v;
  }
}

// Dilated bytes of pixels (lx..lx+3, ly) of the tile, lx % 4 == 0.
__device__ __forceinline__ unsigned dil_apply4(const unsigned char* tile, const EllipseRows& e, int lx, int ly) {
  const int r = e.ksize >> 1;
  const int r4 = dil_r4(r), pitch = dil_pitch(r);
  unsigned acc = 0;
  for (int i = 0; i < e.ksize; ++i) {
    const int half = e.half[i];
    const unsigned char* row = tile + (ly + i) * pitch + r4 + lx;  // column of dx = 0
    for (int dx = -half; dx <= half; ++dx) {
      const int a = dx & 3;
      const unsigned* q = reinterpret_cast<const unsigned*>(row + (dx - a));
      const unsigned lo = q[0];
      const unsigned hi = a ? q[1] : 0u;
      acc = __vmaxu4(acc, __funnelshift_r(lo, hi, a * 8));
    }
  }
  return acc;
}

}  // namespace sdof
