// cv2.resize(INTER_CUBIC) for float32 images (warp_frame_latent, pdcnet_of.py:19-32, resizes the 4-channel
// latent to frame size and back around the warp).  OpenCV's algorithm (imgproc/resize.cpp, float path): per destination
// column f = float((dx + 0.5) * scale_x - 0.5) with scale_x = 1 / (dst_w / src_w) evaluated in double, s = floor(f), the four
// taps s-1 .. s+2 replicated at the borders, weights = interpolateCubic(f - s) in float32 (A = -0.75, w3 = 1 - w0 - w1 - w2);
// a horizontal pass into float rows (taps summed left to right) and then the vertical pass.  No antialiasing (INTER_CUBIC
// has none: an 8x down-scale samples 4x4 source pixels per output).  Pinned against cv2 by tests/test_oracle_warp.py
// through the NumPy restatement; the x8 / /8 ratios the reference uses agree with cv2 to 2e-7 of the image range
// (OpenCV's SIMD / IPP paths fuse and reorder the same float operations).
//
// HBM-bound byte work, but tiny (a 4 x 96 x 64 latent <-> 4 x 768 x 512): one thread per output pixel, all channels.
#include "sdof_common.cuh"

namespace sdof {

__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
  const float A = -0.75f;
  const float xp = __fadd_rn(x, 1.0f), xm = __fsub_rn(1.0f, x);
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp), 5.0f * A), xp), 8.0f * A), xp), 4.0f * A);
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, x), A + 3.0f), x), x), 1.0f);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, xm), A + 3.0f), xm), xm), 1.0f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c[0]), c[1]), c[2]);
}

__device__ __forceinline__ void axis_taps(int d, double scale, int n_src, int (&idx)[4], float (&w)[4]) {
  const float f = (float)(((double)d + 0.5) * scale - 0.5);
  const float fl = floorf(f);
  const int s = (int)fl;
  cubic_coeffs(__fsub_rn(f, fl), w);
#pragma unroll
  for (int k = 0; k < 4; ++k) idx[k] = min(max(s - 1 + k, 0), n_src - 1);
}

template <int C_T>
__global__ void __launch_bounds__(256) resize_cubic_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int Hs,
                                                               int Ws, int C_rt, int Hd, int Wd, double scale_y, double scale_x) {
  const int C = C_T ? C_T : C_rt;
  const long long npix = (long long)B * Hd * Wd;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(p % Wd);
    const long long t = p / Wd;
    const int dy = (int)(t % Hd);
    const int b = (int)(t / Hd);
    int xi[4], yi[4];
    float xw[4], yw[4];
    axis_taps(dx, scale_x, Ws, xi, xw);
    axis_taps(dy, scale_y, Hs, yi, yw);
    const float* img = src + (long long)b * Hs * Ws * C;
    for (int c0 = 0; c0 < C; c0 += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const float* row = img + (long long)yi[ky] * Ws * C;
        float r[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c = c0 + cc;
          if (c < C) {
            float v = __fmul_rn(__ldg(row + (long long)xi[0] * C + c), xw[0]);
            v = __fadd_rn(v, __fmul_rn(__ldg(row + (long long)xi[1] * C + c), xw[1]));
            v = __fadd_rn(v, __fmul_rn(__ldg(row + (long long)xi[2] * C + c), xw[2]));
            v = __fadd_rn(v, __fmul_rn(__ldg(row + (long long)xi[3] * C + c), xw[3]));
            r[cc] = v;
          } else {
            r[cc] = 0.f;
          }
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) acc[cc] = ky == 0 ? __fmul_rn(r[cc], yw[0]) : __fadd_rn(acc[cc], __fmul_rn(r[cc], yw[ky]));
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        if (c0 + cc < C) dst[p * C + c0 + cc] = acc[cc];
    }
  }
}

}  // namespace sdof

extern "C" {

int sdof_resize_cubic_f32(const float* src, int B, int Hs, int Ws, int C, int Hd, int Wd, float* dst, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(src && dst, "sdof_resize_cubic_f32: NULL pointer");
  SDOF_REQUIRE(B >= 0 && Hs >= 1 && Ws >= 1 && Hd >= 1 && Wd >= 1 && C >= 1, "sdof_resize_cubic_f32: bad sizes B=%d %dx%d -> %dx%d C=%d", B, Hs, Ws,
               Hd, Wd, C);
  if (B == 0) return SDOF_OK;
  const double scale_x = 1.0 / ((double)Wd / (double)Ws), scale_y = 1.0 / ((double)Hd / (double)Hs);
  const long long npix = (long long)B * Hd * Wd;
  const int grid = grid_for(npix, 256, 8);
  if (C == 4)
    resize_cubic_f32_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(src, dst, B, Hs, Ws, C, Hd, Wd, scale_y, scale_x);
  else
    resize_cubic_f32_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(src, dst, B, Hs, Ws, C, Hd, Wd, scale_y, scale_x);
  SDOF_LAUNCH_CHECK("resize_cubic_f32_kernel");
  return SDOF_OK;
}

}  // extern "C"
