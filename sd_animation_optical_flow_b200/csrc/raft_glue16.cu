// fp16-activation forms of the update-loop glue kernels (csrc/raft_glue.cu), for the update block run with cuDNN fp16
// tensor-op convolutions (fp32 accumulation): activations between the convolutions live in fp16 -- the 11 significant bits a
// TF32 convolution keeps of an fp32 operand anyway -- at half the bytes; the hidden state's master copy, the coordinates, the
// flow and the per-pair bias maps stay fp32.  Measured cuDNN time of the twelve convolutions of one iteration at 768x512,
// batch 1: 113 us in TF32, 93 us in fp16 (profiles/r2_conv_probe_768x512.jsonl).
//
//   motion_tail16_h  : hx16[:, 128:254] = relu(mc16 + mf16 + bias)[:, :126], hx16[:, 254:256] = flow     (update.py:95-96)
//   gru_rh_h         : rh16 = sigmoid(zr16[:, 128:256] + map) * h                                         (update.py:48-49, 55-56)
//   gru_update_h     : h = (1 - z) h + z tanh(q16 + zr16[:, 256:384] + map); hx16[:, :128] = h16 = fp16(h) (update.py:50-52, 57-59)
//   conv7x7_c2_relu_h: BasicMotionEncoder.convf1 on the fp32 flow -> fp16 [B,h,w,128]                     (update.py:85,93)
//   flowhead2_taps_h : the taps kernel of FlowHead.conv2 reading fp16 activations                          (update.py:10,14)
#include <cuda_fp16.h>
#include <stdlib.h>

#include "sdof_common.cuh"

namespace sdof {

__device__ __forceinline__ float4 ldh4(const __half* p) {   // 4 halves (8-byte aligned) -> float4
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void sth4(__half* p, float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&a);
  o.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = o;
}
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256) motion_tail16_h_kernel(const __half* __restrict__ mc, const __half* __restrict__ mf,
                                                              const float4* __restrict__ bias, const float2* __restrict__ flow,
                                                              __half* __restrict__ hx16, int hx16_stride, int64_t npix) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = npix * 32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i >> 5;
    const int c4 = (int)(i & 31);
    const float4 a = ldh4(mc + i * 4), u = ldh4(mf + i * 4), bv = __ldg(bias + c4);
    float4 v = make_float4(fmaxf(a.x + u.x + bv.x, 0.f), fmaxf(a.y + u.y + bv.y, 0.f), fmaxf(a.z + u.z + bv.z, 0.f), fmaxf(a.w + u.w + bv.w, 0.f));
    if (c4 == 31) {
      const float2 f = flow[p];
      v.z = f.x;
      v.w = f.y;
    }
    sth4(hx16 + p * hx16_stride + 128 + 4 * c4, v);
  }
}

__global__ void __launch_bounds__(256) gru_rh_h_kernel(const __half* __restrict__ zr, int zr_stride, const float4* __restrict__ zrmap,
                                                       const float4* __restrict__ h, __half* __restrict__ rh16, int64_t npix) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = npix * 32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i >> 5;
    const int c4 = (int)(i & 31);
    const float4 r = ldh4(zr + p * zr_stride + 128 + 4 * c4), m = __ldg(zrmap + p * 64 + 32 + c4), hv = h[i];
    sth4(rh16 + i * 4, make_float4(sigm(r.x + m.x) * hv.x, sigm(r.y + m.y) * hv.y, sigm(r.z + m.z) * hv.z, sigm(r.w + m.w) * hv.w));
  }
}

__global__ void __launch_bounds__(256) gru_update_h_kernel(const __half* __restrict__ zr, int zr_stride, const float4* __restrict__ zrmap,
                                                           const __half* __restrict__ q, const float4* __restrict__ qmap,
                                                           float4* __restrict__ h, __half* __restrict__ hx16, int hx16_stride,
                                                           __half* __restrict__ h16, int64_t npix) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = npix * 32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i >> 5;
    const int c4 = (int)(i & 31);
    const __half* zp = zr + p * zr_stride + 4 * c4;
    const float4 z = ldh4(zp), qx = ldh4(zp + 256), qq = ldh4(q + i * 4);
    const float4 mz = __ldg(zrmap + p * 64 + c4), mq = __ldg(qmap + i), hv = h[i];
    float4 o;
    float s;
    s = sigm(z.x + mz.x); o.x = (1.f - s) * hv.x + s * tanhf(qq.x + qx.x + mq.x);
    s = sigm(z.y + mz.y); o.y = (1.f - s) * hv.y + s * tanhf(qq.y + qx.y + mq.y);
    s = sigm(z.z + mz.z); o.z = (1.f - s) * hv.z + s * tanhf(qq.z + qx.z + mq.z);
    s = sigm(z.w + mz.w); o.w = (1.f - s) * hv.w + s * tanhf(qq.w + qx.w + mq.w);
    h[i] = o;
    sth4(hx16 + p * hx16_stride + 4 * c4, o);
    if (h16) sth4(h16 + i * 4, o);
  }
}

// ---- convf1 (7x7, 2 -> 128) with fp16 output: same tiling as conv7x7_c2_relu_kernel (raft_glue.cu)
constexpr int kH7Tile = 8, kH7Threads = 256, kH7Patch = kH7Tile + 6, kH7K = 98;
constexpr size_t kH7Smem = (size_t)kH7K * 128 * 4 + (size_t)kH7Patch * kH7Patch * 8;

// coords != nullptr: the input is flow = coords + gather(taps) - pixel grid, computed while the patch is staged (the deferred coords
// update of sdof_corr_lookup_gather_h: this kernel runs beside the lookup and must not read what the lookup writes); `flow` unused.
__global__ void __launch_bounds__(kH7Threads) conv7x7_c2_relu_h_kernel(const float2* __restrict__ flow, const float* __restrict__ wT,
                                                                      const float* __restrict__ bias, __half* __restrict__ out, int h, int w,
                                                                      int tiles_x, int tiles_y, const float2* __restrict__ coords = nullptr,
                                                                      const float* __restrict__ taps = nullptr,
                                                                      float2 tap_bias = make_float2(0.f, 0.f)) {
  extern __shared__ __align__(16) unsigned char h7_smem[];
  float* ws = reinterpret_cast<float*>(h7_smem);                                   // [98][128]
  float2* patch = reinterpret_cast<float2*>(h7_smem + (size_t)kH7K * 128 * 4);     // [14][14]
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int trem = tile - b * tiles_x * tiles_y;
  const int ty0 = (trem / tiles_x) * kH7Tile, tx0 = (trem % tiles_x) * kH7Tile;
  for (int i = threadIdx.x; i < kH7K * 128 / 4; i += kH7Threads) reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(wT) + i);
  pdl_wait();        // the 50 KB of filters above are staged while the kernel that produces `flow` drains
  pdl_trigger();
  const float2* fb = (coords ? coords : flow) + (int64_t)b * h * w;
  for (int i = threadIdx.x; i < kH7Patch * kH7Patch; i += kH7Threads) {
    const int py = i / kH7Patch, pxx = i - py * kH7Patch;
    const int y = ty0 + py - 3, x = tx0 + pxx - 3;
    float2 v = make_float2(0.f, 0.f);
    if ((unsigned)y < (unsigned)h && (unsigned)x < (unsigned)w) {
      v = fb[y * w + x];
      if (coords) {
        if (taps) {
          const float2 d = flowhead2_gather(taps, tap_bias, (int64_t)b * h * w + y * w + x, y, x, h, w);
          v.x += d.x;
          v.y += d.y;
        }
        v.x -= (float)x;
        v.y -= (float)y;
      }
    }
    patch[i] = v;
  }
  __syncthreads();
  const int cg = threadIdx.x & 15, pg = threadIdx.x >> 4;
  const int r = pg >> 1, c0 = (pg & 1) * 4;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
    float in[20];
    const float4* prow = reinterpret_cast<const float4*>(patch + (r + ky) * kH7Patch + c0);
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float4 v = prow[q];
      in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int kx = 0; kx < 7; ++kx)
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const float4* wp = reinterpret_cast<const float4*>(ws + ((ky * 7 + kx) * 2 + ci) * 128 + cg * 8);
        const float4 wa = wp[0], wb = wp[1];
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = in[2 * (i + kx) + ci];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(v, wv[c], acc[i][c]);
        }
      }
  }
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + cg * 2), b1 = __ldg(reinterpret_cast<const float4*>(bias) + cg * 2 + 1);
  const int y = ty0 + r;
  if (y >= h) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = tx0 + c0 + i;
    if (x >= w) continue;
    __half* o = out + (((int64_t)b * h + y) * w + x) * 128 + cg * 8;
    sth4(o, make_float4(fmaxf(acc[i][0] + b0.x, 0.f), fmaxf(acc[i][1] + b0.y, 0.f), fmaxf(acc[i][2] + b0.z, 0.f), fmaxf(acc[i][3] + b0.w, 0.f)));
    sth4(o + 4, make_float4(fmaxf(acc[i][4] + b1.x, 0.f), fmaxf(acc[i][5] + b1.y, 0.f), fmaxf(acc[i][6] + b1.z, 0.f), fmaxf(acc[i][7] + b1.w, 0.f)));
  }
}

// ---- flow-head taps on fp16 activations: y[p][tap][co] = <x[p, :], w[tap][co][:]>, a warp owns 2 pixels, lanes split the 256
// channels (8 each), two transposing butterflies reduce the 36 partials (same scheme as flowhead2_taps_kernel<2>)
// ---- convf1 as a tensor-core GEMM: the im2col matrix of the 7x7 x 2-channel convolution, in fp16 with a hi/lo split of the flow
// (flow = hi + lo to 2^-22 relative, so the fp16 operand costs no accuracy; the filter is the fp16-rounded one for both halves).
// Row of pixel p (out_channels halves, >= 392): [tap 0..48: (fx_hi, fy_hi)] [tap 0..48: (fx_lo, fy_lo)] [zeros], tap = ky * 7 + kx,
// out-of-image taps 0 (the convolution's zero padding).  A cuDNN 1x1 convolution over these rows runs in ~4.7 us at 96x64 where
// the 7x7 convolution itself stays on a CUDA-core engine (9-13 us in cuDNN with 4, 8 or 16 padded input channels; 12.6 us for the
// FMA kernel above, whose SM time also delays the correlation branch running beside it).  The flow is coords + gather(taps) - grid
// as in sdof_conv7x7_c2_relu_coords_h.
constexpr int kI2cTile = 8, kI2cThreads = 256, kI2cPatch = kI2cTile + 6;

__global__ void __launch_bounds__(kI2cThreads) flow_im2col7_h_kernel(const float2* __restrict__ coords, const float* __restrict__ taps, float2 tap_bias,
                                                                    __half2* __restrict__ out, int words_per_px, int h, int w, int tiles_x,
                                                                    int tiles_y) {
  __shared__ float2 patch[kI2cPatch * kI2cPatch];
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int trem = tile - b * tiles_x * tiles_y;
  const int ty0 = (trem / tiles_x) * kI2cTile, tx0 = (trem % tiles_x) * kI2cTile;
  pdl_wait();
  pdl_trigger();
  const float2* cb = coords + (int64_t)b * h * w;
  for (int i = threadIdx.x; i < kI2cPatch * kI2cPatch; i += kI2cThreads) {
    const int py = i / kI2cPatch, pxx = i - py * kI2cPatch;
    const int y = ty0 + py - 3, x = tx0 + pxx - 3;
    float2 v = make_float2(0.f, 0.f);
    if ((unsigned)y < (unsigned)h && (unsigned)x < (unsigned)w) {
      v = cb[y * w + x];
      if (taps) {
        const float2 d = flowhead2_gather(taps, tap_bias, (int64_t)b * h * w + y * w + x, y, x, h, w);
        v.x += d.x;
        v.y += d.y;
      }
      v.x -= (float)x;
      v.y -= (float)y;
    }
    patch[i] = v;
  }
  __syncthreads();
  const int total = kI2cTile * kI2cTile * words_per_px;
  for (int i = threadIdx.x; i < total; i += kI2cThreads) {
    const int px = i / words_per_px, k = i - px * words_per_px;
    const int r = px / kI2cTile, c = px - r * kI2cTile;
    const int y = ty0 + r, x = tx0 + c;
    if (y >= h || x >= w) continue;
    __half2 o = __floats2half2_rn(0.f, 0.f);
    if (k < 98) {
      const int t = k < 49 ? k : k - 49;
      const float2 f = patch[(r + t / 7) * kI2cPatch + c + t % 7];
      const __half hx = __float2half_rn(f.x), hy = __float2half_rn(f.y);
      o = k < 49 ? __halves2half2(hx, hy) : __floats2half2_rn(f.x - __half2float(hx), f.y - __half2float(hy));
    }
    out[((int64_t)b * h * w + (int64_t)y * w + x) * words_per_px + k] = o;
  }
}

__device__ __forceinline__ float reduce_transpose32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(256) flowhead2_taps_h_kernel(const __half* __restrict__ x, const float* __restrict__ w2, float* __restrict__ y,
                                                               int64_t npix) {
  __shared__ __align__(16) float ws[18 * 256];
  for (int i = threadIdx.x; i < 18 * 256 / 4; i += blockDim.x) reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(w2) + i);
  pdl_wait();        // filter staging (the kernel's main fixed cost) overlaps the tail of the flow head's first convolution
  pdl_trigger();
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ngroups = (npix + 1) / 2;
  for (int64_t g = (int64_t)blockIdx.x * 8 + wib; g < ngroups; g += (int64_t)gridDim.x * 8) {
    const int64_t p0 = g * 2;
    float4 xa[2], xb[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t p = p0 + i < npix ? p0 + i : npix - 1;
      const uint4 raw = *(reinterpret_cast<const uint4*>(x + p * 256) + lane);   // 8 halves = channels 8*lane .. 8*lane+7
      const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
      const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&raw.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&raw.w));
      xa[i] = make_float4(f0.x, f0.y, f1.x, f1.y);
      xb[i] = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
    float v[2][32];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 32; ++j) v[q][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      const float4* wp = reinterpret_cast<const float4*>(ws + j * 256) + lane * 2;
      const float4 u0 = wp[0], u1 = wp[1];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = i * 18 + j;
        v[idx >> 5][idx & 31] = xa[i].x * u0.x + xa[i].y * u0.y + xa[i].z * u0.z + xa[i].w * u0.w + xb[i].x * u1.x + xb[i].y * u1.y +
                                xb[i].z * u1.z + xb[i].w * u1.w;
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float tot = reduce_transpose32(v[q], lane);
      const int idx = q * 32 + lane;
      if (idx < 36 && p0 + idx / 18 < npix) y[p0 * 18 + idx] = tot;
    }
  }
}

}  // namespace sdof

extern "C" {

int sdof_motion_tail16_h(const void* mc16, const void* mf16, const float* bias, const float* flow, int64_t npix, void* hx16, int hx16_stride,
                         sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(mc16 && mf16 && bias && flow && hx16, "sdof_motion_tail16_h: NULL pointer");
  SDOF_REQUIRE(hx16_stride >= 256 && hx16_stride % 8 == 0, "sdof_motion_tail16_h: hx16_stride must be >= 256 and a multiple of 8");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(mc16) | reinterpret_cast<uintptr_t>(mf16) | reinterpret_cast<uintptr_t>(hx16)) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(flow) & 7) == 0,
               "sdof_motion_tail16_h: misaligned pointer");
  if (npix <= 0) return SDOF_OK;
  SDOF_CUDA(launch_pdl(motion_tail16_h_kernel, dim3(grid_for(npix * 32, 256, 8)), dim3(256), 0, as_stream(stream),
                       reinterpret_cast<const __half*>(mc16), reinterpret_cast<const __half*>(mf16), reinterpret_cast<const float4*>(bias),
                       reinterpret_cast<const float2*>(flow), reinterpret_cast<__half*>(hx16), hx16_stride, npix));
  SDOF_LAUNCH_CHECK("motion_tail16_h_kernel");
  return SDOF_OK;
}

int sdof_gru_rh_h(const void* zr16, int zr_channels, const float* zrmap, const float* h, void* rh16, int64_t npix, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(zr16 && zrmap && h && rh16, "sdof_gru_rh_h: NULL pointer");
  SDOF_REQUIRE(zr_channels == 256 || zr_channels == 384, "sdof_gru_rh_h: zr must have 256 or 384 channels (hidden = 128)");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(zr16) | reinterpret_cast<uintptr_t>(rh16)) & 7) == 0 &&
                   ((reinterpret_cast<uintptr_t>(zrmap) | reinterpret_cast<uintptr_t>(h)) & 15) == 0, "sdof_gru_rh_h: misaligned pointer");
  if (npix <= 0) return SDOF_OK;
  SDOF_CUDA(launch_pdl(gru_rh_h_kernel, dim3(grid_for(npix * 32, 256, 8)), dim3(256), 0, as_stream(stream),
                       reinterpret_cast<const __half*>(zr16), zr_channels, reinterpret_cast<const float4*>(zrmap),
                       reinterpret_cast<const float4*>(h), reinterpret_cast<__half*>(rh16), npix));
  SDOF_LAUNCH_CHECK("gru_rh_h_kernel");
  return SDOF_OK;
}

int sdof_gru_update_h(const void* zr16, const float* zrmap, const void* q16, const float* qmap, float* h, void* hx16, int hx16_stride, void* h16,
                      int64_t npix, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(zr16 && zrmap && q16 && qmap && h && hx16, "sdof_gru_update_h: NULL pointer");
  SDOF_REQUIRE(hx16_stride >= 128 && hx16_stride % 8 == 0, "sdof_gru_update_h: bad hx16_stride");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(zr16) | reinterpret_cast<uintptr_t>(q16) | reinterpret_cast<uintptr_t>(hx16) |
                 reinterpret_cast<uintptr_t>(h16)) & 7) == 0 &&
                   ((reinterpret_cast<uintptr_t>(zrmap) | reinterpret_cast<uintptr_t>(qmap) | reinterpret_cast<uintptr_t>(h)) & 15) == 0,
               "sdof_gru_update_h: misaligned pointer");
  if (npix <= 0) return SDOF_OK;
  SDOF_CUDA(launch_pdl(gru_update_h_kernel, dim3(grid_for(npix * 32, 256, 8)), dim3(256), 0, as_stream(stream),
                       reinterpret_cast<const __half*>(zr16), 384, reinterpret_cast<const float4*>(zrmap), reinterpret_cast<const __half*>(q16),
                       reinterpret_cast<const float4*>(qmap), reinterpret_cast<float4*>(h), reinterpret_cast<__half*>(hx16), hx16_stride,
                       reinterpret_cast<__half*>(h16), npix));
  SDOF_LAUNCH_CHECK("gru_update_h_kernel");
  return SDOF_OK;
}

static int conv7x7_c2_relu_h_impl(const float* flow, const float* wT, const float* bias, void* out16, int B, int h, int w, const float* coords,
                                  const float* taps, float tap_bias_x, float tap_bias_y, sdof_stream_t stream);

int sdof_conv7x7_c2_relu_h(const float* flow, const float* wT, const float* bias, void* out16, int B, int h, int w, sdof_stream_t stream) {
  SDOF_REQUIRE(flow, "sdof_conv7x7_c2_relu_h: NULL pointer");
  return conv7x7_c2_relu_h_impl(flow, wT, bias, out16, B, h, w, nullptr, nullptr, 0.f, 0.f, stream);
}

int sdof_conv7x7_c2_relu_coords_h(const float* coords, const float* taps, float tap_bias_x, float tap_bias_y, const float* wT, const float* bias,
                                  void* out16, int B, int h, int w, sdof_stream_t stream) {
  SDOF_REQUIRE(coords, "sdof_conv7x7_c2_relu_coords_h: NULL pointer");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(coords) | reinterpret_cast<uintptr_t>(taps)) & 7) == 0, "sdof_conv7x7_c2_relu_coords_h: misaligned pointer");
  return conv7x7_c2_relu_h_impl(coords, wT, bias, out16, B, h, w, coords, taps, tap_bias_x, tap_bias_y, stream);
}

static int conv7x7_c2_relu_h_impl(const float* flow, const float* wT, const float* bias, void* out16, int B, int h, int w, const float* coords,
                                  const float* taps, float tap_bias_x, float tap_bias_y, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(flow && wT && bias && out16, "sdof_conv7x7_c2_relu_h: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h >= 1 && w >= 1, "sdof_conv7x7_c2_relu_h: bad sizes");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(flow) & 7) | (reinterpret_cast<uintptr_t>(wT) & 15) | (reinterpret_cast<uintptr_t>(bias) & 15) |
                (reinterpret_cast<uintptr_t>(out16) & 7)) == 0, "sdof_conv7x7_c2_relu_h: misaligned pointer");
  if (B == 0) return SDOF_OK;
  static bool attr_set[64] = {};
  int dev = 0;
  SDOF_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    SDOF_CUDA(cudaFuncSetAttribute(conv7x7_c2_relu_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kH7Smem));
    attr_set[dev] = true;
  }
  const int tx = ceil_div(w, kH7Tile), ty = ceil_div(h, kH7Tile);
  const int64_t tiles = (int64_t)tx * ty * B;
  SDOF_REQUIRE(tiles < 0x7fffffffLL, "sdof_conv7x7_c2_relu_h: too many tiles");
  SDOF_CUDA(launch_pdl(conv7x7_c2_relu_h_kernel, dim3((unsigned)tiles), dim3(kH7Threads), kH7Smem, as_stream(stream),
                       reinterpret_cast<const float2*>(flow), wT, bias, reinterpret_cast<__half*>(out16), h, w, tx, ty,
                       reinterpret_cast<const float2*>(coords), taps, make_float2(tap_bias_x, tap_bias_y)));
  SDOF_LAUNCH_CHECK("conv7x7_c2_relu_h_kernel");
  return SDOF_OK;
}

int sdof_flow_im2col7_h(const float* coords, const float* taps, float tap_bias_x, float tap_bias_y, void* out16, int out_channels, int B, int h,
                        int w, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(coords && out16, "sdof_flow_im2col7_h: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h >= 1 && w >= 1, "sdof_flow_im2col7_h: bad sizes");
  SDOF_REQUIRE(out_channels >= 196 && out_channels <= 512 && out_channels % 8 == 0, "sdof_flow_im2col7_h: out_channels must be a multiple of 8 in [196, 512]");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(coords) | reinterpret_cast<uintptr_t>(taps)) & 7) == 0 && (reinterpret_cast<uintptr_t>(out16) & 3) == 0,
               "sdof_flow_im2col7_h: misaligned pointer");
  if (B == 0) return SDOF_OK;
  const int tx = ceil_div(w, kI2cTile), ty = ceil_div(h, kI2cTile);
  const int64_t tiles = (int64_t)tx * ty * B;
  SDOF_REQUIRE(tiles < 0x7fffffffLL, "sdof_flow_im2col7_h: too many tiles");
  SDOF_CUDA(launch_pdl(flow_im2col7_h_kernel, dim3((unsigned)tiles), dim3(kI2cThreads), 0, as_stream(stream), reinterpret_cast<const float2*>(coords),
                       taps, make_float2(tap_bias_x, tap_bias_y), reinterpret_cast<__half2*>(out16), out_channels / 2, h, w, tx, ty));
  SDOF_LAUNCH_CHECK("flow_im2col7_h_kernel");
  return SDOF_OK;
}

int sdof_flowhead2_taps_h(const void* x16, const float* w2, int64_t npix, float* scratch, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(x16 && w2 && scratch, "sdof_flowhead2_taps_h: NULL pointer");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(x16) | reinterpret_cast<uintptr_t>(w2)) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 7) == 0,
               "sdof_flowhead2_taps_h: misaligned pointer");
  if (npix <= 0) return SDOF_OK;
  const int64_t want = ceil_div64(ceil_div64(npix, 2), 8);
  int cap_mul = 1;             // CTAs per SM (each stages the 18 KB filter once); SDOF_FH_CAP_H overrides for experiments
  if (const char* e = getenv("SDOF_FH_CAP_H")) {
    cap_mul = atoi(e);
    if (cap_mul < 1 || cap_mul > 8) cap_mul = 1;
  }
  const int64_t cap = (int64_t)sm_count() * cap_mul;
  SDOF_CUDA(launch_pdl(flowhead2_taps_h_kernel, dim3((unsigned)(want < cap ? want : cap)), dim3(256), 0, as_stream(stream),
                       reinterpret_cast<const __half*>(x16), w2, scratch, npix));
  SDOF_LAUNCH_CHECK("flowhead2_taps_h_kernel");
  return SDOF_OK;
}

}  // extern "C"
