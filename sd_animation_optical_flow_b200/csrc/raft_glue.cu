#include <stdlib.h>
// Per-iteration glue of the RAFT update loop on dense channels-last (NHWC) buffers.
//
// The reference's update block (RAFT/core/update.py:79-136) runs ~100 small kernels per GRU iteration
// in eager PyTorch (cat, relu, sigmoid, tanh, mul, add, plus two layout transposes around every cuDNN
// convolution).  With the activations kept in NHWC and the `torch.cat`s replaced by persistent
// concatenated buffers, everything between the convolutions collapses into the four element-wise
// kernels below; the convolutions themselves stay in cuDNN (raft_fast.py).
//
//   relu_scatter   : dst[:, off:off+C] = relu(src (+ src2) + bias)   (optionally into two buffers)
//   gru_rh         : rhx[:, 0:Hd] = sigmoid(zr[:, Hd:2Hd]) * h
//   gru_update     : h' = (1 - sigmoid(z)) * h + sigmoid(z) * tanh(q)   -> h (dense) and hx[:, 0:Hd]
//   flow_update    : coords1 += delta; flow = coords1 - grid  -> dense flow, hx / rhx flow slots
//   convex_upsample: RAFT.upsample_flow (RAFT/core/raft.py:72-83): softmax over the 9 taps of the 8x8 masks,
//                    weighted sum of the 3x3 neighbourhood of 8*flow  -> full-resolution [B,8h,8w,2]
// All are HBM/L2-bound element-wise work: float4 vectorised, grid-stride.
#include "sdof_common.cuh"

#include <cuda_fp16.h>
namespace sdof {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256) relu_scatter_kernel(const float4* __restrict__ src, const float4* __restrict__ src2,
                                                           const float4* __restrict__ bias, int64_t npix, int C4, float* __restrict__ d1, int d1_stride, int d1_off,
                                                           float* __restrict__ d2, int d2_stride, int d2_off, int C_valid) {
  const int64_t total = npix * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C4;
    const int c = (int)(i - p * C4) * 4;
    float4 v = src[i];
    if (src2) {  // a convolution split over two input-channel groups (computed on two streams): sum of the partial results
      const float4 u = src2[i];
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (bias) {
      const float4 bv = __ldg(bias + (c >> 2));
      v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
    }
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    float* o1 = d1 + p * d1_stride + d1_off + c;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c + k < C_valid) o1[k] = vv[k];
    if (d2) {
      float* o2 = d2 + p * d2_stride + d2_off + c;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c + k < C_valid) o2[k] = vv[k];
    }
  }
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// bias_zr / bias_q: per-channel vectors (bias_map = 0) or per-pixel maps [npix][2*Hd] / [npix][Hd] (bias_map = 1): the
// maps carry the convolution bias PLUS the contribution of the context features `inp`, which is the same in every
// iteration and is therefore convolved once per pair instead of 20 x 6 times (raft_fast.py).
__global__ void __launch_bounds__(256) gru_rh_kernel(const float4* __restrict__ zr, const float4* __restrict__ bias_zr,
                                                     const float4* __restrict__ h, float* __restrict__ rhx, int64_t npix,
                                                     int Hd4, int rhx_stride, int bias_map, int zr_stride4) {
  const int64_t total = npix * Hd4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / Hd4;
    const int c4 = (int)(i - p * Hd4);
    float4 r = zr[p * zr_stride4 + Hd4 + c4];
    if (bias_zr) r = add4(r, __ldg(bias_zr + (bias_map ? p * (2 * Hd4) : 0) + Hd4 + c4));
    const float4 hv = h[i];
    float4 o;
    o.x = sigmoidf_(r.x) * hv.x; o.y = sigmoidf_(r.y) * hv.y; o.z = sigmoidf_(r.z) * hv.z; o.w = sigmoidf_(r.w) * hv.w;
    *reinterpret_cast<float4*>(rhx + p * rhx_stride + 4 * c4) = o;
  }
}

__global__ void __launch_bounds__(256) gru_update_kernel(const float4* __restrict__ zr, const float4* __restrict__ bias_zr,
                                                         const float4* __restrict__ q, const float4* __restrict__ bias_q,
                                                         float4* __restrict__ h, float* __restrict__ hx, int64_t npix,
                                                         int Hd4, int hx_stride, int bias_map, int zr_stride4, int q_extra) {
  const int64_t total = npix * Hd4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / Hd4;
    const int c4 = (int)(i - p * Hd4);
    float4 z = zr[p * zr_stride4 + c4];
    float4 qv = q[i];
    if (q_extra) qv = add4(qv, zr[p * zr_stride4 + 2 * Hd4 + c4]);  // share of convq computed with the z|r convolution
    if (bias_zr) z = add4(z, __ldg(bias_zr + (bias_map ? p * (2 * Hd4) : 0) + c4));
    if (bias_q) qv = add4(qv, __ldg(bias_q + (bias_map ? p * Hd4 : 0) + c4));
    const float4 hv = h[i];
    float4 o;
    float s;
    s = sigmoidf_(z.x); o.x = (1.f - s) * hv.x + s * tanhf(qv.x);
    s = sigmoidf_(z.y); o.y = (1.f - s) * hv.y + s * tanhf(qv.y);
    s = sigmoidf_(z.z); o.z = (1.f - s) * hv.z + s * tanhf(qv.z);
    s = sigmoidf_(z.w); o.w = (1.f - s) * hv.w + s * tanhf(qv.w);
    h[i] = o;
    *reinterpret_cast<float4*>(hx + p * hx_stride + 4 * c4) = o;
  }
}

// delta [npix,2] (flow head output), coords1 [npix,2] in/out; flow = coords1 - (x, y) grid
__global__ void __launch_bounds__(256) flow_update_kernel(const float2* __restrict__ delta, float2 delta_bias,
                                                          float2* __restrict__ coords1,
                                                          float2* __restrict__ flow, float* __restrict__ hx, int hx_stride,
                                                          int hx_off, float* __restrict__ rhx, int rhx_stride, int rhx_off,
                                                          int64_t npix, int h, int w) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int rem = (int)(p % ((int64_t)h * w));
    const int y = rem / w, x = rem - y * w;
    float2 c = coords1[p];
    if (delta) {
      const float2 d = delta[p];
      c.x += d.x + delta_bias.x;
      c.y += d.y + delta_bias.y;
      coords1[p] = c;
    }
    const float2 f = make_float2(c.x - (float)x, c.y - (float)y);
    flow[p] = f;
    if (hx) *reinterpret_cast<float2*>(hx + p * hx_stride + hx_off) = f;
    if (rhx) *reinterpret_cast<float2*>(rhx + p * rhx_stride + rhx_off) = f;
  }
}

// mask [B,h,w,576] NHWC with channel = k*64 + i*8 + j (k: 3x3 tap row-major, (i,j): sub-pixel) exactly as
// mask.view(N,1,9,8,8,H,W) indexes the reference's NCHW tensor; mask_scale = 0.25 (update.py:135).
// One thread per (low-res pixel, sub-pixel): softmax over the 9 taps, sum of weights * 8*flow of the 3x3
// neighbourhood (zero padding, F.unfold(padding=1)).  Output up [B, 8h, 8w, 2].
__global__ void __launch_bounds__(256) convex_upsample_kernel(const float* __restrict__ mask, const float* __restrict__ mask_bias,
                                                              float mask_scale,
                                                              const float2* __restrict__ flow, int B, int h, int w,
                                                              float2* __restrict__ up) {
  const int64_t total = (int64_t)B * h * w * 64;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int sub = (int)(i & 63);
    const int64_t p = i >> 6;
    const int b = (int)(p / ((int64_t)h * w));
    const int rem = (int)(p - (int64_t)b * h * w);
    const int y = rem / w, x = rem - y * w;
    const float* m = mask + p * 576 + sub;
    float lg[9];
    float mx = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      lg[k] = (m[k * 64] + (mask_bias ? __ldg(mask_bias + k * 64 + sub) : 0.f)) * mask_scale;
      mx = fmaxf(mx, lg[k]);
    }
    float s = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float e = __expf(lg[k] - mx);
      s += e;
      const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
      if ((unsigned)yy < (unsigned)h && (unsigned)xx < (unsigned)w) {
        const float2 f = flow[((int64_t)b * h + yy) * w + xx];
        ax += e * f.x;
        ay += e * f.y;
      }
    }
    const float inv = 8.0f / s;
    const int si = sub >> 3, sj = sub & 7;
    up[((int64_t)b * 8 * h + 8 * y + si) * (8 * w) + 8 * x + sj] = make_float2(ax * inv, ay * inv);
  }
}

// InstanceNorm2d (no affine, biased variance, eps) + optional ReLU on NCHW planes, one CTA per (n, c) plane:
// mean, then variance around the mean, then normalise -- three sweeps over a plane that stays in L2.  Replaces
// the batch_norm_collect_statistics / calc_invstd / transform_input / clamp kernels eager PyTorch runs for
// `relu(norm(conv(x)))` in the feature encoder (RAFT/core/extractor.py:49-50, 172-173).
__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(1024) instnorm_relu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t hw,
                                                             float eps, int relu) {
  __shared__ float red[32];
  const float* xp = x + (int64_t)blockIdx.x * hw;
  float* yp = y + (int64_t)blockIdx.x * hw;
  const bool vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(yp)) & 15) == 0;
  const int64_t n4 = vec ? hw / 4 : 0;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(xp)[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
  for (int64_t i = 4 * n4 + threadIdx.x; i < hw; i += blockDim.x) s += xp[i];
  const float mean = block_sum(s, red) / (float)hw;
  float q = 0.f;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(xp)[i];
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  for (int64_t i = 4 * n4 + threadIdx.x; i < hw; i += blockDim.x) {
    const float a = xp[i] - mean;
    q += a * a;
  }
  const float inv = rsqrtf(block_sum(q, red) / (float)hw + eps);
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(xp)[i];
    v.x = (v.x - mean) * inv; v.y = (v.y - mean) * inv; v.z = (v.z - mean) * inv; v.w = (v.w - mean) * inv;
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    reinterpret_cast<float4*>(yp)[i] = v;
  }
  for (int64_t i = 4 * n4 + threadIdx.x; i < hw; i += blockDim.x) {
    float v = (xp[i] - mean) * inv;
    yp[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

}  // namespace sdof

extern "C" {

int sdof_instnorm_relu_nchw(const float* x, float* y, int64_t planes, int64_t hw, float eps, int relu, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(x && y, "sdof_instnorm_relu_nchw: NULL pointer");
  SDOF_REQUIRE(planes >= 0 && planes <= 0x7fffffff && hw >= 1, "sdof_instnorm_relu_nchw: bad sizes");
  if (planes == 0) return SDOF_OK;
  instnorm_relu_kernel<<<(unsigned)planes, 1024, 0, as_stream(stream)>>>(x, y, hw, eps, relu);
  SDOF_LAUNCH_CHECK("instnorm_relu_kernel");
  return SDOF_OK;
}


int sdof_relu_scatter(const float* src, const float* src2, const float* bias, int64_t npix, int C, float* dst1, int dst1_stride, int dst1_off, float* dst2,
                      int dst2_stride, int dst2_off, int C_valid, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(src && dst1, "sdof_relu_scatter: NULL pointer");
  SDOF_REQUIRE(C > 0 && C % 4 == 0 && C_valid > 0 && C_valid <= C, "sdof_relu_scatter: C must be a multiple of 4, 0 < C_valid <= C");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(src2)) & 15) == 0, "sdof_relu_scatter: src, src2 must be 16-byte aligned");
  if (npix <= 0) return SDOF_OK;
  relu_scatter_kernel<<<grid_for(npix * (C / 4), 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<const float4*>(src2), reinterpret_cast<const float4*>(bias), npix, C / 4, dst1, dst1_stride, dst1_off, dst2,
      dst2_stride, dst2_off, C_valid);
  SDOF_LAUNCH_CHECK("relu_scatter_kernel");
  return SDOF_OK;
}

int sdof_gru_rh(const float* zr, const float* bias_zr, const float* h, float* rhx, int64_t npix, int hidden, int rhx_stride,
                int bias_map, int zr_channels, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(zr && h && rhx, "sdof_gru_rh: NULL pointer");
  SDOF_REQUIRE(hidden > 0 && hidden % 4 == 0 && rhx_stride % 4 == 0, "sdof_gru_rh: hidden and stride must be multiples of 4");
  SDOF_REQUIRE(zr_channels == 2 * hidden || zr_channels == 3 * hidden, "sdof_gru_rh: zr must have 2*hidden or 3*hidden channels");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(zr) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(rhx)) & 15) == 0,
               "sdof_gru_rh: pointers must be 16-byte aligned");
  if (npix <= 0) return SDOF_OK;
  gru_rh_kernel<<<grid_for(npix * (hidden / 4), 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(zr), reinterpret_cast<const float4*>(bias_zr), reinterpret_cast<const float4*>(h), rhx, npix,
      hidden / 4, rhx_stride, bias_map, zr_channels / 4);
  SDOF_LAUNCH_CHECK("gru_rh_kernel");
  return SDOF_OK;
}

int sdof_gru_update(const float* zr, const float* bias_zr, const float* q, const float* bias_q, float* h, float* hx,
                    int64_t npix, int hidden, int hx_stride, int bias_map, int zr_channels, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(zr && q && h && hx, "sdof_gru_update: NULL pointer");
  SDOF_REQUIRE(hidden > 0 && hidden % 4 == 0 && hx_stride % 4 == 0, "sdof_gru_update: hidden and stride must be multiples of 4");
  SDOF_REQUIRE(zr_channels == 2 * hidden || zr_channels == 3 * hidden, "sdof_gru_update: zr must have 2*hidden or 3*hidden channels");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(zr) | reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(h) |
                 reinterpret_cast<uintptr_t>(hx)) & 15) == 0, "sdof_gru_update: pointers must be 16-byte aligned");
  if (npix <= 0) return SDOF_OK;
  gru_update_kernel<<<grid_for(npix * (hidden / 4), 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(zr), reinterpret_cast<const float4*>(bias_zr), reinterpret_cast<const float4*>(q),
      reinterpret_cast<const float4*>(bias_q), reinterpret_cast<float4*>(h), hx, npix, hidden / 4, hx_stride, bias_map, zr_channels / 4,
      zr_channels == 3 * hidden);
  SDOF_LAUNCH_CHECK("gru_update_kernel");
  return SDOF_OK;
}

int sdof_flow_update(const float* delta, float delta_bias_x, float delta_bias_y, float* coords1, float* flow, float* hx, int hx_stride, int hx_off, float* rhx,
                     int rhx_stride, int rhx_off, int B, int h, int w, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(coords1 && flow, "sdof_flow_update: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h >= 1 && w >= 1, "sdof_flow_update: bad sizes");
  SDOF_REQUIRE((hx_stride % 2 == 0) && (hx_off % 2 == 0) && (rhx_stride % 2 == 0) && (rhx_off % 2 == 0),
               "sdof_flow_update: strides/offsets must be even (float2 stores)");
  const int64_t npix = (int64_t)B * h * w;
  if (npix == 0) return SDOF_OK;
  flow_update_kernel<<<grid_for(npix, 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2*>(delta), make_float2(delta_bias_x, delta_bias_y), reinterpret_cast<float2*>(coords1),
      reinterpret_cast<float2*>(flow), hx, hx_stride,
      hx_off, rhx, rhx_stride, rhx_off, npix, h, w);
  SDOF_LAUNCH_CHECK("flow_update_kernel");
  return SDOF_OK;
}

int sdof_convex_upsample(const float* mask, const float* mask_bias, float mask_scale, const float* flow, int B, int h, int w, float* up,
                         sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(mask && flow && up, "sdof_convex_upsample: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h >= 1 && w >= 1, "sdof_convex_upsample: bad sizes");
  const int64_t total = (int64_t)B * h * w * 64;
  if (total == 0) return SDOF_OK;
  convex_upsample_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(
      mask, mask_bias, mask_scale, reinterpret_cast<const float2*>(flow), B, h, w, reinterpret_cast<float2*>(up));
  SDOF_LAUNCH_CHECK("convex_upsample_kernel");
  return SDOF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Channels-last feature encoders (RAFT/core/extractor.py:118-192).  Eager PyTorch / cuDNN wrap every NCHW
// convolution of the encoders in an NCHW->NHWC and an NHWC->NCHW transpose (round-1 launch list: 0.86 ms of a
// 5.9 ms step); with the activations kept NHWC the tensor-core convolutions run directly, and the
// normalisation between them is the three kernels below.
//
//   instnorm_stats_nhwc : per (image, channel) sum and sum of squares over H*W, fp32 partials per thread,
//                         fp64 atomics into stats[N][C][2] (zeroed by the caller)
//   instnorm_apply_nhwc : y = relu?((x - mean) * rsqrt(var + eps)), optionally followed by the residual
//                         y = relu(res + y) of the ResidualBlock (extractor.py:49-58)
//   add_relu            : y = relu(a + b)   (cnet, whose BatchNorm is folded into the convolutions)
namespace sdof {

constexpr int kInThreads = 256;

// channel quads of an fp32 (float4) or fp16 (uint2 = 4 halves) channels-last tensor
__device__ __forceinline__ float4 ldq(const float4* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float4 ldq(const uint2* p, int64_t i) {
  const uint2 v = p[i];
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void stq(float4* p, int64_t i, float4 v) { p[i] = v; }
__device__ __forceinline__ void stq(uint2* p, int64_t i, float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&a);
  o.y = *reinterpret_cast<const uint32_t*>(&b);
  p[i] = o;
}

// Round 2: the statistics pass was 18-21 us per layer-1 tensor (25-50 MB): four 8/16-byte loads in flight per thread at one
// CTA-pair per SM leave the memory system idle, and 592 CTAs x 16 threads x 8 fp64 atomics pile ~300 serialised atomics on
// each of the 256 addresses.  Now: 16-byte loads (8 halves / 4 floats), eight of them in flight per thread, two 512-thread
// CTAs per SM for the whole tensor (64 KB in flight per SM), and one atomic pair per channel per CTA.
constexpr int kStThreads = 512;
template <typename ST> struct StatLoad;
template <> struct StatLoad<float4> {      // one 16-byte load = 4 channels
  static constexpr int kCh = 4;
  static __device__ __forceinline__ void ld(const float4* p, int64_t i, float (&v)[8]) {
    const float4 a = p[i];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
};
template <> struct StatLoad<uint2> {       // one 16-byte load = 8 channels (two quads)
  static constexpr int kCh = 8;
  static __device__ __forceinline__ void ld(const uint2* p, int64_t i, float (&v)[8]) {
    const uint4 a = reinterpret_cast<const uint4*>(p)[i];
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&a.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&a.w));
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
  }
};

template <typename ST>
__global__ void __launch_bounds__(kStThreads, 2) instnorm_stats_nhwc_kernel(const ST* __restrict__ x, double* __restrict__ stats,
                                                                        int64_t hw, int C, int px_per_cta) {
  constexpr int kCh = StatLoad<ST>::kCh;
  extern __shared__ float st_red[];            // [2][kStThreads][kCh]
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y;
  const int G = C / kCh;                        // 16-byte groups per pixel
  const int npl = kStThreads / G;               // pixel lanes
  const int T = npl * G;                        // active threads (a multiple of G: a thread keeps its channel group)
  const int64_t p0 = (int64_t)blockIdx.x * px_per_cta;
  const int64_t p1 = p0 + px_per_cta < hw ? p0 + px_per_cta : hw;
  const ST* xs = x + ((int64_t)n * hw + p0) * (C / 4);   // ST counts channel quads
  const int64_t cnt = (p1 - p0) * G;
  float s[kCh], q[kCh];
#pragma unroll
  for (int k = 0; k < kCh; ++k) s[k] = q[k] = 0.f;
  if ((int)threadIdx.x < T) {
    constexpr int kU = 32 / kCh;   // loads in flight per thread: 32 values = 64 (fp16) / 128 (fp32) bytes; two CTAs per SM fit 64 registers
    int64_t i = threadIdx.x;
    for (; i + (kU - 1) * (int64_t)T < cnt; i += kU * (int64_t)T) {
      float v[kU][8];
#pragma unroll
      for (int u = 0; u < kU; ++u) StatLoad<ST>::ld(xs, i + u * (int64_t)T, v[u]);   // default caching: the apply kernel re-reads x from L2
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        float ps = 0.f, pq = 0.f;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          ps += v[u][k];
          pq += v[u][k] * v[u][k];
        }
        s[k] += ps;
        q[k] += pq;
      }
    }
    for (; i < cnt; i += T) {
      float v[8];
      StatLoad<ST>::ld(xs, i, v);
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        s[k] += v[k];
        q[k] += v[k] * v[k];
      }
    }
  }
  float* rs = st_red + (size_t)threadIdx.x * kCh;
  float* rq = st_red + (size_t)(kStThreads + threadIdx.x) * kCh;
#pragma unroll
  for (int k = 0; k < kCh; ++k) {
    rs[k] = s[k];
    rq[k] = q[k];
  }
  __syncthreads();
  // thread c < C sums channel c over the pixel lanes (fp32 partial of this CTA, fp64 across CTAs)
  for (int c = threadIdx.x; c < C; c += kStThreads) {
    const int g = c / kCh, k = c - g * kCh;
    float ts = 0.f, tq = 0.f;
    for (int l = 0; l < npl; ++l) {
      ts += st_red[(size_t)(l * G + g) * kCh + k];
      tq += st_red[(size_t)(kStThreads + l * G + g) * kCh + k];
    }
    double* o = stats + ((int64_t)n * C + c) * 2;
    atomicAdd(o, (double)ts);
    atomicAdd(o + 1, (double)tq);
  }
}

template <typename ST>
__global__ void __launch_bounds__(kInThreads) instnorm_apply_nhwc_kernel(const ST* __restrict__ x, const double* __restrict__ stats,
                                                                        const ST* __restrict__ res, ST* __restrict__ y,
                                                                        int64_t hw, int C4, int px_per_cta, float eps, int relu) {
  __shared__ float mean_s[512], rstd_s[512];
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y;
  const int C = C4 * 4;
  for (int c = threadIdx.x; c < C; c += kInThreads) {
    const double s = stats[((int64_t)n * C + c) * 2], q = stats[((int64_t)n * C + c) * 2 + 1];
    const double m = s / (double)hw;
    double var = q / (double)hw - m * m;
    if (var < 0.0) var = 0.0;
    mean_s[c] = (float)m;
    rstd_s[c] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int npl = kInThreads / C4;
  const int T = npl * C4;
  if ((int)threadIdx.x >= T) return;
  const int cq = threadIdx.x % C4;
  const float4 m = *reinterpret_cast<const float4*>(mean_s + 4 * cq);
  const float4 r = *reinterpret_cast<const float4*>(rstd_s + 4 * cq);
  const int64_t p0 = (int64_t)blockIdx.x * px_per_cta;
  const int64_t p1 = p0 + px_per_cta < hw ? p0 + px_per_cta : hw;
  const int64_t base = ((int64_t)n * hw + p0) * C4;
  const int64_t cnt = (p1 - p0) * C4;
  for (int64_t i = threadIdx.x; i < cnt; i += T) {
    const float4 v = ldq(x, base + i);
    float4 o;
    o.x = (v.x - m.x) * r.x; o.y = (v.y - m.y) * r.y; o.z = (v.z - m.z) * r.z; o.w = (v.w - m.w) * r.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (res) {
      const float4 a = ldq(res, base + i);
      o.x = fmaxf(a.x + o.x, 0.f); o.y = fmaxf(a.y + o.y, 0.f); o.z = fmaxf(a.z + o.z, 0.f); o.w = fmaxf(a.w + o.w, 0.f);
    }
    stq(y, base + i, o);
  }
}

__global__ void __launch_bounds__(256) add_relu_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y,
                                                       int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 u = a[i], v = b[i];
    y[i] = make_float4(fmaxf(u.x + v.x, 0.f), fmaxf(u.y + v.y, 0.f), fmaxf(u.z + v.z, 0.f), fmaxf(u.w + v.w, 0.f));
  }
}

static int instnorm_px_per_cta(int N, int64_t hw) {
  // ~4 CTAs per SM over the whole tensor, at least 64 pixels each
  int64_t want = ((int64_t)N * hw) / ((int64_t)sm_count() * 4);
  if (want < 64) want = 64;
  if (want > 4096) want = 4096;
  return (int)want;
}

}  // namespace sdof

extern "C" {

static int instnorm_stats_impl(const void* x, int elem_bytes, int N, int64_t hw, int C, double* stats, sdof_stream_t stream);
static int instnorm_apply_impl(const void* x, int elem_bytes, const double* stats, const void* residual, void* y, int N, int64_t hw, int C,
                               float eps, int relu, sdof_stream_t stream);

int sdof_instnorm_stats_nhwc(const float* x, int N, int64_t hw, int C, double* stats, sdof_stream_t stream) {
  return instnorm_stats_impl(x, 4, N, hw, C, stats, stream);
}
int sdof_instnorm_stats_nhwc_h(const void* x, int N, int64_t hw, int C, double* stats, sdof_stream_t stream) {
  return instnorm_stats_impl(x, 2, N, hw, C, stats, stream);
}
int sdof_instnorm_apply_nhwc(const float* x, const double* stats, const float* residual, float* y, int N, int64_t hw, int C, float eps,
                             int relu, sdof_stream_t stream) {
  return instnorm_apply_impl(x, 4, stats, residual, y, N, hw, C, eps, relu, stream);
}
int sdof_instnorm_apply_nhwc_h(const void* x, const double* stats, const void* residual, void* y, int N, int64_t hw, int C, float eps,
                               int relu, sdof_stream_t stream) {
  return instnorm_apply_impl(x, 2, stats, residual, y, N, hw, C, eps, relu, stream);
}

static int instnorm_stats_impl(const void* x, int elem_bytes, int N, int64_t hw, int C, double* stats, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(x && stats, "sdof_instnorm_stats_nhwc: NULL pointer");
  SDOF_REQUIRE(N >= 0 && N <= 65535 && hw >= 1 && C >= 4 && C % 4 == 0 && C <= 512, "sdof_instnorm_stats_nhwc: need C %% 4 == 0, 4 <= C <= 512");
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sdof_instnorm_stats_nhwc: x must be 16-byte aligned");
  if (N == 0) return SDOF_OK;
  {
    // the whole tensor over two CTAs per SM (at least 32 pixels each)
    int64_t want = ceil_div64((int64_t)N * hw, (int64_t)sm_count() * 2);
    if (want < 32) want = 32;
    const int sppc = (int)(want > (1 << 20) ? (1 << 20) : want);
    dim3 sgrid((unsigned)ceil_div64(hw, sppc), N);
    if (elem_bytes == 2) {
      SDOF_REQUIRE(C % 8 == 0, "sdof_instnorm_stats_nhwc_h: C must be a multiple of 8");
      SDOF_CUDA(launch_pdl(instnorm_stats_nhwc_kernel<uint2>, sgrid, dim3(kStThreads), 2 * kStThreads * 8 * sizeof(float), as_stream(stream),
                           reinterpret_cast<const uint2*>(x), stats, hw, C, sppc));
    } else {
      SDOF_CUDA(launch_pdl(instnorm_stats_nhwc_kernel<float4>, sgrid, dim3(kStThreads), 2 * kStThreads * 4 * sizeof(float), as_stream(stream),
                           reinterpret_cast<const float4*>(x), stats, hw, C, sppc));
    }
  }
  SDOF_LAUNCH_CHECK("instnorm_stats_nhwc_kernel");
  return SDOF_OK;
}

static int instnorm_apply_impl(const void* x, int elem_bytes, const double* stats, const void* residual, void* y, int N, int64_t hw, int C,
                               float eps, int relu, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(x && stats && y, "sdof_instnorm_apply_nhwc: NULL pointer");
  SDOF_REQUIRE(N >= 0 && N <= 65535 && hw >= 1 && C >= 4 && C % 4 == 0 && C <= 512, "sdof_instnorm_apply_nhwc: need C %% 4 == 0, 4 <= C <= 512");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0,
               "sdof_instnorm_apply_nhwc: pointers must be 16-byte aligned");
  if (N == 0) return SDOF_OK;
  const int ppc = instnorm_px_per_cta(N, hw);
  dim3 grid((unsigned)ceil_div64(hw, ppc), N);
  if (elem_bytes == 2)
    SDOF_CUDA(launch_pdl(instnorm_apply_nhwc_kernel<uint2>, grid, dim3(kInThreads), 0, as_stream(stream), reinterpret_cast<const uint2*>(x),
                         stats, reinterpret_cast<const uint2*>(residual), reinterpret_cast<uint2*>(y), hw, C / 4, ppc, eps, relu));
  else
    SDOF_CUDA(launch_pdl(instnorm_apply_nhwc_kernel<float4>, grid, dim3(kInThreads), 0, as_stream(stream), reinterpret_cast<const float4*>(x),
                         stats, reinterpret_cast<const float4*>(residual), reinterpret_cast<float4*>(y), hw, C / 4, ppc, eps, relu));
  SDOF_LAUNCH_CHECK("instnorm_apply_nhwc_kernel");
  return SDOF_OK;
}

int sdof_add_relu(const float* a, const float* b, float* y, int64_t n, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(a && b && y, "sdof_add_relu: NULL pointer");
  SDOF_REQUIRE(n >= 0 && n % 4 == 0, "sdof_add_relu: n must be a multiple of 4");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
               "sdof_add_relu: pointers must be 16-byte aligned");
  if (n == 0) return SDOF_OK;
  add_relu_kernel<<<grid_for(n / 4, 256, 8), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                                         reinterpret_cast<float4*>(y), n / 4);
  SDOF_LAUNCH_CHECK("add_relu_kernel");
  return SDOF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// The two convolutions of the update block that cuDNN serves badly (round-1 launch list, per GRU iteration):
//   convf1  : 7x7, 2 -> 128 channels on the flow (update.py:85,93): cuDNN falls back to a non-tensor-core
//             engine, 14 us for 0.15 GFLOP;
//   fh.conv2: 3x3, 256 -> 2 channels (update.py:10,14): 13.7 us + two cuDNN padding kernels (6 us) for 0.06 GFLOP,
//             followed by the flow_update kernel.
// Both are tiny and fp32 CUDA-core work here (closer to the fp32 reference than cuDNN's TF32):
//   conv7x7_c2_relu_kernel     : out = relu(conv(flow) + bias), NHWC, register tile 4 px x 8 channels per thread
//   flowhead2_{taps,gather_update}_kernel : delta = conv3x3(x) + bias as per-pixel tap products followed by a
//                                9-neighbour gather fused with the flow_update step (see below)
namespace sdof {

constexpr int kC7Tile = 8;                 // 8x8 output pixels per CTA
constexpr int kC7Threads = 256;            // 16 pixel groups (row, half) x 16 channel groups of 8
constexpr int kC7Patch = kC7Tile + 6;      // 14
constexpr int kC7K = 98;                   // 7*7*2
constexpr size_t kC7Smem = (size_t)kC7K * 128 * 4 + (size_t)kC7Patch * kC7Patch * 8;

__global__ void __launch_bounds__(kC7Threads) conv7x7_c2_relu_kernel(const float2* __restrict__ flow, const float* __restrict__ wT,
                                                                    const float* __restrict__ bias, float* __restrict__ out,
                                                                    int h, int w, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) unsigned char c7_smem[];
  float* ws = reinterpret_cast<float*>(c7_smem);                                   // [98][128]
  float2* patch = reinterpret_cast<float2*>(c7_smem + (size_t)kC7K * 128 * 4);     // [14][14]
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int trem = tile - b * tiles_x * tiles_y;
  const int ty0 = (trem / tiles_x) * kC7Tile, tx0 = (trem % tiles_x) * kC7Tile;
  for (int i = threadIdx.x; i < kC7K * 128 / 4; i += kC7Threads)
    reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(wT) + i);
  const float2* fb = flow + (int64_t)b * h * w;
  for (int i = threadIdx.x; i < kC7Patch * kC7Patch; i += kC7Threads) {
    const int py = i / kC7Patch, pxx = i - py * kC7Patch;
    const int y = ty0 + py - 3, x = tx0 + pxx - 3;
    patch[i] = ((unsigned)y < (unsigned)h && (unsigned)x < (unsigned)w) ? fb[y * w + x] : make_float2(0.f, 0.f);
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
    float in[20];  // 10 pixels x 2 channels of patch row r+ky, columns c0 .. c0+9
    const float4* prow = reinterpret_cast<const float4*>(patch + (r + ky) * kC7Patch + c0);
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
    float4* o = reinterpret_cast<float4*>(out + (((int64_t)b * h + y) * w + x) * 128 + cg * 8);
    o[0] = make_float4(fmaxf(acc[i][0] + b0.x, 0.f), fmaxf(acc[i][1] + b0.y, 0.f), fmaxf(acc[i][2] + b0.z, 0.f), fmaxf(acc[i][3] + b0.w, 0.f));
    o[1] = make_float4(fmaxf(acc[i][4] + b1.x, 0.f), fmaxf(acc[i][5] + b1.y, 0.f), fmaxf(acc[i][6] + b1.z, 0.f), fmaxf(acc[i][7] + b1.w, 0.f));
  }
}

// Transposing butterfly: every lane holds 32 partial values v[0..31]; after 31 shuffles lane L holds the sum over
// all lanes of v[L].
__device__ __forceinline__ float warp_reduce_transpose32(float (&v)[32], int lane) {
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

// FlowHead.conv2 as "taps first": a 3x3 convolution with 2 output channels reads each 1 KB input pixel nine
// times when done output-stationary (55 MB of L2 traffic at 96x64, 16 us measured).  Instead
//   taps kernel  : y[p][tap][co] = <x[p, :], w[tap][co][:]>   (18 dot products per pixel, x read ONCE;
//                  a warp owns 2 pixels, lanes split the 256 channels, weights from shared memory are reused
//                  across the pixels, two transposing butterflies reduce the 36 partials)
//   gather kernel: delta[p][co] = bias[co] + sum_tap y[p + off(tap)][tap][co] (zero padding), then the
//                  coords / flow update of RAFT.forward (raft.py:128-131).

template <int kFhPx>   // pixels per warp
__global__ void __launch_bounds__(256) flowhead2_taps_kernel(const float* __restrict__ x, const float* __restrict__ w2,
                                                             float* __restrict__ y, int64_t npix) {
  constexpr int kFhGroups = (kFhPx * 18 + 31) / 32;  // butterflies per warp
  __shared__ __align__(16) float ws[18 * 256];
  for (int i = threadIdx.x; i < 18 * 256 / 4; i += blockDim.x) reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(w2) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ngroups = (npix + kFhPx - 1) / kFhPx;
  for (int64_t g = (int64_t)blockIdx.x * 8 + wib; g < ngroups; g += (int64_t)gridDim.x * 8) {
    const int64_t p0 = g * kFhPx;
    float4 xa[kFhPx], xb[kFhPx];
#pragma unroll
    for (int i = 0; i < kFhPx; ++i) {
      const int64_t p = p0 + i < npix ? p0 + i : npix - 1;
      const float4* xp = reinterpret_cast<const float4*>(x + p * 256) + lane * 2;
      xa[i] = xp[0];
      xb[i] = xp[1];
    }
    float v[kFhGroups][32];
#pragma unroll
    for (int q = 0; q < kFhGroups; ++q)
#pragma unroll
      for (int j = 0; j < 32; ++j) v[q][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      const float4* wp = reinterpret_cast<const float4*>(ws + j * 256) + lane * 2;
      const float4 u0 = wp[0], u1 = wp[1];
#pragma unroll
      for (int i = 0; i < kFhPx; ++i) {
        const int idx = i * 18 + j;
        v[idx >> 5][idx & 31] = xa[i].x * u0.x + xa[i].y * u0.y + xa[i].z * u0.z + xa[i].w * u0.w + xb[i].x * u1.x + xb[i].y * u1.y +
                                xb[i].z * u1.z + xb[i].w * u1.w;
      }
    }
#pragma unroll
    for (int q = 0; q < kFhGroups; ++q) {
      const float tot = warp_reduce_transpose32(v[q], lane);
      const int idx = q * 32 + lane;  // = pixel * 18 + j
      if (idx < kFhPx * 18 && p0 + idx / 18 < npix) y[p0 * 18 + idx] = tot;
    }
  }
}

__global__ void __launch_bounds__(256) flowhead2_gather_update_kernel(const float* __restrict__ y, float2 bias, float2* __restrict__ coords1,
                                                                      float2* __restrict__ flow, float* __restrict__ hx, int hx_stride,
                                                                      int hx_off, float* __restrict__ rhx, int rhx_stride, int rhx_off,
                                                                      int64_t npix, int h, int w) {
  pdl_wait();      // no-op unless launched with the PDL attribute (sdof_flowhead2_gather_update)
  pdl_trigger();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int rem = (int)(p % ((int64_t)h * w));
    const int yy = rem / w, xx = rem - yy * w;
    const float2 d = flowhead2_gather(y, bias, p, yy, xx, h, w);
    float2 c = coords1[p];
    c.x += d.x;
    c.y += d.y;
    coords1[p] = c;
    const float2 f = make_float2(c.x - (float)xx, c.y - (float)yy);
    flow[p] = f;
    if (hx) *reinterpret_cast<float2*>(hx + p * hx_stride + hx_off) = f;
    if (rhx) *reinterpret_cast<float2*>(rhx + p * rhx_stride + rhx_off) = f;
  }
}

}  // namespace sdof

extern "C" {

int sdof_conv7x7_c2_relu(const float* flow, const float* wT, const float* bias, float* out, int B, int h, int w, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(flow && wT && bias && out, "sdof_conv7x7_c2_relu: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h >= 1 && w >= 1, "sdof_conv7x7_c2_relu: bad sizes");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(flow) & 7) | (reinterpret_cast<uintptr_t>(wT) & 15) | (reinterpret_cast<uintptr_t>(bias) & 15) |
                (reinterpret_cast<uintptr_t>(out) & 15)) == 0, "sdof_conv7x7_c2_relu: misaligned pointer");
  if (B == 0) return SDOF_OK;
  static bool attr_set[64] = {};
  int dev = 0;
  SDOF_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    SDOF_CUDA(cudaFuncSetAttribute(conv7x7_c2_relu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC7Smem));
    attr_set[dev] = true;
  }
  const int tx = ceil_div(w, kC7Tile), ty = ceil_div(h, kC7Tile);
  const int64_t tiles = (int64_t)tx * ty * B;
  SDOF_REQUIRE(tiles < 0x7fffffffLL, "sdof_conv7x7_c2_relu: too many tiles");
  conv7x7_c2_relu_kernel<<<(unsigned)tiles, kC7Threads, kC7Smem, as_stream(stream)>>>(reinterpret_cast<const float2*>(flow), wT, bias, out,
                                                                                     h, w, tx, ty);
  SDOF_LAUNCH_CHECK("conv7x7_c2_relu_kernel");
  return SDOF_OK;
}

int sdof_flowhead2_update(const float* x, const float* w2, float bias_x, float bias_y, float* coords1, float* flow, float* hx, int hx_stride,
                          int hx_off, float* rhx, int rhx_stride, int rhx_off, int B, int h, int w, float* scratch, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(x && w2 && coords1 && flow && scratch, "sdof_flowhead2_update: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h >= 1 && w >= 1, "sdof_flowhead2_update: bad sizes");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w2)) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 7) == 0,
               "sdof_flowhead2_update: x, w2 must be 16-byte aligned, scratch 8-byte aligned");
  SDOF_REQUIRE((hx_stride % 2 == 0) && (hx_off % 2 == 0) && (rhx_stride % 2 == 0) && (rhx_off % 2 == 0),
               "sdof_flowhead2_update: strides/offsets must be even (float2 stores)");
  const int64_t npix = (int64_t)B * h * w;
  if (npix == 0) return SDOF_OK;
  {
    // one CTA per SM: every CTA first stages the 18 KB of filters, which is the kernel's main fixed cost (in-graph at
    // 96x64, taps + gather: 1 CTA/SM 8.0 us, 2/SM 9.6, 3/SM 11.2, 6/SM with one pixel per warp 16.7); SDOF_FH_PX / SDOF_FH_CAP
    // are the experiment switches behind those numbers
    static int px_mode = -1, cap_mode = 1;
    if (px_mode < 0) {
      const char* e = getenv("SDOF_FH_PX");
      px_mode = e ? atoi(e) : 2;
      const char* c = getenv("SDOF_FH_CAP");
      cap_mode = c && atoi(c) > 0 ? atoi(c) : 1;
    }
    const int px = px_mode == 1 ? 1 : 2;
    const int64_t want = ceil_div64(ceil_div64(npix, px), 8);
    const int64_t cap = (int64_t)sm_count() * cap_mode;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (px == 1)
      flowhead2_taps_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(x, w2, scratch, npix);
    else
      flowhead2_taps_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(x, w2, scratch, npix);
  }
  SDOF_LAUNCH_CHECK("flowhead2_taps_kernel");
  flowhead2_gather_update_kernel<<<grid_for(npix, 256, 8), 256, 0, as_stream(stream)>>>(
      scratch, make_float2(bias_x, bias_y), reinterpret_cast<float2*>(coords1), reinterpret_cast<float2*>(flow), hx, hx_stride, hx_off, rhx,
      rhx_stride, rhx_off, npix, h, w);
  SDOF_LAUNCH_CHECK("flowhead2_gather_update_kernel");
  return SDOF_OK;
}

// The second half of sdof_flowhead2_update alone (the tap products come from sdof_flowhead2_taps_h on fp16 activations).
int sdof_flowhead2_gather_update(const float* scratch, float bias_x, float bias_y, float* coords1, float* flow, float* hx, int hx_stride,
                                 int hx_off, int B, int h, int w, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(scratch && coords1 && flow, "sdof_flowhead2_gather_update: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h >= 1 && w >= 1, "sdof_flowhead2_gather_update: bad sizes");
  SDOF_REQUIRE((hx_stride % 2 == 0) && (hx_off % 2 == 0), "sdof_flowhead2_gather_update: strides/offsets must be even (float2 stores)");
  const int64_t npix = (int64_t)B * h * w;
  if (npix == 0) return SDOF_OK;
  SDOF_CUDA(launch_pdl(flowhead2_gather_update_kernel, dim3(grid_for(npix, 256, 8)), dim3(256), 0, as_stream(stream), scratch,
                       make_float2(bias_x, bias_y), reinterpret_cast<float2*>(coords1), reinterpret_cast<float2*>(flow), hx, hx_stride, hx_off,
                       static_cast<float*>(nullptr), 0, 0, npix, h, w));
  SDOF_LAUNCH_CHECK("flowhead2_gather_update_kernel");
  return SDOF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Input side of RAFT_2.calc / RAFT.forward (ofgen.py:72-76, raft.py:89-90) in one pass: uint8 HWC frame ->
// replicate-padded (InputPadder, utils/utils.py:7-19), normalised 2*(x/255)-1 fp32 NHWC image that the channels-last
// encoder convolutions consume directly.  Replaces permute + float + pad + contiguous + div + mul + sub + the
// NCHW -> NHWC copy (about a dozen ATen launches per pair).
namespace sdof {
// Cout = 3, or 4 with a zero fourth channel: cuDNN only runs its tensor-core implicit GEMM on NHWC inputs whose channel
// count is a multiple of 4 (the 3-channel stem convolution otherwise falls back to a CUDA-core engine, 266 us per pair).
__global__ void __launch_bounds__(256) normalize_pad_u8_nhwc_kernel(const unsigned char* __restrict__ img, int B, int H, int W, int top, int left,
                                                                    int Hp, int Wp, int Cout, int swap_rb, float* __restrict__ out) {
  const int64_t total = (int64_t)B * Hp * Wp;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(p % Wp);
    const int y = (int)((p / Wp) % Hp);
    const int b = (int)(p / ((int64_t)Wp * Hp));
    const int sy = min(max(y - top, 0), H - 1), sx = min(max(x - left, 0), W - 1);
    const unsigned char* q = img + (((int64_t)b * H + sy) * W + sx) * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn((float)q[swap_rb ? 2 - c : c], 255.0f)), 1.0f);
    if (Cout == 4) {
      reinterpret_cast<float4*>(out)[p] = make_float4(v[0], v[1], v[2], 0.f);
    } else {
      float* o = out + p * 3;
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
  }
}
}  // namespace sdof

extern "C" int sdof_normalize_pad_u8_nhwc(const uint8_t* img, int B, int H, int W, int top, int left, int Hp, int Wp, int Cout, int swap_rb,
                                          float* out, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(img && out, "sdof_normalize_pad_u8_nhwc: NULL pointer");
  SDOF_REQUIRE(B >= 0 && H >= 1 && W >= 1 && top >= 0 && left >= 0 && Hp >= H + top && Wp >= W + left,
               "sdof_normalize_pad_u8_nhwc: bad sizes H=%d W=%d top=%d left=%d Hp=%d Wp=%d", H, W, top, left, Hp, Wp);
  SDOF_REQUIRE(Cout == 3 || (Cout == 4 && (reinterpret_cast<uintptr_t>(out) & 15) == 0),
               "sdof_normalize_pad_u8_nhwc: Cout must be 3, or 4 with a 16-byte aligned output");
  const int64_t total = (int64_t)B * Hp * Wp;
  if (total == 0) return SDOF_OK;
  normalize_pad_u8_nhwc_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(img, B, H, W, top, left, Hp, Wp, Cout, swap_rb, out);
  SDOF_LAUNCH_CHECK("normalize_pad_u8_nhwc_kernel");
  return SDOF_OK;
}
