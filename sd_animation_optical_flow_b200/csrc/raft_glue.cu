// Per-iteration glue of the RAFT update loop on dense channels-last (NHWC) buffers.
//
// The reference's update block (RAFT/core/update.py:79-136) runs ~100 small kernels per GRU iteration
// in eager PyTorch (cat, relu, sigmoid, tanh, mul, add, plus two layout transposes around every cuDNN
// convolution).  With the activations kept in NHWC and the `torch.cat`s replaced by persistent
// concatenated buffers, everything between the convolutions collapses into the four element-wise
// kernels below; the convolutions themselves stay in cuDNN (raft_fast.py).
//
//   relu_scatter   : dst[:, off:off+C] = relu(src)            (optionally into two buffers)
//   gru_rh         : rhx[:, 0:Hd] = sigmoid(zr[:, Hd:2Hd]) * h
//   gru_update     : h' = (1 - sigmoid(z)) * h + sigmoid(z) * tanh(q)   -> h (dense) and hx[:, 0:Hd]
//   flow_update    : coords1 += delta; flow = coords1 - grid  -> dense flow, hx / rhx flow slots
//   convex_upsample: RAFT.upsample_flow (RAFT/core/raft.py:72-83): softmax over the 9 taps of the 8x8 masks,
//                    weighted sum of the 3x3 neighbourhood of 8*flow  -> full-resolution [B,8h,8w,2]
// All are HBM/L2-bound element-wise work: float4 vectorised, grid-stride.
#include "sdof_common.cuh"

namespace sdof {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256) relu_scatter_kernel(const float4* __restrict__ src, const float4* __restrict__ bias,
                                                           int64_t npix, int C4, float* __restrict__ d1, int d1_stride, int d1_off,
                                                           float* __restrict__ d2, int d2_stride, int d2_off, int C_valid) {
  const int64_t total = npix * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C4;
    const int c = (int)(i - p * C4) * 4;
    float4 v = src[i];
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

__global__ void __launch_bounds__(256) gru_rh_kernel(const float4* __restrict__ zr, const float4* __restrict__ bias_zr,
                                                     const float4* __restrict__ h, float* __restrict__ rhx, int64_t npix,
                                                     int Hd4, int rhx_stride) {
  const int64_t total = npix * Hd4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / Hd4;
    const int c4 = (int)(i - p * Hd4);
    float4 r = zr[p * (2 * Hd4) + Hd4 + c4];
    if (bias_zr) r = add4(r, __ldg(bias_zr + Hd4 + c4));
    const float4 hv = h[i];
    float4 o;
    o.x = sigmoidf_(r.x) * hv.x; o.y = sigmoidf_(r.y) * hv.y; o.z = sigmoidf_(r.z) * hv.z; o.w = sigmoidf_(r.w) * hv.w;
    *reinterpret_cast<float4*>(rhx + p * rhx_stride + 4 * c4) = o;
  }
}

__global__ void __launch_bounds__(256) gru_update_kernel(const float4* __restrict__ zr, const float4* __restrict__ bias_zr,
                                                         const float4* __restrict__ q, const float4* __restrict__ bias_q,
                                                         float4* __restrict__ h, float* __restrict__ hx, int64_t npix,
                                                         int Hd4, int hx_stride) {
  const int64_t total = npix * Hd4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / Hd4;
    const int c4 = (int)(i - p * Hd4);
    float4 z = zr[p * (2 * Hd4) + c4];
    float4 qv = q[i];
    if (bias_zr) z = add4(z, __ldg(bias_zr + c4));
    if (bias_q) qv = add4(qv, __ldg(bias_q + c4));
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


int sdof_relu_scatter(const float* src, const float* bias, int64_t npix, int C, float* dst1, int dst1_stride, int dst1_off, float* dst2,
                      int dst2_stride, int dst2_off, int C_valid, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(src && dst1, "sdof_relu_scatter: NULL pointer");
  SDOF_REQUIRE(C > 0 && C % 4 == 0 && C_valid > 0 && C_valid <= C, "sdof_relu_scatter: C must be a multiple of 4, 0 < C_valid <= C");
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0, "sdof_relu_scatter: src must be 16-byte aligned");
  if (npix <= 0) return SDOF_OK;
  relu_scatter_kernel<<<grid_for(npix * (C / 4), 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<const float4*>(bias), npix, C / 4, dst1, dst1_stride, dst1_off, dst2,
      dst2_stride, dst2_off, C_valid);
  SDOF_LAUNCH_CHECK("relu_scatter_kernel");
  return SDOF_OK;
}

int sdof_gru_rh(const float* zr, const float* bias_zr, const float* h, float* rhx, int64_t npix, int hidden, int rhx_stride,
                sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(zr && h && rhx, "sdof_gru_rh: NULL pointer");
  SDOF_REQUIRE(hidden > 0 && hidden % 4 == 0 && rhx_stride % 4 == 0, "sdof_gru_rh: hidden and stride must be multiples of 4");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(zr) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(rhx)) & 15) == 0,
               "sdof_gru_rh: pointers must be 16-byte aligned");
  if (npix <= 0) return SDOF_OK;
  gru_rh_kernel<<<grid_for(npix * (hidden / 4), 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(zr), reinterpret_cast<const float4*>(bias_zr), reinterpret_cast<const float4*>(h), rhx, npix,
      hidden / 4, rhx_stride);
  SDOF_LAUNCH_CHECK("gru_rh_kernel");
  return SDOF_OK;
}

int sdof_gru_update(const float* zr, const float* bias_zr, const float* q, const float* bias_q, float* h, float* hx,
                    int64_t npix, int hidden, int hx_stride, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(zr && q && h && hx, "sdof_gru_update: NULL pointer");
  SDOF_REQUIRE(hidden > 0 && hidden % 4 == 0 && hx_stride % 4 == 0, "sdof_gru_update: hidden and stride must be multiples of 4");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(zr) | reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(h) |
                 reinterpret_cast<uintptr_t>(hx)) & 15) == 0, "sdof_gru_update: pointers must be 16-byte aligned");
  if (npix <= 0) return SDOF_OK;
  gru_update_kernel<<<grid_for(npix * (hidden / 4), 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(zr), reinterpret_cast<const float4*>(bias_zr), reinterpret_cast<const float4*>(q),
      reinterpret_cast<const float4*>(bias_q), reinterpret_cast<float4*>(h), hx, npix, hidden / 4, hx_stride);
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
