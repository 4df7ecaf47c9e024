// Per-iteration correlation lookups.
//
//  corr_lookup_kernel : windowed bilinear lookup in the precomputed pyramid.
//  alt_corr_kernel    : on-the-fly windowed correlation (no volume), warp-shuffle
//                       reductions over the channel dimension.
//
// Both are gathers, so the layout work is on the OUTPUT side: a CTA owns 32 consecutive
// source pixels, one warp per pixel; each warp stages its (2r+2)^2 integer-grid window
// in shared memory, blends the (2r+1)^2 outputs (all taps of a level share one (fx,fy)),
// and parks them in a [channels][32] shared tile so the [B, channels, h1, w1] result is
// written as full 128-byte lines.  Algorithmic bytes per source pixel and iteration:
// L*(2r+2)^2*4 read + L*(2r+1)^2*4 written = 2896 B for L=4, r=4.
#include <cuda_fp16.h>

#include "corr.cuh"

namespace sdof {

constexpr int kLookupPx = 32;                   // source pixels per CTA = warps per CTA
constexpr int kLookupThreads = kLookupPx * 32;  // 1024
constexpr int kStagePitch = kLookupPx + 1;      // +1: conflict-free column writes

struct LookupLevels {
  int levels;
  const void* base[SDOF_MAX_LEVELS];   // float or __half maps (template parameter of the kernel)
  const void* header;                  // start of the pyramid buffer (fp16 pyramids: float[0] = factor)
  long long pitch[SDOF_MAX_LEVELS];
  int h[SDOF_MAX_LEVELS], w[SDOF_MAX_LEVELS], wp[SDOF_MAX_LEVELS];
};

// floor with NaN / huge values mapped far outside any map (whole window reads zero)
__device__ __forceinline__ int safe_floor(float v, float* frac) {
  if (!(v > -1.0e6f && v < 1.0e6f)) {
    *frac = 0.f;
    return -(1 << 24);
  }
  const float f = floorf(v);
  *frac = v - f;
  return (int)f;
}

// blend the (D+1)x(D+1) window `win` (row-major, [wy][wx]) into D*D outputs, x-major
// channel order k = D*ix + iy (RAFT/core/corr.py:37-43), and park them in the stage tile.
__device__ __forceinline__ void blend_window(const float* __restrict__ win, int D, float fx, float fy, float scale,
                                             float* __restrict__ stage_col, int lane, int T1 = 0) {
  if (T1 == 0) T1 = D + 1;   // window row stride
  // `scale` goes into the four blend weights (same arithmetic as blend_window_strided: both output layouts agree bit for bit)
  const float w00 = (1.f - fx) * (1.f - fy) * scale, w01 = fx * (1.f - fy) * scale, w10 = (1.f - fx) * fy * scale, w11 = fx * fy * scale;
  for (int k = lane; k < D * D; k += 32) {
    const int ix = k / D, iy = k - ix * D;
    const float* q = win + iy * T1 + ix;
    stage_col[k * kStagePitch] = q[0] * w00 + q[1] * w01 + q[T1] * w10 + q[T1 + 1] * w11;
  }
}

__device__ __forceinline__ void blend_window_strided(const float* __restrict__ win, int D, float fx, float fy,
                                                     float* __restrict__ dst, int stride, int lane, int T1 = 0, float scale = 1.0f) {
  if (T1 == 0) T1 = D + 1;   // window row stride
  // the (power-of-two, or 1) scale of an fp16 pyramid goes into the four blend weights
  const float w00 = (1.f - fx) * (1.f - fy) * scale, w01 = fx * (1.f - fy) * scale, w10 = (1.f - fx) * fy * scale, w11 = fx * fy * scale;
  for (int k = lane; k < D * D; k += 32) {
    const int ix = k / D, iy = k - ix * D;
    const float* q = win + iy * T1 + ix;
    dst[k * stride] = q[0] * w00 + q[1] * w01 + q[T1] * w10 + q[T1 + 1] * w11;
  }
}

// Compile-time window: the (output index -> window offset) map of a lane is the same for every level, so it is computed once
// (`BlendMap`) and a level's blend is 4 LDS + 4 FMA + 1 store per output, fully unrolled (the generic loops above spend ~35
// instructions per output on index arithmetic: 760 instructions per pixel in the round-1 kernel).
template <int D>
struct BlendMap {
  static constexpr int kTrips = (D * D + 31) / 32;
  int off[kTrips];   // iy * WS + ix of output k = lane + 32 t (window offset of its top-left tap)
  __device__ __forceinline__ void init(int lane, int WS) {
#pragma unroll
    for (int t = 0; t < kTrips; ++t) {
      const int k = lane + 32 * t;
      const int ix = k / D, iy = k - ix * D;
      off[t] = iy * WS + ix;
    }
  }
};

__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__half* p, float v) { *p = __float2half_rn(v); }

template <int D, typename OT>
__device__ __forceinline__ void blend_fast(const float* __restrict__ win, const BlendMap<D>& bm, int WS, float fx, float fy, float scale,
                                           OT* __restrict__ dst, int stride, int lane) {
  const float w00 = (1.f - fx) * (1.f - fy) * scale, w01 = fx * (1.f - fy) * scale, w10 = (1.f - fx) * fy * scale, w11 = fx * fy * scale;
#pragma unroll
  for (int t = 0; t < BlendMap<D>::kTrips; ++t) {
    const int k = lane + 32 * t;
    if (k < D * D) {
      const float* q = win + bm.off[t];
      store_out(dst + k * stride, q[0] * w00 + q[1] * w01 + q[WS] * w10 + q[WS + 1] * w11);
    }
  }
}

// coalesced write-out of the CTA's [channels][32] stage tile
__device__ __forceinline__ void flush_stage(const float* __restrict__ stage, int channels, float* __restrict__ out,
                                            int64_t out_chan_stride, int p0, int N1) {
  for (int i = threadIdx.x; i < channels * kLookupPx; i += blockDim.x) {
    const int k = i >> 5, j = i & 31;
    if (p0 + j < N1) out[(int64_t)k * out_chan_stride + p0 + j] = stage[k * kStagePitch + j];
  }
}

// PX = source pixels (= warps) per CTA.  The planar output needs PX = 32 (full 128-byte lines through the stage
// tile); the channels-last output has no stage, so it runs with 8-warp CTAs: 768 CTAs at 96x64 instead of 192
// 1024-thread CTAs that fill the 148 SMs 1.3 times.
constexpr int kLookupPxNhwc = 8;

// ET = float: the fp32 pyramid (window taps gathered one by one).  ET = __half: the fp16 pyramid -- a window row is gathered
// as 2r/2+2 aligned 32-bit words (two taps each), half the load instructions and half the sectors of the fp32 form.
// Deferred coords update (channels-last form only): the previous iteration's flow head left its tap products in `taps`
// (sdof_flowhead2_taps_h); instead of a separate kernel that sums them into coords1, the lookup does
// coords = coords_in + gather(taps) itself (every lane computes the same scalar sum: broadcast loads) and lane 0 writes the new
// coordinates / flow for the kernels after it.  coords_out == nullptr: plain lookup.
struct LookupGather {
  const float* taps = nullptr;     // [B*h1*w1][18]; nullptr with coords_out set: coords pass through (first iteration)
  float2 bias = {0.f, 0.f};
  float2* coords_out = nullptr;    // [B,h1,w1,2]   (must not alias coords_in: the flow branch reads coords_in concurrently)
  float2* flow_out = nullptr;      // [B,h1,w1,2] = coords_out - pixel grid
  int w1 = 0;
};

template <int R_T, int L_T, int PX, typename ET>
__global__ void __launch_bounds__(PX * 32) corr_lookup_kernel(LookupLevels lv, const float* __restrict__ coords,
                                                              int N1, int r_rt, float* __restrict__ out, __half* __restrict__ out16 = nullptr,
                                                              int out16_channels = 0, LookupGather g = LookupGather()) {
  constexpr bool nhwc = PX != kLookupPx;   // the 8-warp CTA shape is the channels-last form (no stage tile)
  extern __shared__ __align__(16) float smem[];
  constexpr bool kHalf = sizeof(ET) == 2;
  const int r = R_T ? R_T : r_rt;
  const int L = L_T ? L_T : lv.levels;
  const int D = 2 * r + 1, T1 = D + 1, DD = D * D;
  const int WS = kHalf ? T1 + 2 : T1;       // window row stride in shared memory (the word gather may start one tap early)
  const int T = WS * T1;                    // floats per level window
  float* stage = smem;                                                    // [L*DD][kStagePitch]   (planar only)
  float* win_all = smem + (PX == kLookupPx ? L * DD * kStagePitch : 0);   // [PX warps][L][T]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * PX;
  const int p = p0 + warp;
  float* win = win_all + warp * L * T;
  pdl_wait();      // no-op unless launched with the PDL attribute (sdof_corr_lookup_h: overlaps the tail of the coords update)
  pdl_trigger();
  // fp16 pyramid: stored values are raw accumulators of the auto-ranged operands; the header holds the factor back to
  // correlation units (sdof_corr_pyramid_layout_ex).  levels share one header, lv.base[0] is 128 bytes behind it.
  const float factor = kHalf ? __ldg(reinterpret_cast<const float*>(lv.header)) : 1.0f;
  if (p < N1) {
    // nhwc: coords [B,h,w,2] and out [B,h,w,CH] (channels-last, for the NHWC update loop); else the
    // reference's planar layouts coords [B,2,h,w] / out [B,CH,h,w]
    float cx = nhwc ? coords[((int64_t)b * N1 + p) * 2] : coords[((int64_t)b * 2 + 0) * N1 + p];
    float cy = nhwc ? coords[((int64_t)b * N1 + p) * 2 + 1] : coords[((int64_t)b * 2 + 1) * N1 + p];
    const int64_t row = (int64_t)b * N1 + p;
    if constexpr (nhwc) {
      if (g.coords_out != nullptr) {
        const int yy = p / g.w1, xx = p - yy * g.w1;
        if (g.taps != nullptr) {
          const float2 d = flowhead2_gather(g.taps, g.bias, row, yy, xx, N1 / g.w1, g.w1);
          cx += d.x;
          cy += d.y;
        }
        if (lane == 0) {
          g.coords_out[row] = make_float2(cx, cy);
          g.flow_out[row] = make_float2(cx - (float)xx, cy - (float)yy);
        }
      }
    }
    float fxs[L_T ? L_T : SDOF_MAX_LEVELS], fys[L_T ? L_T : SDOF_MAX_LEVELS];
    int xos[L_T ? L_T : SDOF_MAX_LEVELS];
    // gather: with compile-time (r, L) every level's loads are issued before any is consumed
    constexpr int kItems = kHalf ? (R_T ? (2 * R_T + 2) * (R_T + 2) : 18 * 10) : (R_T ? (2 * R_T + 2) * (2 * R_T + 2) : 18 * 18);
    constexpr int kTrips = (kItems + 31) / 32;
#pragma unroll
    for (int l = 0; l < (L_T ? L_T : SDOF_MAX_LEVELS); ++l) {
      if (l >= L) break;
      const float s = 1.0f / (float)(1 << l);  // coords / 2**l is exact
      const int x0 = safe_floor(cx * s, &fxs[l]) - r;
      const int y0 = safe_floor(cy * s, &fys[l]) - r;
      const int h = lv.h[l], w = lv.w[l], wp = lv.wp[l];
      if constexpr (kHalf) {
        const __half* map = reinterpret_cast<const __half*>(lv.base[l]) + row * lv.pitch[l];
        const int xa = x0 & ~1;               // floor to even (also for negative x0): words are 4-byte aligned (wp % 8 == 0)
        xos[l] = x0 - xa;
        const int WR = (T1 >> 1) + 1;         // words per window row
#pragma unroll
        for (int j = 0; j < kTrips; ++j) {
          const int t = lane + 32 * j;
          if (t < WR * T1) {
            const int wy = t / WR, wx = t - wy * WR;
            const int gy = y0 + wy, gx = xa + 2 * wx;
            float2 v = make_float2(0.f, 0.f);
            if ((unsigned)gy < (unsigned)h && (unsigned)gx < (unsigned)w) {
              v = __half22float2(__ldg(reinterpret_cast<const __half2*>(map + (int64_t)gy * wp + gx)));
              if (gx + 1 >= w) v.y = 0.f;     // odd width: the pair's second tap is the padding column
            }
            *reinterpret_cast<float2*>(win + l * T + wy * WS + 2 * wx) = v;
          }
        }
      } else {
        const float* map = reinterpret_cast<const float*>(lv.base[l]) + row * lv.pitch[l];
        xos[l] = 0;
#pragma unroll
        for (int j = 0; j < kTrips; ++j) {
          const int t = lane + 32 * j;
          if (t < T) {
            const int wy = t / T1, wx = t - wy * T1;
            const int gy = y0 + wy, gx = x0 + wx;
            float v = 0.f;
            if ((unsigned)gy < (unsigned)h && (unsigned)gx < (unsigned)w) v = __ldg(map + (int64_t)gy * wp + gx);
            win[l * T + t] = v;
          }
        }
      }
    }
    __syncwarp();
    if constexpr (R_T != 0) {
      BlendMap<2 * R_T + 1> bm;
      bm.init(lane, WS);
#pragma unroll
      for (int l = 0; l < (L_T ? L_T : SDOF_MAX_LEVELS); ++l) {
        if (l >= L) break;
        if (nhwc && out16)  // fp16 channels-last output, rows padded to out16_channels (the cuDNN fp16 convc1 needs C % 8 == 0)
          blend_fast<2 * R_T + 1>(win + l * T + xos[l], bm, WS, fxs[l], fys[l], factor, out16 + ((int64_t)b * N1 + p) * out16_channels + l * DD, 1, lane);
        else if (nhwc)  // the pixel's channels are contiguous: write them straight out (stride 1 between channels)
          blend_fast<2 * R_T + 1>(win + l * T + xos[l], bm, WS, fxs[l], fys[l], factor, out + ((int64_t)b * N1 + p) * (L * DD) + l * DD, 1, lane);
        else
          blend_fast<2 * R_T + 1>(win + l * T + xos[l], bm, WS, fxs[l], fys[l], factor, stage + l * DD * kStagePitch + warp, kStagePitch, lane);
      }
      if (nhwc && out16 && lane < out16_channels - L * DD)   // zero the padding channels
        out16[((int64_t)b * N1 + p) * out16_channels + L * DD + lane] = __float2half_rn(0.f);
    } else {
#pragma unroll
      for (int l = 0; l < (L_T ? L_T : SDOF_MAX_LEVELS); ++l) {
        if (l >= L) break;
        if (nhwc)  // the pixel's channels are contiguous: write them straight out (stride 1 between channels)
          blend_window_strided(win + l * T + xos[l], D, fxs[l], fys[l], out + ((int64_t)b * N1 + p) * (L * DD) + l * DD, 1, lane, WS, factor);
        else
          blend_window(win + l * T + xos[l], D, fxs[l], fys[l], factor, stage + l * DD * kStagePitch + warp, lane, WS);
      }
    }
  }
  if (nhwc) return;
  __syncthreads();
  flush_stage(stage, L * DD, out + (int64_t)b * L * DD * N1, N1, p0, N1);
}

// ---------------------------------------------------------------------------------------------
// On-the-fly correlation: for each of the (2r+2)^2 integer taps around floor(coords)-r take the
// C-long dot product <fmap1[p], fmap2[tap]> (correlation_kernel.cu:59-90), then blend exactly like
// the lookup.  One warp per source pixel: lanes split the channels (float4 per lane per 128
// channels); 32 taps are reduced together with a transposing butterfly (31 shuffles per 32 taps).
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
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
  return v[0];  // lane L holds the total of tap L of the group
}

struct AltCorrArgs {
  const float* fmap1;   // [B, N1, C]
  const float* fmap2;   // [B, H2*W2, C]
  const float* coords;  // element (bn, p, xy) at coords[bn*c_set + p*c_px + xy*c_xy]
  long long c_set, c_px, c_xy;
  int sets;  // coordinate sets per batch item (N of alt_cuda_corr.forward)
  int N1, H2, W2, C, r;
  float coord_scale, out_scale;
  float* out;            // channel (chan_offset + set*DD + k) of [B, out_channels, N1]
  int out_channels, chan_offset;
};

template <int R_T>
__global__ void __launch_bounds__(kLookupThreads) alt_corr_kernel(AltCorrArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int r = R_T ? R_T : a.r;
  const int D = 2 * r + 1, T1 = D + 1, T = T1 * T1, DD = D * D;
  const int C = a.C, C4 = C >> 2;
  float* stage = smem;                            // [DD][kStagePitch]
  float* win_all = stage + DD * kStagePitch;      // [32][T]
  float* f1_all = smem + ((DD * kStagePitch + kLookupPx * T + 3) & ~3);  // [32][C], 16-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = blockIdx.y, b = bn / a.sets, set = bn - b * a.sets;
  const int p0 = blockIdx.x * kLookupPx;
  const int p = p0 + warp;
  float* win = win_all + warp * T;
  if (p < a.N1) {
    float4* f1s = reinterpret_cast<float4*>(f1_all + warp * C);
    const float4* f1g = reinterpret_cast<const float4*>(a.fmap1 + ((int64_t)b * a.N1 + p) * C);
    for (int c = lane; c < C4; c += 32) f1s[c] = __ldg(f1g + c);
    const float* cp = a.coords + (int64_t)bn * a.c_set + (int64_t)p * a.c_px;
    float fx, fy;
    const int x0 = safe_floor(cp[0] * a.coord_scale, &fx) - r;
    const int y0 = safe_floor(cp[a.c_xy] * a.coord_scale, &fy) - r;
    __syncwarp();
    const float4* f2b = reinterpret_cast<const float4*>(a.fmap2 + (int64_t)b * a.H2 * a.W2 * C);
    for (int t0 = 0; t0 < T; t0 += 32) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = 0.f;
        const int t = t0 + i;
        if (t < T) {
          const int wy = t / T1, wx = t - wy * T1;
          const int gy = y0 + wy, gx = x0 + wx;
          if ((unsigned)gy < (unsigned)a.H2 && (unsigned)gx < (unsigned)a.W2) {  // warp-uniform
            const float4* f2 = f2b + ((int64_t)gy * a.W2 + gx) * C4;
            float s = 0.f;
            for (int c = lane; c < C4; c += 32) {
              const float4 u = f1s[c], q = __ldg(f2 + c);
              s = fmaf(u.x, q.x, s);
              s = fmaf(u.y, q.y, s);
              s = fmaf(u.z, q.z, s);
              s = fmaf(u.w, q.w, s);
            }
            v[i] = s;
          }
        }
      }
      const float tot = butterfly32(v, lane);
      if (t0 + lane < T) win[t0 + lane] = tot;
    }
    __syncwarp();
    blend_window(win, D, fx, fy, a.out_scale, stage + warp, lane);
  }
  __syncthreads();
  flush_stage(stage, DD, a.out + ((int64_t)b * a.out_channels + a.chan_offset + (int64_t)set * DD) * a.N1, a.N1, p0, a.N1);
}

static int launch_alt_corr(const char* name, const AltCorrArgs& a, int B, cudaStream_t st) {
  const int D = 2 * a.r + 1, T = (D + 1) * (D + 1), DD = D * D;
  // stage + windows, rounded so the feature rows start 16-byte aligned
  size_t floats = (size_t)DD * kStagePitch + (size_t)kLookupPx * T;
  floats = (floats + 3) & ~(size_t)3;
  const size_t smem = (floats + (size_t)kLookupPx * a.C) * sizeof(float);
  SDOF_REQUIRE(smem <= 200 * 1024, "%s: C=%d / radius=%d need %zu bytes of shared memory (limit 200 KB)", name, a.C, a.r,
               smem);
  AltCorrArgs args = a;
  dim3 grid(ceil_div(a.N1, kLookupPx), B * a.sets);
  SDOF_REQUIRE(grid.y <= 65535, "%s: B*N > 65535 not supported", name);
#define SDOF_ALT_LAUNCH(RT)                                                                                        \
  do {                                                                                                             \
    SDOF_CUDA(cudaFuncSetAttribute(alt_corr_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    alt_corr_kernel<RT><<<grid, kLookupThreads, smem, st>>>(args);                                                 \
  } while (0)
  if (a.r == 4)
    SDOF_ALT_LAUNCH(4);
  else if (a.r == 3)
    SDOF_ALT_LAUNCH(3);
  else
    SDOF_ALT_LAUNCH(0);
#undef SDOF_ALT_LAUNCH
  SDOF_LAUNCH_CHECK(name);
  return SDOF_OK;
}

}  // namespace sdof

extern "C" {

static int corr_lookup_impl(const char* name, const void* pyramid, int elem_bytes, const float* coords, int B, int h1, int w1,
                            int h2, int w2, int levels, int radius, float* out, int nhwc, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(pyramid && coords && out, "%s: NULL pointer", name);
  SDOF_REQUIRE(B >= 0 && h1 >= 1 && w1 >= 1 && h2 >= 1 && w2 >= 1, "%s: bad sizes", name);
  SDOF_REQUIRE(radius >= 0 && radius <= 8, "%s: radius must be in [0,8], got %d", name, radius);
  SDOF_REQUIRE(B <= 65535, "%s: B > 65535 not supported", name);
  SDOF_REQUIRE(elem_bytes == 4 || elem_bytes == 2, "%s: elem_bytes must be 4 (fp32 pyramid) or 2 (fp16 pyramid)", name);
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(pyramid) & 3) == 0, "%s: pyramid must be 4-byte aligned", name);
  sdof_pyramid_layout lay;
  const int N1 = h1 * w1;
  int rc = sdof_corr_pyramid_layout_ex((int64_t)B * N1, h2, w2, levels, elem_bytes, &lay);
  if (rc) return rc;
  if (B == 0) return SDOF_OK;
  LookupLevels lv;
  lv.levels = levels;
  lv.header = pyramid;
  for (int l = 0; l < SDOF_MAX_LEVELS; ++l) {
    const int ll = l < levels ? l : 0;
    lv.base[l] = reinterpret_cast<const uint8_t*>(pyramid) + lay.offset[ll] * elem_bytes;
    lv.pitch[l] = lay.pitch[ll];
    lv.h[l] = lay.h[ll];
    lv.w[l] = lay.w[ll];
    lv.wp[l] = lay.wp[ll];
  }
  const int D = 2 * radius + 1, T1 = D + 1, DD = D * D;
  const int T = (elem_bytes == 2 ? T1 + 2 : T1) * T1;
  const int px = nhwc ? kLookupPxNhwc : kLookupPx;
  const size_t smem = ((nhwc ? 0 : (size_t)levels * DD * kStagePitch) + (size_t)px * levels * T) * sizeof(float);
  SDOF_REQUIRE(smem <= 200 * 1024, "%s: levels=%d radius=%d need %zu bytes of shared memory", name, levels, radius, smem);
  dim3 grid(ceil_div(N1, px), B);
  cudaStream_t st = as_stream(stream);
#define SDOF_LOOKUP_LAUNCH(RT, LT, PX, TT)                                                                                       \
  do {                                                                                                                           \
    SDOF_CUDA(cudaFuncSetAttribute(corr_lookup_kernel<RT, LT, PX, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    corr_lookup_kernel<RT, LT, PX, TT><<<grid, PX * 32, smem, st>>>(lv, coords, N1, radius, out);                          \
  } while (0)
#define SDOF_LOOKUP_DISPATCH(RT, LT)                          \
  do {                                                        \
    if (nhwc && elem_bytes == 2)                              \
      SDOF_LOOKUP_LAUNCH(RT, LT, kLookupPxNhwc, __half);      \
    else if (nhwc)                                            \
      SDOF_LOOKUP_LAUNCH(RT, LT, kLookupPxNhwc, float);       \
    else if (elem_bytes == 2)                                 \
      SDOF_LOOKUP_LAUNCH(RT, LT, kLookupPx, __half);          \
    else                                                      \
      SDOF_LOOKUP_LAUNCH(RT, LT, kLookupPx, float);           \
  } while (0)
  if (radius == 4 && levels == 4)
    SDOF_LOOKUP_DISPATCH(4, 4);
  else if (radius == 3 && levels == 4)
    SDOF_LOOKUP_DISPATCH(3, 4);
  else
    SDOF_LOOKUP_DISPATCH(0, 0);
#undef SDOF_LOOKUP_DISPATCH
#undef SDOF_LOOKUP_LAUNCH
  SDOF_LAUNCH_CHECK("corr_lookup_kernel");
  return SDOF_OK;
}

int sdof_corr_lookup(const float* pyramid, const float* coords, int B, int h1, int w1, int h2, int w2, int levels,
                     int radius, float* out, sdof_stream_t stream) {
  return corr_lookup_impl("sdof_corr_lookup", pyramid, 4, coords, B, h1, w1, h2, w2, levels, radius, out, 0, stream);
}

int sdof_corr_lookup_nhwc(const float* pyramid, const float* coords, int B, int h1, int w1, int h2, int w2, int levels,
                          int radius, float* out, sdof_stream_t stream) {
  return corr_lookup_impl("sdof_corr_lookup_nhwc", pyramid, 4, coords, B, h1, w1, h2, w2, levels, radius, out, 1, stream);
}

// fp16 channels-last output for the fp16 update block: out16 [B,h1,w1,out_channels] halves, out_channels >= levels*(2r+1)^2 a
// multiple of 8 (padding channels are written as zeros); radius 4, 4 levels only (RAFT's configuration).
static int corr_lookup_h_impl(const char* name, const void* pyramid, int elem_bytes, const float* coords, int B, int h1, int w1, int h2, int w2,
                              int levels, int radius, void* out16, int out_channels, sdof::LookupGather g, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(pyramid && coords && out16, "%s: NULL pointer", name);
  SDOF_REQUIRE(B >= 0 && B <= 65535 && h1 >= 1 && w1 >= 1 && h2 >= 1 && w2 >= 1, "%s: bad sizes", name);
  SDOF_REQUIRE(radius == 4 && levels == 4, "%s: radius 4 / 4 levels only", name);
  SDOF_REQUIRE(out_channels >= 324 && out_channels <= 356 && out_channels % 8 == 0, "%s: out_channels must be a multiple of 8 in [328, 352]", name);
  SDOF_REQUIRE(elem_bytes == 4 || elem_bytes == 2, "%s: elem_bytes must be 4 or 2", name);
  sdof_pyramid_layout lay;
  const int N1 = h1 * w1;
  int rc = sdof_corr_pyramid_layout_ex((int64_t)B * N1, h2, w2, levels, elem_bytes, &lay);
  if (rc) return rc;
  if (B == 0) return SDOF_OK;
  LookupLevels lv;
  lv.levels = levels;
  lv.header = pyramid;
  for (int l = 0; l < SDOF_MAX_LEVELS; ++l) {
    const int ll = l < levels ? l : 0;
    lv.base[l] = reinterpret_cast<const uint8_t*>(pyramid) + lay.offset[ll] * elem_bytes;
    lv.pitch[l] = lay.pitch[ll];
    lv.h[l] = lay.h[ll];
    lv.w[l] = lay.w[ll];
    lv.wp[l] = lay.wp[ll];
  }
  const int T1 = 10;
  const int T = (elem_bytes == 2 ? T1 + 2 : T1) * T1;
  const size_t smem = (size_t)kLookupPxNhwc * levels * T * sizeof(float);
  dim3 grid(ceil_div(N1, kLookupPxNhwc), B);
  cudaStream_t st = as_stream(stream);
  // the pyramid is read 20 times per pair: its leading bytes (level 0 first) may stay in the persisting part of L2
  L2Window win;
  win.ptr = pyramid;
  win.bytes = (size_t)lay.total_floats * elem_bytes;
  {
    const size_t cap = l2_persist_bytes();
    if (cap > 0 && win.bytes > cap) win.bytes = cap;   // a window that fits entirely: level 0 (and what follows it) up to the cap
  }
  if (elem_bytes == 2)
    SDOF_CUDA(launch_pdl_win(corr_lookup_kernel<4, 4, kLookupPxNhwc, __half>, grid, dim3(kLookupPxNhwc * 32), smem, st, win, lv, coords, N1,
                             radius, static_cast<float*>(nullptr), reinterpret_cast<__half*>(out16), out_channels, g));
  else
    SDOF_CUDA(launch_pdl_win(corr_lookup_kernel<4, 4, kLookupPxNhwc, float>, grid, dim3(kLookupPxNhwc * 32), smem, st, win, lv, coords, N1,
                             radius, static_cast<float*>(nullptr), reinterpret_cast<__half*>(out16), out_channels, g));
  SDOF_LAUNCH_CHECK("corr_lookup_kernel");
  return SDOF_OK;
}

int sdof_corr_lookup_h(const void* pyramid, int elem_bytes, const float* coords, int B, int h1, int w1, int h2, int w2, int levels, int radius,
                       void* out16, int out_channels, sdof_stream_t stream) {
  return corr_lookup_h_impl("sdof_corr_lookup_h", pyramid, elem_bytes, coords, B, h1, w1, h2, w2, levels, radius, out16, out_channels,
                            sdof::LookupGather(), stream);
}

int sdof_corr_lookup_gather_h(const void* pyramid, int elem_bytes, const float* coords_in, const float* taps, float bias_x, float bias_y,
                              float* coords_out, float* flow_out, int B, int h1, int w1, int h2, int w2, int levels, int radius, void* out16,
                              int out_channels, sdof_stream_t stream) {
  using namespace sdof;
  const char* name = "sdof_corr_lookup_gather_h";
  SDOF_REQUIRE(coords_in && coords_out && flow_out, "%s: NULL pointer", name);
  SDOF_REQUIRE(coords_out != coords_in, "%s: coords_out must not alias coords_in (the flow branch reads coords_in concurrently)", name);
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(coords_out) | reinterpret_cast<uintptr_t>(flow_out) | reinterpret_cast<uintptr_t>(taps)) & 7) == 0,
               "%s: misaligned pointer", name);
  LookupGather g;
  g.taps = taps;
  g.bias = make_float2(bias_x, bias_y);
  g.coords_out = reinterpret_cast<float2*>(coords_out);
  g.flow_out = reinterpret_cast<float2*>(flow_out);
  g.w1 = w1;
  return corr_lookup_h_impl(name, pyramid, elem_bytes, coords_in, B, h1, w1, h2, w2, levels, radius, out16, out_channels, g, stream);
}

int sdof_corr_lookup_ex(const void* pyramid, int elem_bytes, const float* coords, int B, int h1, int w1, int h2, int w2, int levels,
                        int radius, float* out, int channels_last, sdof_stream_t stream) {
  return corr_lookup_impl("sdof_corr_lookup_ex", pyramid, elem_bytes, coords, B, h1, w1, h2, w2, levels, radius, out,
                          channels_last ? 1 : 0, stream);
}

static int check_alt(const char* name, const float* fmap1, const float* fmap2, const float* coords, float* out, int B,
                     int H1, int W1, int H2, int W2, int C, int radius) {
  SDOF_REQUIRE(fmap1 && fmap2 && coords && out, "%s: NULL pointer", name);
  SDOF_REQUIRE(B >= 0 && H1 >= 1 && W1 >= 1 && H2 >= 1 && W2 >= 1, "%s: bad sizes", name);
  SDOF_REQUIRE(C >= 4 && C % 4 == 0, "%s: C must be a positive multiple of 4, got %d", name, C);
  SDOF_REQUIRE(radius >= 0 && radius <= 8, "%s: radius must be in [0,8], got %d", name, radius);
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(fmap1) | reinterpret_cast<uintptr_t>(fmap2)) & 15) == 0,
               "%s: feature maps must be 16-byte aligned", name);
  return SDOF_OK;
}

int sdof_alt_corr_forward(const float* fmap1, const float* fmap2, const float* coords, int B, int H1, int W1, int H2,
                          int W2, int C, int N, int radius, float* corr, sdof_stream_t stream) {
  using namespace sdof;
  int rc = check_alt("sdof_alt_corr_forward", fmap1, fmap2, coords, corr, B, H1, W1, H2, W2, C, radius);
  if (rc) return rc;
  SDOF_REQUIRE(N >= 1, "sdof_alt_corr_forward: N must be >= 1");
  if (B == 0) return SDOF_OK;
  const int D = 2 * radius + 1;
  AltCorrArgs a;
  a.fmap1 = fmap1;
  a.fmap2 = fmap2;
  a.coords = coords;
  a.N1 = H1 * W1;
  a.c_set = (long long)a.N1 * 2;  // [B,N,H1,W1,2]: sets are contiguous, (x,y) interleaved
  a.c_px = 2;
  a.c_xy = 1;
  a.sets = N;
  a.H2 = H2;
  a.W2 = W2;
  a.C = C;
  a.r = radius;
  a.coord_scale = 1.f;
  a.out_scale = 1.f;
  a.out = corr;  // [B, N*DD, N1]
  a.out_channels = N * D * D;
  a.chan_offset = 0;
  return launch_alt_corr("sdof_alt_corr_forward", a, B, as_stream(stream));
}

int sdof_alt_corr_level(const float* fmap1, const float* fmap2_level, const float* coords, int B, int H1, int W1,
                        int H2, int W2, int C, int radius, float coord_scale, float out_scale, int chan_offset,
                        int out_channels, float* out, sdof_stream_t stream) {
  using namespace sdof;
  int rc = check_alt("sdof_alt_corr_level", fmap1, fmap2_level, coords, out, B, H1, W1, H2, W2, C, radius);
  if (rc) return rc;
  const int D = 2 * radius + 1;
  SDOF_REQUIRE(chan_offset >= 0 && chan_offset + D * D <= out_channels, "sdof_alt_corr_level: channel window out of range");
  if (B == 0) return SDOF_OK;
  AltCorrArgs a;
  a.fmap1 = fmap1;
  a.fmap2 = fmap2_level;
  a.coords = coords;
  a.N1 = H1 * W1;
  a.c_set = (long long)a.N1 * 2;  // [B,2,H1,W1] planar
  a.c_px = 1;
  a.c_xy = a.N1;
  a.sets = 1;
  a.H2 = H2;
  a.W2 = W2;
  a.C = C;
  a.r = radius;
  a.coord_scale = coord_scale;
  a.out_scale = out_scale;
  a.out = out;
  a.out_channels = out_channels;
  a.chan_offset = chan_offset;
  return launch_alt_corr("sdof_alt_corr_level", a, B, as_stream(stream));
}

}  // extern "C"
