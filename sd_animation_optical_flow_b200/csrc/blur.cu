// The step after the hot path: what GuidedLDM.img2img_inpaint does to the warped frame and the
// inpainting mask before Stable Diffusion runs (guided_ldm_inpainting.py:290-309), bit-exact to Pillow:
//
//   image_mask = mask.filter(ImageFilter.GaussianBlur(mask_blur))     -> blur_composite_kernel (blur part)
//   image      = Image.composite(reference_img, image, image_mask)    -> blur_composite_kernel (composite part)
//   latmask    = around(image_mask.resize((w/8, h/8)) / 255)          -> resample_h_kernel + resample_v_kernel
//
// Pillow's GaussianBlur is three passes per axis of an "extended box" filter in 8.24 fixed point with a uint8
// rounding after every pass and replicated edges (libImaging/BoxBlur.c); its resize is a separable antialiased
// bicubic with 22-bit integer coefficients (libImaging/Resample.c).  Both are byte/integer work bound by HBM:
// per pixel 1 B mask + 3 B image + 3 B reference in, 3 B image + 1 B blurred mask out = 11 B.
//
// blur_composite_kernel: a CTA owns a TW x TH output tile.  The mask tile plus a halo of 3*(radius+1) pixels per
// side is staged as bytes in shared memory (only in-image pixels exist: every read clamps its coordinate to the
// image, which is Pillow's edge replication), then 3 horizontal and 3 vertical passes ping-pong between two
// shared buffers, each pass shrinking the valid rectangle by radius+1; the blurred tile is written out and used
// at once for the composite of the tile's pixels.
#include <math.h>

#include <vector>

#include "sdof_common.cuh"

namespace sdof {

struct BoxParams {
  int radius;       // integer part of the box radius
  unsigned ww, fw;  // 8.24 weights of the inner box pixels and of the two outer pixels
};

// BoxBlur.c::_gaussian_blur_radius (float variables, double sqrt/floor) + ImagingHorizontalBoxBlur's weights
static BoxParams gaussian_box_params(float radius, int passes) {
  float sigma2 = radius * radius / passes;
  float L = (float)sqrt(12.0 * sigma2 + 1.0);
  float l = (float)floor((L - 1.0) / 2.0);
  float a = (2 * l + 1) * (l * (l + 1) - 3 * sigma2);
  a /= 6 * (sigma2 - (l + 1) * (l + 1));
  const float fr = l + a;
  BoxParams p;
  p.radius = (int)fr;
  p.ww = (unsigned)((float)(1u << 24) / (fr * 2 + 1));
  p.fw = ((1u << 24) - (unsigned)(p.radius * 2 + 1) * p.ww) / 2;
  return p;
}

constexpr int kBlurThreads = 256;

// One extended-box pass along x (kVert = false) or y (kVert = true) over the rectangle [x0,x1) x [y0,y1) of the
// staged region (region coordinates; gx0/gy0 = image coordinates of region (0,0); pitch = region row pitch).
// A thread owns a run of kBoxRun outputs along the pass direction and slides the box sum: two shared-memory loads per
// output instead of 2*radius+3.  Lanes map to consecutive rows (horizontal pass) / columns (vertical pass).
constexpr int kBoxRun = 16;

template <bool kVert>
__device__ __forceinline__ void box_pass(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int pitch,
                                         int x0, int x1, int y0, int y1, int gx0, int gy0, int W, int H, BoxParams bp) {
  const int r = bp.radius;
  // along = pass direction, across = the other one
  const int a0 = kVert ? y0 : x0, a1 = kVert ? y1 : x1;
  const int c0 = kVert ? x0 : y0, c1 = kVert ? x1 : y1;
  const int ga0 = kVert ? gy0 : gx0;            // image coordinate of region 0 along the pass
  const int gmax = (kVert ? H : W) - 1;
  const int step = kVert ? pitch : 1;           // byte step along the pass
  const int nacross = c1 - c0;
  const int nruns = (a1 - a0 + kBoxRun - 1) / kBoxRun;
  for (int i = threadIdx.x; i < nacross * nruns; i += kBlurThreads) {
    const int run = i / nacross, c = c0 + (i - run * nacross);
    const int s0 = a0 + run * kBoxRun, s1 = min(s0 + kBoxRun, a1);
    // line(k) = value at IMAGE coordinate k along the pass (clamped = edge replication), on this thread's row / column
    const unsigned char* base = in + (kVert ? c : c * pitch) - ga0 * step;
#define SDOF_LINE(k) ((unsigned)base[min(max((k), 0), gmax) * step])
    int g = ga0 + s0;
    unsigned acc = 0;
    for (int d = -r; d <= r; ++d) acc += SDOF_LINE(g + d);
    unsigned left = SDOF_LINE(g - r - 1);
    unsigned char* o = out + (kVert ? s0 * pitch + c : c * pitch + s0);
    for (int sidx = s0; sidx < s1; ++sidx, ++g, o += step) {
      const unsigned right = SDOF_LINE(g + r + 1);
      const unsigned bulk = acc * bp.ww + (left + right) * bp.fw;  // UINT32 arithmetic as in Pillow
      *o = (unsigned char)((bulk + (1u << 23)) >> 24);
      left = SDOF_LINE(g - r);
      acc += right - left;
    }
#undef SDOF_LINE
  }
}

__device__ __forceinline__ unsigned div255(unsigned a) {
  const unsigned t = a + 128;
  return ((t >> 8) + t) >> 8;
}

// mask [B,H,W]; image, reference, out [B,H,W,C] (image/reference/out may be NULL: blur only); blurred [B,H,W].
__global__ void __launch_bounds__(kBlurThreads) blur_composite_kernel(const unsigned char* __restrict__ mask, const unsigned char* __restrict__ image,
                                                                      const unsigned char* __restrict__ reference, int H, int W, int C,
                                                                      BoxParams bp, int passes, int TW, int TH, int pitch, int vec_ok,
                                                                      unsigned char* __restrict__ blurred, unsigned char* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char bl_smem[];
  const int halo = passes * (bp.radius + 1);
  const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH, b = blockIdx.z;
  // region = tile + halo, clipped to the image (coordinates outside are never read: reads clamp to the image)
  const int gx0 = max(tx0 - halo, 0), gy0 = max(ty0 - halo, 0);
  const int gx1 = min(tx0 + TW + halo, W), gy1 = min(ty0 + TH + halo, H);
  const int rw = gx1 - gx0, rh = gy1 - gy0;
  unsigned char* buf0 = bl_smem;
  unsigned char* buf1 = bl_smem + (size_t)pitch * (TH + 2 * halo);
  const unsigned char* mb = mask + (int64_t)b * H * W;
  for (int ry = threadIdx.x >> 5; ry < rh; ry += kBlurThreads / 32) {  // a warp per staged row: coalesced byte loads
    const unsigned char* srow = mb + (int64_t)(gy0 + ry) * W + gx0;
    for (int rx = threadIdx.x & 31; rx < rw; rx += 32) buf0[ry * pitch + rx] = srow[rx];
  }
  __syncthreads();
  unsigned char* cur = buf0;
  unsigned char* nxt = buf1;
  if (bp.radius != 0 || bp.fw != 0) {
    // horizontal passes: after pass p the columns within p*(radius+1) of a clipped (non-image) region edge are stale
    for (int p = 1; p <= passes; ++p) {
      const int m = p * (bp.radius + 1);
      const int x0 = (gx0 == 0) ? 0 : m, x1 = (gx1 == W) ? rw : rw - m;
      box_pass<false>(cur, nxt, pitch, x0, x1, 0, rh, gx0, gy0, W, H, bp);
      __syncthreads();
      unsigned char* t = cur; cur = nxt; nxt = t;
    }
    const int xa = (gx0 == 0) ? 0 : halo, xb = (gx1 == W) ? rw : rw - halo;
    for (int p = 1; p <= passes; ++p) {
      const int m = p * (bp.radius + 1);
      const int y0 = (gy0 == 0) ? 0 : m, y1 = (gy1 == H) ? rh : rh - m;
      box_pass<true>(cur, nxt, pitch, xa, xb, y0, y1, gx0, gy0, W, H, bp);
      __syncthreads();
      unsigned char* t = cur; cur = nxt; nxt = t;
    }
  }
  // write the blurred tile and composite
  const int ox = tx0 - gx0, oy = ty0 - gy0;
  const int tw = min(TW, W - tx0), th = min(TH, H - ty0);
  if (vec_ok && C == 3 && (tw & 3) == 0) {
    // 4 pixels per thread: 12 image + 12 reference bytes as aligned words (W % 4 == 0, TW % 4 == 0, aligned pointers)
    const int tw4 = tw >> 2;
    for (int i = threadIdx.x; i < tw4 * th; i += kBlurThreads) {
      const int ly = i / tw4, lx = (i - ly * tw4) * 4;
      const unsigned char* mrow = cur + (oy + ly) * pitch + ox + lx;
      const unsigned m[4] = {mrow[0], mrow[1], mrow[2], mrow[3]};
      const int64_t p = ((int64_t)b * H + ty0 + ly) * W + tx0 + lx;
      if (blurred) *reinterpret_cast<unsigned*>(blurred + p) = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
      if (out) {
        const unsigned* iw = reinterpret_cast<const unsigned*>(image + p * 3);
        const unsigned* rw_ = reinterpret_cast<const unsigned*>(reference + p * 3);
        unsigned* ow = reinterpret_cast<unsigned*>(out + p * 3);
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          const unsigned iv = __ldcs(iw + w), rv = __ldcs(rw_ + w);
          unsigned res = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const unsigned mm = m[(w * 4 + k) / 3];
            res |= div255(((iv >> (8 * k)) & 0xff) * (255u - mm) + ((rv >> (8 * k)) & 0xff) * mm) << (8 * k);
          }
          __stcs(ow + w, res);
        }
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < tw * th; i += kBlurThreads) {
    const int ly = i / tw, lx = i - ly * tw;
    const unsigned m = cur[(oy + ly) * pitch + ox + lx];
    const int64_t p = ((int64_t)b * H + ty0 + ly) * W + tx0 + lx;
    if (blurred) blurred[p] = (unsigned char)m;
    if (out) {
      for (int c = 0; c < C; ++c)
        out[p * C + c] = (unsigned char)div255((unsigned)image[p * C + c] * (255u - m) + (unsigned)reference[p * C + c] * m);
    }
  }
}

// ---- Image.resize (BICUBIC, antialias): separable integer resampling -------------------------------------------------
// Resample.c::precompute_coeffs + normalize_coeffs_8bpc, in double exactly like Pillow.
static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

struct ResampleTable {
  int ksize = 0;
  std::vector<int> bounds;  // [out][2]: xmin, count
  std::vector<int> coeffs;  // [out][ksize]
};

static ResampleTable precompute_coeffs(int in_size, int out_size) {
  ResampleTable t;
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  t.ksize = (int)ceil(support) * 2 + 1;
  t.bounds.assign((size_t)out_size * 2, 0);
  t.coeffs.assign((size_t)out_size * t.ksize, 0);
  std::vector<double> k(t.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      t.coeffs[(size_t)xx * t.ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << 22)) : (int)(0.5 + k[x] * (1 << 22));
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
  }
  return t;
}

__device__ __forceinline__ unsigned char clip8_q22(int v) {
  v >>= 22;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// src [B,H,W] -> dst [B,H,ow]
__global__ void __launch_bounds__(256) resample_h_kernel(const unsigned char* __restrict__ src, const int* __restrict__ bounds,
                                                         const int* __restrict__ coeffs, int ksize, int64_t rows, int W, int ow,
                                                         unsigned char* __restrict__ dst) {
  const int64_t total = rows * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / ow;
    const int xx = (int)(i - row * ow);
    const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
    const unsigned char* s = src + row * W + x0;
    const int* k = coeffs + (int64_t)xx * ksize;
    int ss = 1 << 21;
    for (int x = 0; x < n; ++x) ss += (int)s[x] * k[x];
    dst[i] = clip8_q22(ss);
  }
}

// src [B,H,ow] -> dst [B,oh,ow]; latmask (optional) [B,4,oh,ow] = around(dst / 255) = (dst >= 128)
__global__ void __launch_bounds__(256) resample_v_kernel(const unsigned char* __restrict__ src, const int* __restrict__ bounds,
                                                         const int* __restrict__ coeffs, int ksize, int B, int H, int oh, int ow,
                                                         unsigned char* __restrict__ dst, float* __restrict__ latmask) {
  const int64_t total = (int64_t)B * oh * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % ow);
    const int yy = (int)((i / ow) % oh);
    const int b = (int)(i / ((int64_t)ow * oh));
    const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
    const unsigned char* s = src + ((int64_t)b * H + y0) * ow + xx;
    const int* k = coeffs + (int64_t)yy * ksize;
    int ss = 1 << 21;
    for (int y = 0; y < n; ++y) ss += (int)s[(int64_t)y * ow] * k[y];
    const unsigned char v = clip8_q22(ss);
    if (dst) dst[i] = v;
    if (latmask) {
      const float m = v >= 128 ? 1.f : 0.f;
      for (int c = 0; c < 4; ++c) latmask[(((int64_t)b * 4 + c) * oh + yy) * ow + xx] = m;
    }
  }
}

}  // namespace sdof

extern "C" {

int sdof_mask_blur_composite(const uint8_t* mask, const uint8_t* image, const uint8_t* reference, int B, int H, int W, int C,
                             float mask_blur, uint8_t* blurred, uint8_t* out, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(mask && (blurred || out), "sdof_mask_blur_composite: NULL pointer");
  SDOF_REQUIRE(!out || (image && reference), "sdof_mask_blur_composite: image and reference are required with out");
  SDOF_REQUIRE(B >= 0 && B <= 65535 && H >= 1 && W >= 1 && C >= 1 && C <= 4, "sdof_mask_blur_composite: bad sizes B=%d H=%d W=%d C=%d", B, H, W, C);
  SDOF_REQUIRE(mask_blur >= 0.f && mask_blur <= 64.f, "sdof_mask_blur_composite: mask_blur must be in [0, 64]");
  if (B == 0) return SDOF_OK;
  const int passes = 3;
  BoxParams bp = gaussian_box_params(mask_blur, passes);
  if (mask_blur == 0.f) { bp.radius = 0; bp.ww = 1u << 24; bp.fw = 0; }  // Pillow skips the blur for radius 0
  const int halo = passes * (bp.radius + 1);
  // tile: larger for larger halos so that the staged area stays a small multiple of the tile
  int TW = 64, TH = 32;
  if (halo > 24) { TW = 128; TH = 64; }
  if (halo > 64) { TW = 192; TH = 96; }
  const int pitch = (TW + 2 * halo + 3) & ~3;
  const size_t smem = 2 * (size_t)pitch * (TH + 2 * halo);
  SDOF_REQUIRE(smem <= 220 * 1024, "sdof_mask_blur_composite: mask_blur %.2f needs %zu bytes of shared memory", mask_blur, smem);
  SDOF_CUDA(cudaFuncSetAttribute(blur_composite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(W, TW), ceil_div(H, TH), B);
  const int vec_ok = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(reference) |
                                       reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(blurred)) & 3) == 0;
  blur_composite_kernel<<<grid, kBlurThreads, smem, as_stream(stream)>>>(mask, image, reference, H, W, C, bp, passes, TW, TH, pitch, vec_ok,
                                                                        blurred, out);
  SDOF_LAUNCH_CHECK("blur_composite_kernel");
  return SDOF_OK;
}

// Host-side tables, exported so the CPU test-suite can pin them against Pillow's without a GPU.
int sdof_box_blur_params(float mask_blur, int32_t* out3) {
  using namespace sdof;
  SDOF_REQUIRE(out3 && mask_blur >= 0.f, "sdof_box_blur_params: bad arguments");
  const BoxParams bp = gaussian_box_params(mask_blur, 3);
  out3[0] = bp.radius;
  out3[1] = (int32_t)bp.ww;
  out3[2] = (int32_t)bp.fw;
  return SDOF_OK;
}

int sdof_resample_table(int in_size, int out_size, int32_t* ksize, int32_t* bounds, int32_t* coeffs, int64_t coeffs_cap) {
  using namespace sdof;
  SDOF_REQUIRE(in_size >= 1 && out_size >= 1 && ksize && bounds && coeffs, "sdof_resample_table: bad arguments");
  const ResampleTable t = precompute_coeffs(in_size, out_size);
  SDOF_REQUIRE((int64_t)t.coeffs.size() <= coeffs_cap, "sdof_resample_table: coeffs_cap %lld < %zu", (long long)coeffs_cap, t.coeffs.size());
  *ksize = t.ksize;
  for (size_t i = 0; i < t.bounds.size(); ++i) bounds[i] = t.bounds[i];
  for (size_t i = 0; i < t.coeffs.size(); ++i) coeffs[i] = t.coeffs[i];
  return SDOF_OK;
}

int64_t sdof_resize_bicubic_workspace_bytes(int B, int H, int W, int oh, int ow) {
  if (B < 0 || H < 1 || W < 1 || oh < 1 || ow < 1) return -1;
  const sdof::ResampleTable th = sdof::precompute_coeffs(W, ow), tv = sdof::precompute_coeffs(H, oh);
  const int64_t ints = (int64_t)ow * (2 + th.ksize) + (int64_t)oh * (2 + tv.ksize);
  return ints * 4 + (((int64_t)B * H * ow + 15) & ~15LL) + 64;
}

int sdof_resize_bicubic_u8(const uint8_t* src, int B, int H, int W, int oh, int ow, uint8_t* dst, float* latmask, void* workspace,
                           int64_t workspace_bytes, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(src && (dst || latmask) && workspace, "sdof_resize_bicubic_u8: NULL pointer");
  SDOF_REQUIRE(B >= 0 && H >= 1 && W >= 1 && oh >= 1 && ow >= 1, "sdof_resize_bicubic_u8: bad sizes");
  SDOF_REQUIRE(workspace_bytes >= sdof_resize_bicubic_workspace_bytes(B, H, W, oh, ow), "sdof_resize_bicubic_u8: workspace too small");
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sdof_resize_bicubic_u8: workspace must be 16-byte aligned");
  if (B == 0) return SDOF_OK;
  const ResampleTable th = precompute_coeffs(W, ow), tv = precompute_coeffs(H, oh);
  // workspace: [tmp B*H*ow bytes, padded][h bounds][h coeffs][v bounds][v coeffs]
  uint8_t* tmp = static_cast<uint8_t*>(workspace);
  int* tab = reinterpret_cast<int*>(tmp + (((int64_t)B * H * ow + 15) & ~15LL));
  int* hb = tab;
  int* hc = hb + 2 * ow;
  int* vb = hc + (int64_t)ow * th.ksize;
  int* vc = vb + 2 * oh;
  cudaStream_t st = as_stream(stream);
  // the tables are a few KB: synchronous-with-respect-to-host copies from pageable memory, ordered on the stream
  SDOF_CUDA(cudaMemcpyAsync(hb, th.bounds.data(), th.bounds.size() * 4, cudaMemcpyHostToDevice, st));
  SDOF_CUDA(cudaMemcpyAsync(hc, th.coeffs.data(), th.coeffs.size() * 4, cudaMemcpyHostToDevice, st));
  SDOF_CUDA(cudaMemcpyAsync(vb, tv.bounds.data(), tv.bounds.size() * 4, cudaMemcpyHostToDevice, st));
  SDOF_CUDA(cudaMemcpyAsync(vc, tv.coeffs.data(), tv.coeffs.size() * 4, cudaMemcpyHostToDevice, st));
  SDOF_CUDA(cudaStreamSynchronize(st));  // the host vectors die at return
  const int64_t rows = (int64_t)B * H;
  resample_h_kernel<<<grid_for(rows * ow, 256, 8), 256, 0, st>>>(src, hb, hc, th.ksize, rows, W, ow, tmp);
  SDOF_LAUNCH_CHECK("resample_h_kernel");
  resample_v_kernel<<<grid_for((int64_t)B * oh * ow, 256, 8), 256, 0, st>>>(tmp, vb, vc, tv.ksize, B, H, oh, ow, dst, latmask);
  SDOF_LAUNCH_CHECK("resample_v_kernel");
  return SDOF_OK;
}

}  // extern "C"
