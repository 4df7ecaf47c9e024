// Confidence post-processing, mask generation and composites.
// All HBM-bound byte/float work: grid-stride element-wise kernels plus the
// shared-memory tile dilation of dilate.cuh.
#include "dilate.cuh"
#include "warp.cuh"

namespace sdof {

// ------------------------------------------------------------------------------- M1
// softmax / log_softmax over K, component 0 (pdcnet_of.py:72-74).
__global__ void __launch_bounds__(256) confidence_softmax_kernel(const float* __restrict__ wm, int K, int64_t HW,
                                                                 int64_t total, float* __restrict__ conf,
                                                                 float* __restrict__ logconf) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, p = i - b * HW;
    const float* w = wm + b * K * HW + p;
    float m = w[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, w[k * HW]);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = __fadd_rn(s, expf(__fsub_rn(w[k * HW], m)));
    const float z0 = __fsub_rn(w[0], m);
    if (conf) conf[i] = __fdiv_rn(expf(z0), s);
    if (logconf) logconf[i] = __fsub_rn(z0, logf(s));
  }
}

// ------------------------------------------------------------------------------- M2
// of_calc's travel distance (ofgen_pixel_inpaint.py:105-116): the displacement is pushed
// through float32(float64 grid + flow) and back (float32(float64(map) - grid)).
__global__ void __launch_bounds__(256) travel_distance_kernel(const float* __restrict__ flow,
                                                              const float* __restrict__ conf, int H, int W,
                                                              int64_t total, float thres, float* __restrict__ v) {
  const int64_t hw = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    decompose_pixel(i, hw, W, b, y, x);
    const float2 f = *reinterpret_cast<const float2*>(flow + i * 2);
    // float32(float64 grid + flow) and back: single fp32 adds give the same bits (see warp.cuh map_coord;
    // the subtraction of an integer grid value from its own rounded sum is exact or single-rounded)
    const float mx = __fadd_rn((float)x, f.x);
    const float my = __fadd_rn((float)y, f.y);
    const float dx = __fsub_rn(mx, (float)x);
    const float dy = __fsub_rn(my, (float)y);
    const float r = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    v[i] = (conf[i] < thres) ? 0.f : r;
  }
}

// ------------------------------------------------------------------------------- dilation family
enum { kSrcU8 = 0, kSrcU8Inverted = 1, kSrcConfBelow = 2 };

template <int kSrc, bool kOrWith>
__global__ void __launch_bounds__(kDilThreads) dilate_kernel(const void* __restrict__ src_, float thres,
                                                             const unsigned char* __restrict__ or_with, int H, int W,
                                                             EllipseRows e, unsigned char* __restrict__ dst) {
  extern __shared__ __align__(16) unsigned char tile[];
  const int r = e.ksize >> 1;
  const int tx0 = blockIdx.x * kDilTW, ty0 = blockIdx.y * kDilTH;
  const int64_t img = (int64_t)blockIdx.z * H * W;
  if (kSrc == kSrcConfBelow) {
    const float* c = reinterpret_cast<const float*>(src_) + img;
    dil_stage(tile, r, ty0, tx0, H, W, [&](int gy, int gx) -> unsigned char {
      return c[(int64_t)gy * W + gx] < thres ? 255 : 0;
    });
  } else {
    const unsigned char* s = reinterpret_cast<const unsigned char*>(src_) + img;
    dil_stage(tile, r, ty0, tx0, H, W, [&](int gy, int gx) -> unsigned char {
      const unsigned char v = s[(int64_t)gy * W + gx];
      return kSrc == kSrcU8Inverted ? (unsigned char)(255 - v) : v;
    });
  }
  __syncthreads();
  const int lx = (threadIdx.x % (kDilTW / 4)) * 4, ly = threadIdx.x / (kDilTW / 4);
  const int gx = tx0 + lx, gy = ty0 + ly;
  if (gy >= H || gx >= W) return;
  unsigned out = dil_apply4(tile, e, lx, ly);
  unsigned char* o = dst + img + (int64_t)gy * W + gx;
  const unsigned char* ow = kOrWith ? or_with + img + (int64_t)gy * W + gx : nullptr;
  if (gx + 4 <= W && ((reinterpret_cast<uintptr_t>(o) & 3) == 0) &&
      (!kOrWith || (reinterpret_cast<uintptr_t>(ow) & 3) == 0)) {
    if (kOrWith) out |= *reinterpret_cast<const unsigned*>(ow);
    *reinterpret_cast<unsigned*>(o) = out;
  } else {
    for (int i = 0; i < 4 && gx + i < W; ++i) {
      unsigned char b = (unsigned char)((out >> (8 * i)) & 0xff);
      if (kOrWith) b |= ow[i];
      o[i] = b;
    }
  }
}

// log_conf[conf < thres] = 0 (generate_mask's in-place side effect)
__global__ void __launch_bounds__(256) reset_log_conf_kernel(const float* __restrict__ conf, float thres, int64_t total,
                                                             float* __restrict__ log_conf) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    if (conf[i] < thres) log_conf[i] = 0.f;
}

// ------------------------------------------------------------------------------- M6 (first stage)
// 255 * (RGB2GRAY(|Laplacian(img)| mod 256) > 20): 3x3 [0 1 0;1 -4 1;0 1 0], BORDER_REFLECT_101,
// the float64->uint8 cast wraps, gray = (9798 c0 + 19235 c1 + 3735 c2 + 2^14) >> 15.
__global__ void __launch_bounds__(256) laplacian_edges_kernel(const unsigned char* __restrict__ image, int H, int W,
                                                              int64_t total, unsigned char* __restrict__ out) {
  const int64_t hw = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    decompose_pixel(i, hw, W, b, y, x);
    const unsigned char* img = image + (int64_t)b * hw * 3;
    // reflect-101: -1 -> 1, H -> H-2 (a 1-pixel-wide image reflects onto itself)
    const int yu = y > 0 ? y - 1 : (H > 1 ? 1 : 0), yd = y < H - 1 ? y + 1 : (H > 1 ? H - 2 : 0);
    const int xl = x > 0 ? x - 1 : (W > 1 ? 1 : 0), xr = x < W - 1 ? x + 1 : (W > 1 ? W - 2 : 0);
    const unsigned char* pc = img + ((int64_t)y * W + x) * 3;
    const unsigned char* pu = img + ((int64_t)yu * W + x) * 3;
    const unsigned char* pd = img + ((int64_t)yd * W + x) * 3;
    const unsigned char* pl = img + ((int64_t)y * W + xl) * 3;
    const unsigned char* pr = img + ((int64_t)y * W + xr) * 3;
    int g = 0;
    const int coef[3] = {9798, 19235, 3735};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int lap = (int)pu[c] + (int)pd[c] + (int)pl[c] + (int)pr[c] - 4 * (int)pc[c];
      lap = (lap < 0 ? -lap : lap) & 0xff;
      g += coef[c] * lap;
    }
    g = (g + 16384) >> 15;
    out[i] = g > 20 ? 255 : 0;
  }
}

// ------------------------------------------------------------------------------- M4
// mix_propagated_ai_frame (ofgen_pixel_inpaint.py:251-260): fp32 blend, separate
// multiply/add roundings like numpy, clip, truncating cast.
__global__ void __launch_bounds__(256) mix_kernel(const unsigned char* __restrict__ raw,
                                                  const unsigned char* __restrict__ warped,
                                                  const unsigned char* __restrict__ mask, int C, int64_t npix,
                                                  float w_keep, float w_inpaint, unsigned char* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const float w = mask[p] <= 127 ? w_keep : w_inpaint;
    const float omw = __fsub_rn(1.f, w);
    for (int c = 0; c < C; ++c) {
      float v = __fadd_rn(__fmul_rn((float)raw[p * C + c], omw), __fmul_rn((float)warped[p * C + c], w));
      v = fminf(fmaxf(v, 0.f), 255.f);
      out[p * C + c] = (unsigned char)(int)v;
    }
  }
}

// ------------------------------------------------------------------------------- M5 select
__global__ void __launch_bounds__(256) merge_select_kernel(const unsigned char* __restrict__ base,
                                                           const unsigned char* __restrict__ second,
                                                           const unsigned char* __restrict__ mask, int C, int64_t npix,
                                                           unsigned char* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const bool take = mask[p] == 255;
    for (int c = 0; c < C; ++c) out[p * C + c] = take ? second[p * C + c] : base[p * C + c];
  }
}

// ------------------------------------------------------------------------------- A2
__global__ void __launch_bounds__(256) confidence_sums_kernel(const float* __restrict__ flow_mat, int64_t per_source,
                                                              double* __restrict__ sums) {
  const int s = blockIdx.y;
  const float* base = flow_mat + (int64_t)s * per_source * 3 + 2;
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_source; i += (int64_t)gridDim.x * blockDim.x)
    acc += (double)base[i * 3];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    atomicAdd(sums + s, t);
  }
}

// ------------------------------------------------------------------------------- M5 greedy
// Round state lives in device memory so the n rounds run without host synchronisation:
//   counts[s]  = number of still-uncovered confident pixels of reference s (exact integers,
//                equal to the reference's fp32 sums of {0,1} values)
//   order[i]   = reference chosen in round i (np.argmax: first maximum)
__global__ void __launch_bounds__(256) greedy_binarise_kernel(float* __restrict__ flow_mat, int n, int64_t hw,
                                                              float thres, unsigned long long* __restrict__ counts) {
  const int s = blockIdx.y;
  float* conf = flow_mat + (int64_t)s * hw * 3 + 2;
  unsigned cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const bool on = conf[i * 3] > thres;
    conf[i * 3] = on ? 1.f : 0.f;
    cnt += on;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(counts + s, (unsigned long long)cnt);
}

__global__ void greedy_pick_kernel(unsigned long long* __restrict__ counts, int n, int round,
                                   int32_t* __restrict__ order) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int best = 0;
    unsigned long long bv = counts[0];
    for (int s = 1; s < n; ++s)
      if (counts[s] > bv) {
        bv = counts[s];
        best = s;
      }
    order[round] = best;
    for (int s = 0; s < n; ++s) counts[s] = 0;  // re-accumulated by the apply kernel
  }
}

constexpr int kGreedyMaxRefs = 64;

__global__ void __launch_bounds__(256) greedy_apply_kernel(const int16_t* __restrict__ tab, float* __restrict__ flow_mat,
                                                           const unsigned char* __restrict__ ai_frames,
                                                           const unsigned char* __restrict__ frames_end, int n, int H,
                                                           int W, int round, const int32_t* __restrict__ order,
                                                           unsigned char* __restrict__ ret,
                                                           unsigned char* __restrict__ mask,
                                                           unsigned long long* __restrict__ counts) {
  const int64_t hw = (int64_t)H * W;
  const int ref = order[round];
  const float* fref = flow_mat + (int64_t)ref * hw * 3;
  const unsigned char* frame = ai_frames + (int64_t)ref * hw * 3;
  __shared__ unsigned s_cnt[kGreedyMaxRefs];
  for (int s = threadIdx.x; s < n; s += blockDim.x) s_cnt[s] = 0;
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < hw; p += (int64_t)gridDim.x * blockDim.x) {
    const float last = fref[p * 3 + 2];
    const unsigned char cur = (unsigned char)(int)(last * 255.f);  // 0 or 255
    mask[p] = round == 0 ? cur : (unsigned char)(mask[p] | cur);
    if (round == 0 || cur == 255) {
      const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
      const FixedCoord fc = fixed_coord(map_coord(x, fref[p * 3], 1.f), map_coord(y, fref[p * 3 + 1], 1.f));
      const unsigned px = cubic_u8_c3(tab, frame, frames_end, H, W, fc);
      ret[p * 3] = (unsigned char)(px & 0xff);
      ret[p * 3 + 1] = (unsigned char)((px >> 8) & 0xff);
      ret[p * 3 + 2] = (unsigned char)((px >> 16) & 0xff);
    }
    // subtract the covered pixels from every reference, clip to [0,1], count what is left
    for (int s = 0; s < n; ++s) {
      float* c = flow_mat + ((int64_t)s * hw + p) * 3 + 2;
      float v = __fsub_rn(*c, last);
      v = fminf(fmaxf(v, 0.f), 1.f);
      *c = v;
      if (v > 0.f) atomicAdd(&s_cnt[s], 1u);
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < n; s += blockDim.x)
    if (s_cnt[s]) atomicAdd(counts + s, (unsigned long long)s_cnt[s]);
}

static int check_img(const char* name, int B, int H, int W) {
  SDOF_REQUIRE(B >= 0 && H >= 1 && W >= 1, "%s: bad sizes B=%d H=%d W=%d", name, B, H, W);
  return SDOF_OK;
}

template <int kSrc, bool kOrWith>
static int launch_dilate(const char* name, const void* src, float thres, const unsigned char* or_with, int B, int H,
                         int W, int ksize, unsigned char* dst, sdof_stream_t stream) {
  EllipseRows e;
  if (make_ellipse(ksize, &e)) return fail(SDOF_ERR_INVALID, "%s: ksize must be odd in [1,%d], got %d", name, kDilMaxK, ksize);
  if (B == 0) return SDOF_OK;
  SDOF_REQUIRE(B <= 65535, "%s: B > 65535 not supported", name);
  const int r = ksize >> 1;
  dim3 grid(ceil_div(W, kDilTW), ceil_div(H, kDilTH), B);
  dilate_kernel<kSrc, kOrWith><<<grid, kDilThreads, dil_smem_bytes(r), as_stream(stream)>>>(src, thres, or_with, H, W, e, dst);
  SDOF_LAUNCH_CHECK(name);
  return SDOF_OK;
}

}  // namespace sdof

extern "C" {

int sdof_confidence_softmax(const float* weight_map, int B, int K, int H, int W, float* conf, float* logconf,
                            sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(weight_map && (conf || logconf), "sdof_confidence_softmax: NULL pointer");
  SDOF_REQUIRE(K >= 1, "sdof_confidence_softmax: K must be >= 1");
  int rc = check_img("sdof_confidence_softmax", B, H, W);
  if (rc) return rc;
  const int64_t hw = (int64_t)H * W, total = hw * B;
  if (total == 0) return SDOF_OK;
  confidence_softmax_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(weight_map, K, hw, total, conf, logconf);
  SDOF_LAUNCH_CHECK("confidence_softmax_kernel");
  return SDOF_OK;
}

int sdof_travel_distance(const float* flow, const float* conf, int B, int H, int W, float conf_thres, float* v,
                         sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(flow && conf && v, "sdof_travel_distance: NULL pointer");
  int rc = check_img("sdof_travel_distance", B, H, W);
  if (rc) return rc;
  const int64_t total = (int64_t)B * H * W;
  if (total == 0) return SDOF_OK;
  travel_distance_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(flow, conf, H, W, total, conf_thres, v);
  SDOF_LAUNCH_CHECK("travel_distance_kernel");
  return SDOF_OK;
}

int sdof_generate_mask(const float* conf, float* log_conf, int B, int H, int W, float thres, int ksize, uint8_t* mask,
                       sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(conf && mask, "sdof_generate_mask: NULL pointer");
  int rc = check_img("sdof_generate_mask", B, H, W);
  if (rc) return rc;
  rc = launch_dilate<kSrcConfBelow, false>("sdof_generate_mask", conf, thres, nullptr, B, H, W, ksize, mask, stream);
  if (rc) return rc;
  const int64_t total = (int64_t)B * H * W;
  if (log_conf && total) {
    reset_log_conf_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(conf, thres, total, log_conf);
    SDOF_LAUNCH_CHECK("reset_log_conf_kernel");
  }
  return SDOF_OK;
}

int sdof_dilate_ellipse_u8(const uint8_t* src, int B, int H, int W, int ksize, int invert, uint8_t* dst,
                           sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(src && dst, "sdof_dilate_ellipse_u8: NULL pointer");
  SDOF_REQUIRE(src != dst, "sdof_dilate_ellipse_u8: in-place dilation is not supported");
  int rc = check_img("sdof_dilate_ellipse_u8", B, H, W);
  if (rc) return rc;
  if (invert) return launch_dilate<kSrcU8Inverted, false>("sdof_dilate_ellipse_u8", src, 0.f, nullptr, B, H, W, ksize, dst, stream);
  return launch_dilate<kSrcU8, false>("sdof_dilate_ellipse_u8", src, 0.f, nullptr, B, H, W, ksize, dst, stream);
}

int sdof_expand_mask(const uint8_t* mask, const uint8_t* image, int B, int H, int W, int ksize, uint8_t* scratch,
                     uint8_t* out, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(mask && image && scratch && out, "sdof_expand_mask: NULL pointer");
  SDOF_REQUIRE(scratch != out && scratch != mask, "sdof_expand_mask: scratch must not alias mask/out");
  int rc = check_img("sdof_expand_mask", B, H, W);
  if (rc) return rc;
  const int64_t total = (int64_t)B * H * W;
  if (total == 0) return SDOF_OK;
  laplacian_edges_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(image, H, W, total, scratch);
  SDOF_LAUNCH_CHECK("laplacian_edges_kernel");
  return launch_dilate<kSrcU8, true>("sdof_expand_mask", scratch, 0.f, mask, B, H, W, ksize, out, stream);
}

int sdof_mix_propagated(const uint8_t* raw, const uint8_t* warped, const uint8_t* mask, int B, int H, int W, int C,
                        float ppw, uint8_t* out, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(raw && warped && mask && out, "sdof_mix_propagated: NULL pointer");
  SDOF_REQUIRE(C >= 1 && C <= 4, "sdof_mix_propagated: C must be in 1..4");
  int rc = check_img("sdof_mix_propagated", B, H, W);
  if (rc) return rc;
  const int64_t npix = (int64_t)B * H * W;
  if (npix == 0) return SDOF_OK;
  if (ppw < 0.001f) {  // reference returns raw_ai_frame unchanged
    SDOF_CUDA(cudaMemcpyAsync(out, raw, (size_t)npix * C, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return SDOF_OK;
  }
  // weights[mask<=127] = ppw; weights[mask>127] = 1 - ppw, with (1 - ppw) evaluated in
  // double by Python before the float32 store
  const float w_keep = ppw;
  const float w_inpaint = (float)(1.0 - (double)ppw);
  mix_kernel<<<grid_for(npix, 256, 8), 256, 0, as_stream(stream)>>>(raw, warped, mask, C, npix, w_keep, w_inpaint, out);
  SDOF_LAUNCH_CHECK("mix_kernel");
  return SDOF_OK;
}

int sdof_merge_select(const uint8_t* base, const uint8_t* second, const uint8_t* mask, int B, int H, int W, int C,
                      uint8_t* out, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(base && second && mask && out, "sdof_merge_select: NULL pointer");
  SDOF_REQUIRE(C >= 1 && C <= 4, "sdof_merge_select: C must be in 1..4");
  int rc = check_img("sdof_merge_select", B, H, W);
  if (rc) return rc;
  const int64_t npix = (int64_t)B * H * W;
  if (npix == 0) return SDOF_OK;
  merge_select_kernel<<<grid_for(npix, 256, 8), 256, 0, as_stream(stream)>>>(base, second, mask, C, npix, out);
  SDOF_LAUNCH_CHECK("merge_select_kernel");
  return SDOF_OK;
}

int sdof_confidence_sums(const float* flow_mat, int S, int64_t per_source, double* sums, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(flow_mat && sums, "sdof_confidence_sums: NULL pointer");
  SDOF_REQUIRE(S >= 0 && S <= 65535 && per_source >= 0, "sdof_confidence_sums: bad sizes");
  if (S == 0) return SDOF_OK;
  SDOF_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * S, as_stream(stream)));
  if (per_source == 0) return SDOF_OK;
  int gx = grid_for(per_source, 256, 8) / (S < 8 ? S : 8);
  if (gx < 1) gx = 1;
  confidence_sums_kernel<<<dim3(gx, S), 256, 0, as_stream(stream)>>>(flow_mat, per_source, sums);
  SDOF_LAUNCH_CHECK("confidence_sums_kernel");
  return SDOF_OK;
}

int64_t sdof_greedy_workspace_bytes(int n, int H, int W) {
  (void)H;
  (void)W;
  return (int64_t)sizeof(unsigned long long) * (n > 0 ? n : 1);
}

int sdof_greedy_composite(float* flow_mat, const uint8_t* ai_frames, int n, int H, int W, float thres, uint8_t* ret,
                          uint8_t* mask, int32_t* order, void* workspace, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(flow_mat && ai_frames && ret && mask && order && workspace, "sdof_greedy_composite: NULL pointer");
  SDOF_REQUIRE(n >= 1 && n <= kGreedyMaxRefs, "sdof_greedy_composite: n must be in [1,%d], got %d", kGreedyMaxRefs, n);
  int rc = check_img("sdof_greedy_composite", 1, H, W);
  if (rc) return rc;
  SDOF_REQUIRE(W <= 32767 && H <= 32767, "sdof_greedy_composite: frame larger than 32767");
  CubicTables tabs;
  if ((rc = get_cubic_tables(&tabs))) return rc;
  cudaStream_t st = as_stream(stream);
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(workspace);
  const int64_t hw = (int64_t)H * W;
  SDOF_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * n, st));
  int gx = grid_for(hw, 256, 8) / (n < 8 ? n : 8);
  if (gx < 1) gx = 1;
  greedy_binarise_kernel<<<dim3(gx, n), 256, 0, st>>>(flow_mat, n, hw, thres, counts);
  SDOF_LAUNCH_CHECK("greedy_binarise_kernel");
  const bool aligned = (reinterpret_cast<uintptr_t>(ai_frames) & 3) == 0;
  const uint8_t* frames_end = aligned ? ai_frames + (int64_t)n * hw * 3 : ai_frames;
  for (int round = 0; round < n; ++round) {
    greedy_pick_kernel<<<1, 32, 0, st>>>(counts, n, round, order);
    SDOF_LAUNCH_CHECK("greedy_pick_kernel");
    greedy_apply_kernel<<<grid_for(hw, 256, 4), 256, 0, st>>>(tabs.i16, flow_mat, ai_frames, frames_end, n, H, W, round, order,
                                                             ret, mask, counts);
    SDOF_LAUNCH_CHECK("greedy_apply_kernel");
  }
  return SDOF_OK;
}

}  // extern "C"
