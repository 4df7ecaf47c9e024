// Standalone driver for compute-sanitizer (memcheck / racecheck / synccheck): calls the C ABI of libsdof_b200.so directly
// with cudaMalloc'ed buffers, no Python and no torch in the process (round 1's runs died inside the interpreter before
// reaching a kernel, so their "0 errors" was no evidence).  Sizes are small (the tools slow kernels down 10-100x) but
// chosen so that every code path of the hot kernels runs: partial correlation tiles and odd pooled sizes, staged AND
// fallback warp tiles, image borders, hysteresis relaunches, several greedy rounds.
//
//   tools/run_sanitizer.sh            # builds this file against the in-tree library and runs the three tools
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sdof_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)
#define SD(x)                                                                      \
  do {                                                                             \
    int r_ = (x);                                                                  \
    if (r_ != 0) {                                                                 \
      fprintf(stderr, "%s -> %d: %s\n", #x, r_, sdof_last_error());                \
      exit(3);                                                                     \
    }                                                                              \
  } while (0)

static uint32_t g_seed = 12345u;
static float frand() {  // uniform (-1, 1)
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) * (1.0f / 8388608.0f)) - 1.0f;
}
template <typename T>
static T* dalloc(size_t n) {
  void* p = nullptr;
  CK(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
  return static_cast<T*>(p);
}
static float* dfloats(size_t n, float scale) {
  std::vector<float> h(n);
  for (auto& v : h) v = scale * frand();
  float* d = dalloc<float>(n);
  CK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}
static uint8_t* dbytes(size_t n) {
  std::vector<uint8_t> h(n);
  for (auto& v : h) v = (uint8_t)(255.f * 0.5f * (frand() + 1.f));
  uint8_t* d = dalloc<uint8_t>(n);
  CK(cudaMemcpy(d, h.data(), n, cudaMemcpyHostToDevice));
  return d;
}

static void corr_case(int B, int h, int w, int C, int precision, const char* name) {
  const int levels = 4, r = 4;
  sdof_pyramid_layout lay;
  SD(sdof_corr_pyramid_layout((int64_t)B * h * w, h, w, levels, &lay));
  float* f1 = dfloats((size_t)B * h * w * C, 1.f);
  float* f2 = dfloats((size_t)B * h * w * C, 1.f);
  float* pyr = dalloc<float>((size_t)lay.total_floats);
  const int64_t wsb = sdof_corr_volume_workspace_bytes(B, h, w, h, w, C, levels, precision);
  void* ws = wsb ? dalloc<uint8_t>((size_t)wsb) : nullptr;
  SD(sdof_corr_volume_pyramid(f1, f2, B, h, w, h, w, C, levels, precision, pyr, ws, wsb, nullptr));
  // lookup (planar + channels-last) with coordinates that leave the map on every side
  std::vector<float> hc((size_t)B * 2 * h * w), hn((size_t)B * h * w * 2);
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const float cx = x + 6.f * frand(), cy = y + 6.f * frand();
        hc[((size_t)(b * 2 + 0) * h + y) * w + x] = cx;
        hc[((size_t)(b * 2 + 1) * h + y) * w + x] = cy;
        hn[(((size_t)b * h + y) * w + x) * 2 + 0] = cx;
        hn[(((size_t)b * h + y) * w + x) * 2 + 1] = cy;
      }
  float* dc = dalloc<float>(hc.size());
  float* dn = dalloc<float>(hn.size());
  CK(cudaMemcpy(dc, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dn, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice));
  float* out = dalloc<float>((size_t)B * 324 * h * w);
  SD(sdof_corr_lookup(pyr, dc, B, h, w, h, w, levels, r, out, nullptr));
  SD(sdof_corr_lookup_nhwc(pyr, dn, B, h, w, h, w, levels, r, out, nullptr));
  CK(cudaDeviceSynchronize());
  printf("ok corr %-8s B=%d %dx%d C=%d\n", name, B, h, w, C);
  cudaFree(f1); cudaFree(f2); cudaFree(pyr); cudaFree(ws); cudaFree(dc); cudaFree(dn); cudaFree(out);
}

// round-2 path: split operands (shared key-frame target, B2 = 1), fp16-stored pyramid, lookup reading it
static void corr_half_case(int B, int h, int w, int C) {
  const int levels = 4, r = 4;
  sdof_pyramid_layout lay;
  SD(sdof_corr_pyramid_layout_ex((int64_t)B * h * w, h, w, levels, 2, &lay));
  float* f1 = dfloats((size_t)B * h * w * C, 300.f);      // auto-ranging: far from unit scale
  float* key = dfloats((size_t)h * w * C, 1e-3f);
  const int64_t sb = sdof_corr_src_operand_bytes(B, h, w, C), tb = sdof_corr_tgt_operand_bytes(1, h, w, C, levels);
  void* so = dalloc<uint8_t>((size_t)sb);
  void* to = dalloc<uint8_t>((size_t)tb);
  uint16_t* pyr = dalloc<uint16_t>((size_t)lay.total_floats);
  SD(sdof_corr_prepare_tgt(key, 1, h, w, C, levels, SDOF_PREC_FP16, to, tb, nullptr));
  SD(sdof_corr_prepare_src(f1, B, h, w, C, SDOF_PREC_FP16, so, sb, nullptr));
  SD(sdof_corr_pyramid_from_parts(so, to, B, h, w, 1, h, w, C, levels, SDOF_PREC_FP16, 2, pyr, nullptr));
  std::vector<float> hn((size_t)B * h * w * 2);
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        hn[(((size_t)b * h + y) * w + x) * 2 + 0] = x + 7.f * frand();
        hn[(((size_t)b * h + y) * w + x) * 2 + 1] = y + 7.f * frand();
      }
  float* dn = dalloc<float>(hn.size());
  CK(cudaMemcpy(dn, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice));
  float* out = dalloc<float>((size_t)B * 324 * h * w);
  SD(sdof_corr_lookup_ex(pyr, 2, dn, B, h, w, h, w, levels, r, out, 1, nullptr));
  SD(sdof_corr_lookup_ex(pyr, 2, dn, B, h, w, h, w, levels, r, out, 0, nullptr));   // (coords read as planar: any values do)
  CK(cudaDeviceSynchronize());
  printf("ok corr fp16-stored pyramid, shared target B=%d %dx%d C=%d\n", B, h, w, C);
  cudaFree(f1); cudaFree(key); cudaFree(so); cudaFree(to); cudaFree(pyr); cudaFree(dn); cudaFree(out);
}

static void alt_corr_case(int B, int h, int w, int C) {
  float* f1 = dfloats((size_t)B * h * w * C, 1.f);
  float* f2 = dfloats((size_t)B * h * w * C, 1.f);
  std::vector<float> hc((size_t)B * h * w * 2);
  for (int i = 0; i < B * h * w; ++i) {
    hc[2 * i] = (i % w) + 5.f * frand();
    hc[2 * i + 1] = ((i / w) % h) + 5.f * frand();
  }
  float* dc = dalloc<float>(hc.size());
  CK(cudaMemcpy(dc, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
  float* out = dalloc<float>((size_t)B * 81 * h * w);
  SD(sdof_alt_corr_forward(f1, f2, dc, B, h, w, h, w, C, 1, 4, out, nullptr));
  float* pooled = dalloc<float>((size_t)B * (h / 2) * (w / 2) * C);
  SD(sdof_avgpool2_nhwc(f2, B, h, w, C, pooled, nullptr));
  CK(cudaDeviceSynchronize());
  printf("ok alt_corr B=%d %dx%d C=%d\n", B, h, w, C);
  cudaFree(f1); cudaFree(f2); cudaFree(dc); cudaFree(out); cudaFree(pooled);
}

static void warp_case(int B, int H, int W, float noise, float shift, const char* name) {
  uint8_t* src = dbytes((size_t)B * H * W * 3);
  uint8_t* base = dbytes((size_t)B * H * W * 3);
  std::vector<float> hf((size_t)B * H * W * 2);
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        float* f = &hf[(((size_t)b * H + y) * W + x) * 2];
        f[0] = shift + 3.f * sinf(0.02f * y + b) + noise * frand();
        f[1] = -shift + 3.f * cosf(0.015f * x) + noise * frand();
      }
  hf[10] = NAN;           // cv2's out-of-range rule
  hf[20] = 1e9f;
  hf[33] = -INFINITY;
  float* flow = dalloc<float>(hf.size());
  CK(cudaMemcpy(flow, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice));
  uint8_t* dst = dalloc<uint8_t>((size_t)B * H * W * 3);
  int64_t st[2];
  SD(sdof_warp_tile_stats(st, 1));
  SD(sdof_warp_cubic_u8(src, flow, B, 1, H, W, 3, H, W, 1.0f, dst, nullptr));
  SD(sdof_warp_cubic_u8(src, flow, B, 0, H, W, 3, H, W, -1.0f, dst, nullptr));
  SD(sdof_warp_tile_stats(st, 1));
  SD(sdof_warp_bilinear_u8(src, flow, B, 1, H, W, 3, H, W, 1.0f, dst, nullptr));
  float* fsrc = dfloats((size_t)B * H * W * 2, 5.f);
  float* fdst = dalloc<float>((size_t)B * H * W * 2);
  SD(sdof_warp_cubic_f32(fsrc, flow, B, 1, H, W, 2, H, W, 1.0f, fdst, nullptr));
  SD(sdof_warp_bilinear_f32(fsrc, flow, B, 1, H, W, 2, H, W, 1.0f, fdst, nullptr));
  // fused warp + confidence mask + composite
  float* wm = dfloats((size_t)B * 2 * H * W, 3.f);
  uint8_t* mask = dalloc<uint8_t>((size_t)B * H * W);
  SD(sdof_warp_mask_composite(src, base, flow, wm, B, 1, H, W, 0.6f, 7, dst, mask, nullptr));
  CK(cudaDeviceSynchronize());
  printf("ok warp %-8s B=%d %dx%d: %lld staged / %lld fallback tiles\n", name, B, H, W, (long long)st[0], (long long)st[1]);
  cudaFree(src); cudaFree(base); cudaFree(flow); cudaFree(dst); cudaFree(fsrc); cudaFree(fdst); cudaFree(wm); cudaFree(mask);
}

static void mask_case(int B, int H, int W) {
  float* conf = dfloats((size_t)B * H * W, 0.5f);
  float* logc = dfloats((size_t)B * H * W, 1.f);
  uint8_t* mask = dalloc<uint8_t>((size_t)B * H * W);
  uint8_t* m2 = dalloc<uint8_t>((size_t)B * H * W);
  uint8_t* scratch = dalloc<uint8_t>((size_t)B * H * W);
  uint8_t* img = dbytes((size_t)B * H * W * 3);
  uint8_t* img2 = dbytes((size_t)B * H * W * 3);
  uint8_t* out = dalloc<uint8_t>((size_t)B * H * W * 3);
  SD(sdof_generate_mask(conf, logc, B, H, W, 0.1f, 7, mask, nullptr));
  SD(sdof_dilate_ellipse_u8(mask, B, H, W, 15, 1, m2, nullptr));
  SD(sdof_expand_mask(mask, img, B, H, W, 7, scratch, m2, nullptr));
  SD(sdof_mix_propagated(img, img2, mask, B, H, W, 3, 0.3f, out, nullptr));
  SD(sdof_merge_select(img, img2, mask, B, H, W, 3, out, nullptr));
  uint8_t* blurred = dalloc<uint8_t>((size_t)B * H * W);
  SD(sdof_mask_blur_composite(mask, img, img2, B, H, W, 3, 4.0f, blurred, out, nullptr));
  const int oh = H / 8, ow = W / 8;
  const int64_t rb = sdof_resize_bicubic_workspace_bytes(B, H, W, oh, ow);
  void* rws = dalloc<uint8_t>((size_t)rb);
  uint8_t* small = dalloc<uint8_t>((size_t)B * oh * ow);
  float* lat = dalloc<float>((size_t)B * 4 * oh * ow);
  SD(sdof_resize_bicubic_u8(blurred, B, H, W, oh, ow, small, lat, rws, rb, nullptr));
  // greedy composite over n references
  const int n = 4;
  std::vector<float> fm((size_t)n * H * W * 3);
  for (size_t i = 0; i < fm.size(); i += 3) {
    fm[i] = 3.f * frand();
    fm[i + 1] = 3.f * frand();
    fm[i + 2] = 0.5f * (frand() + 1.f);
  }
  float* dfm = dalloc<float>(fm.size());
  CK(cudaMemcpy(dfm, fm.data(), fm.size() * 4, cudaMemcpyHostToDevice));
  uint8_t* frames = dbytes((size_t)n * H * W * 3);
  uint8_t* ret = dalloc<uint8_t>((size_t)H * W * 3);
  uint8_t* gmask = dalloc<uint8_t>((size_t)H * W);
  int32_t* order = dalloc<int32_t>(n);
  void* gws = dalloc<uint8_t>((size_t)sdof_greedy_workspace_bytes(n, H, W));
  double* sums = dalloc<double>(n);
  SD(sdof_confidence_sums(dfm, n, (int64_t)H * W, sums, nullptr));
  SD(sdof_greedy_composite(dfm, frames, n, H, W, 0.55f, ret, gmask, order, gws, nullptr));
  // key-frame detector (Canny with hysteresis relaunches)
  const int64_t eb = sdof_detect_edges_workspace_bytes(H, W);
  void* ews = dalloc<uint8_t>((size_t)eb);
  uint8_t* edges = dalloc<uint8_t>((size_t)H * W);
  SD(sdof_detect_edges(img, H, W, 3, -1, -1, edges, ews, eb, nullptr));
  unsigned long long* dsum = dalloc<unsigned long long>(1);
  SD(sdof_abs_diff_sum_u8(edges, gmask, (int64_t)H * W, dsum, nullptr));
  CK(cudaDeviceSynchronize());
  printf("ok masks / greedy / blur / detector B=%d %dx%d\n", B, H, W);
}

static void glue_case(int B, int h, int w) {
  const int64_t npix = (int64_t)B * h * w;
  float* mask = dfloats((size_t)npix * 576, 2.f);
  float* flow = dfloats((size_t)npix * 2, 3.f);
  float* up = dalloc<float>((size_t)npix * 64 * 2);
  SD(sdof_convex_upsample(mask, nullptr, 0.25f, flow, B, h, w, up, nullptr));
  float* wT = dfloats(7 * 7 * 2 * 128, 0.1f);
  float* bias = dfloats(128, 0.1f);
  float* o128 = dalloc<float>((size_t)npix * 128);
  SD(sdof_conv7x7_c2_relu(flow, wT, bias, o128, B, h, w, nullptr));
  float* x256 = dfloats((size_t)npix * 256, 1.f);
  float* w2 = dfloats(3 * 3 * 2 * 256, 0.05f);
  float* coords = dfloats((size_t)npix * 2, 10.f);
  float* hx = dalloc<float>((size_t)npix * 256);
  float* scratch = dalloc<float>((size_t)npix * 18);
  SD(sdof_flowhead2_update(x256, w2, 0.1f, -0.1f, coords, flow, hx, 256, 254, nullptr, 0, 0, B, h, w, scratch, nullptr));
  uint8_t* img = dbytes((size_t)B * (8 * h - 3) * (8 * w - 5) * 3);
  float* norm = dalloc<float>((size_t)B * 8 * h * 8 * w * 4);
  SD(sdof_normalize_pad_u8_nhwc(img, B, 8 * h - 3, 8 * w - 5, 1, 2, 8 * h, 8 * w, 4, 1, norm, nullptr));
  double* stats = dalloc<double>((size_t)B * 256 * 2);
  CK(cudaMemset(stats, 0, (size_t)B * 256 * 2 * sizeof(double)));
  SD(sdof_instnorm_stats_nhwc(x256, B, (int64_t)h * w, 256, stats, nullptr));
  SD(sdof_instnorm_apply_nhwc(x256, stats, nullptr, x256, B, (int64_t)h * w, 256, 1e-5f, 1, nullptr));
  CK(cudaDeviceSynchronize());
  printf("ok RAFT glue B=%d %dx%d\n", B, h, w);
}

// tcgen05 SepConvGRU kernels (csrc/conv_tc.cu): both passes, a size with partial tiles
static void gru_tc_case(int B, int h, int w) {
  const size_t npix = (size_t)B * h * w;
  std::vector<uint16_t> hz(npix * 256, 0x3400 /* 0.25 in fp16 */);
  uint16_t* hx16 = dalloc<uint16_t>(npix * 256 + 64);
  CK(cudaMemcpy(hx16, hz.data(), npix * 256 * 2, cudaMemcpyHostToDevice));
  std::vector<uint16_t> wz((size_t)384 * 1280, 0x2000), wq((size_t)128 * 640, 0x2000);
  uint16_t* dwz = dalloc<uint16_t>(wz.size());
  uint16_t* dwq = dalloc<uint16_t>(wq.size());
  CK(cudaMemcpy(dwz, wz.data(), wz.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dwq, wq.data(), wq.size() * 2, cudaMemcpyHostToDevice));
  float* zrmap = dfloats(npix * 256, 0.5f);
  float* qmap = dfloats(npix * 128, 0.5f);
  float* hid = dfloats(npix * 128, 0.9f);
  float* z = dalloc<float>(npix * 128);
  float* qx = dalloc<float>(npix * 128);
  uint16_t* rh16 = dalloc<uint16_t>(npix * 128 + 64);
  float* mc = dfloats(npix * 128, 1.f);
  float* mf = dfloats(npix * 128, 1.f);
  float* bias = dfloats(128, 0.1f);
  float* flow = dfloats(npix * 2, 3.f);
  SD(sdof_motion_tail16(mc, mf, bias, flow, (int64_t)npix, hx16, 256, nullptr));
  for (int horizontal = 1; horizontal >= 0; --horizontal) {
    SD(sdof_gru_zr_tc(hx16, dwz, zrmap, hid, B, h, w, horizontal, z, rh16, qx, nullptr));
    SD(sdof_gru_q_tc(rh16, dwq, qmap, qx, z, B, h, w, horizontal, hid, hx16, 256, nullptr));
  }
  CK(cudaDeviceSynchronize());
  printf("ok tcgen05 GRU B=%d %dx%d\n", B, h, w);
}

// fp16 update-loop glue (csrc/raft_glue16.cu), the fp16 lookup output and fp16 instance norm -- all launched with the
// programmatic-dependent-launch attribute, back to back on one stream like an iteration of the update block
static uint16_t* dhalves(size_t n, uint16_t v) {
  std::vector<uint16_t> h(n, v);
  uint16_t* d = dalloc<uint16_t>(n);
  CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  return d;
}
static void glue16_case(int B, int h, int w) {
  const int levels = 4, r = 4;
  const size_t npix = (size_t)B * h * w;
  sdof_pyramid_layout lay;
  SD(sdof_corr_pyramid_layout_ex((int64_t)npix, h, w, levels, 2, &lay));
  uint16_t* pyr = dhalves((size_t)lay.total_floats, 0x3000);
  float* coords = dfloats(npix * 2, 10.f);
  float* coords2 = dalloc<float>(npix * 2);
  uint16_t* i2c = dalloc<uint16_t>(npix * 200);
  uint16_t* corr16 = dalloc<uint16_t>(npix * 328);
  float* flow = dfloats(npix * 2, 3.f);
  float* wT = dfloats(7 * 7 * 2 * 128, 0.1f);
  float* bias = dfloats(128, 0.1f);
  uint16_t* f1 = dalloc<uint16_t>(npix * 128);
  uint16_t* mc = dhalves(npix * 128, 0x3800);
  uint16_t* mf = dhalves(npix * 128, 0x3400);
  uint16_t* hx16 = dalloc<uint16_t>(npix * 256);
  uint16_t* zr16 = dhalves(npix * 384, 0x3000);
  uint16_t* q16 = dhalves(npix * 128, 0x3000);
  uint16_t* rh16 = dalloc<uint16_t>(npix * 128);
  uint16_t* h16 = dalloc<uint16_t>(npix * 128);
  float* zrmap = dfloats(npix * 256, 0.5f);
  float* qmap = dfloats(npix * 128, 0.5f);
  float* hid = dfloats(npix * 128, 0.9f);
  uint16_t* x16 = dhalves(npix * 256, 0x3400);
  float* w2 = dfloats(3 * 3 * 2 * 256, 0.05f);
  float* scratch = dalloc<float>(npix * 18);
  for (int it = 0; it < 2; ++it) {
    SD(sdof_corr_lookup_h(pyr, 2, coords, B, h, w, h, w, levels, r, corr16, 328, nullptr));
    SD(sdof_conv7x7_c2_relu_h(flow, wT, bias, f1, B, h, w, nullptr));
    SD(sdof_motion_tail16_h(mc, mf, bias, flow, (int64_t)npix, hx16, 256, nullptr));
    SD(sdof_gru_rh_h(zr16, 384, zrmap, hid, rh16, (int64_t)npix, nullptr));
    SD(sdof_gru_update_h(zr16, zrmap, q16, qmap, hid, hx16, 256, h16, (int64_t)npix, nullptr));
    SD(sdof_flowhead2_taps_h(x16, w2, (int64_t)npix, scratch, nullptr));
    SD(sdof_flowhead2_gather_update(scratch, 0.1f, -0.1f, coords, flow, nullptr, 0, 0, B, h, w, nullptr));
    // the deferred form of the same update: applied inside the next lookup / convf1
    SD(sdof_corr_lookup_gather_h(pyr, 2, coords, it ? scratch : nullptr, 0.1f, -0.1f, coords2, flow, B, h, w, h, w, levels, r, corr16, 328, nullptr));
    SD(sdof_conv7x7_c2_relu_coords_h(coords, it ? scratch : nullptr, 0.1f, -0.1f, wT, bias, f1, B, h, w, nullptr));
    SD(sdof_flow_im2col7_h(coords, it ? scratch : nullptr, 0.1f, -0.1f, i2c, 200, B, h, w, nullptr));
  }
  double* stats = dalloc<double>((size_t)B * 256 * 2);
  CK(cudaMemset(stats, 0, (size_t)B * 256 * 2 * sizeof(double)));
  SD(sdof_instnorm_stats_nhwc_h(x16, B, (int64_t)h * w, 256, stats, nullptr));
  SD(sdof_instnorm_apply_nhwc_h(x16, stats, nullptr, x16, B, (int64_t)h * w, 256, 1e-5f, 1, nullptr));
  CK(cudaDeviceSynchronize());
  printf("ok fp16 update-loop glue (PDL launches) B=%d %dx%d\n", B, h, w);
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  const bool all = !strcmp(what, "all");
  if (all || !strcmp(what, "corr")) {
    corr_case(1, 24, 40, 256, SDOF_PREC_FP16, "fp16");    // resident kernel: partial source block (960 = 3.75 x 256), partial patches
    corr_case(2, 18, 22, 64, SDOF_PREC_BF16, "bf16");     // odd pooled sizes 9x11 / 4x5 / 2x2
    corr_case(1, 24, 40, 256, SDOF_PREC_TF32, "tf32");    // streaming kernel
    corr_case(1, 18, 22, 64, SDOF_PREC_3XTF32, "3xtf32");
    corr_case(1, 10, 12, 32, SDOF_PREC_FP32, "fp32");
    corr_half_case(2, 24, 40, 256);                       // 32x4 / 16x8 / 8x16 patches, partial source block
    corr_half_case(1, 18, 22, 64);                        // odd pooled widths 11, 5: the pair store's padding column
    alt_corr_case(1, 20, 24, 256);
  }
  if (all || !strcmp(what, "warp")) {
    warp_case(2, 100, 140, 0.0f, 2.5f, "smooth");     // staged tiles, partial tiles at the right / bottom edge
    warp_case(1, 96, 128, 14.0f, 0.0f, "noisy");      // fallback tiles
  }
  if (all || !strcmp(what, "mask")) mask_case(1, 90, 124);
  if (all || !strcmp(what, "glue")) glue_case(2, 12, 20);
  if (all || !strcmp(what, "glue")) {
    glue16_case(2, 12, 20);
    glue16_case(1, 23, 37);     // sizes no vector width divides
  }
  if (all || !strcmp(what, "gru")) {
    gru_tc_case(1, 24, 64);     // full 32x4 patches
    gru_tc_case(2, 11, 20);     // partial tiles on both axes, B > 1
  }
  printf("driver done\n");
  return 0;
}
