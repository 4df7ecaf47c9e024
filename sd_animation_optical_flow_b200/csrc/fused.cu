// Fused per-non-key-frame pass (SURVEY §8d "warp + confidence + mask + composite",
// 26 B/pixel algorithmic: flow 8 + logits 8 + src 3 + base 3 + out 3 + mask 1):
//   conf = softmax(weight_map)[0]                         (pdcnet_of.py:72-74)
//   mask = dilate_ellipse(255*(conf < thres), ksize)      (ofgen_pixel_inpaint.py:262-267)
//   out  = mask > 127 ? base : cubic_warp(src, x + flow)  (pdcnet_of.py:34-42, ofgen_pixel_inpaint.py:251-260 with ppw=1)
// One CTA per 64x16 tile: the low-confidence indicator of the tile + halo is staged in
// shared memory (logits are re-read only in the halo), dilated with packed byte max,
// and each thread then warps / selects 4 consecutive pixels.
#include "dilate.cuh"
#include "warp.cuh"

namespace sdof {

__device__ __forceinline__ float softmax0(const float* __restrict__ w, int K, int64_t HW) {
  float m = w[0];
  for (int k = 1; k < K; ++k) m = fmaxf(m, w[k * HW]);
  float s = 0.f;
  for (int k = 0; k < K; ++k) s = __fadd_rn(s, expf(__fsub_rn(w[k * HW], m)));
  return __fdiv_rn(expf(__fsub_rn(w[0], m)), s);
}

__global__ void __launch_bounds__(kDilThreads) warp_mask_composite_kernel(
    const int16_t* __restrict__ tab, const unsigned char* __restrict__ src, const unsigned char* __restrict__ src_end,
    const unsigned char* __restrict__ base, const float* __restrict__ flow, const float* __restrict__ wm, int K, int H,
    int W, int64_t src_bstride, float thres, EllipseRows e, unsigned char* __restrict__ out,
    unsigned char* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char tile[];
  const int r = e.ksize >> 1;
  const int tx0 = blockIdx.x * kDilTW, ty0 = blockIdx.y * kDilTH;
  const int b = blockIdx.z;
  const int64_t hw = (int64_t)H * W;
  const float* wmb = wm + (int64_t)b * K * hw;
  dil_stage(tile, r, ty0, tx0, H, W, [&](int gy, int gx) -> unsigned char {
    return softmax0(wmb + (int64_t)gy * W + gx, K, hw) < thres ? 255 : 0;
  });
  __syncthreads();
  const int lx = (threadIdx.x % (kDilTW / 4)) * 4, ly = threadIdx.x / (kDilTW / 4);
  const int gx = tx0 + lx, gy = ty0 + ly;
  if (gy >= H || gx >= W) return;
  const unsigned m4 = dil_apply4(tile, e, lx, ly);
  const int64_t p0 = (int64_t)b * hw + (int64_t)gy * W + gx;
  const unsigned char* img = src + b * src_bstride;
  const bool full = gx + 4 <= W;
  float fl[8];
  if (full && ((reinterpret_cast<uintptr_t>(flow + p0 * 2) & 15) == 0)) {
    const float4 f0 = __ldcs(reinterpret_cast<const float4*>(flow + p0 * 2));
    const float4 f1 = __ldcs(reinterpret_cast<const float4*>(flow + p0 * 2) + 1);
    fl[0] = f0.x; fl[1] = f0.y; fl[2] = f0.z; fl[3] = f0.w;
    fl[4] = f1.x; fl[5] = f1.y; fl[6] = f1.z; fl[7] = f1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = gx + i < W;
      fl[2 * i] = ok ? flow[(p0 + i) * 2] : 0.f;
      fl[2 * i + 1] = ok ? flow[(p0 + i) * 2 + 1] : 0.f;
    }
  }
  unsigned px[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    px[i] = 0;
    if (gx + i < W) {
      const unsigned mb = (m4 >> (8 * i)) & 0xff;
      if (mb > 127) {
        const unsigned char* q = base + (p0 + i) * 3;
        px[i] = (unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16);
      } else {
        const FixedCoord fc = fixed_coord(map_coord(gx + i, fl[2 * i], 1.f), map_coord(gy, fl[2 * i + 1], 1.f));
        px[i] = cubic_u8_c3(tab, img, src_end, H, W, fc);
      }
    }
  }
  unsigned char* o = out + p0 * 3;
  unsigned char* mo = mask + p0;
  if (full && ((reinterpret_cast<uintptr_t>(o) & 3) == 0) && ((reinterpret_cast<uintptr_t>(mo) & 3) == 0)) {
    unsigned* ow = reinterpret_cast<unsigned*>(o);
    __stcs(ow, px[0] | (px[1] << 24));
    __stcs(ow + 1, (px[1] >> 8) | (px[2] << 16));
    __stcs(ow + 2, (px[2] >> 16) | (px[3] << 8));
    __stcs(reinterpret_cast<unsigned*>(mo), m4);
  } else {
    for (int i = 0; i < 4 && gx + i < W; ++i) {
      o[3 * i] = (unsigned char)(px[i] & 0xff);
      o[3 * i + 1] = (unsigned char)((px[i] >> 8) & 0xff);
      o[3 * i + 2] = (unsigned char)((px[i] >> 16) & 0xff);
      mo[i] = (unsigned char)((m4 >> (8 * i)) & 0xff);
    }
  }
}

}  // namespace sdof

extern "C" int sdof_warp_mask_composite(const uint8_t* src, const uint8_t* base, const float* flow,
                                        const float* weight_map, int B, int src_batched, int H, int W, float thres,
                                        int ksize, uint8_t* out, uint8_t* mask, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(src && base && flow && weight_map && out && mask, "sdof_warp_mask_composite: NULL pointer");
  SDOF_REQUIRE(B >= 0 && B <= 65535 && H >= 1 && W >= 1 && H <= 32767 && W <= 32767,
               "sdof_warp_mask_composite: bad sizes B=%d H=%d W=%d", B, H, W);
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(flow) & 7) == 0, "sdof_warp_mask_composite: flow must be 8-byte aligned");
  EllipseRows e;
  if (make_ellipse(ksize, &e))
    return fail(SDOF_ERR_INVALID, "sdof_warp_mask_composite: ksize must be odd in [1,%d], got %d", kDilMaxK, ksize);
  if (B == 0) return SDOF_OK;
  CubicTables tabs;
  int rc = get_cubic_tables(&tabs);
  if (rc) return rc;
  const int64_t img_bytes = (int64_t)H * W * 3;
  const bool aligned = (reinterpret_cast<uintptr_t>(src) & 3) == 0;
  const uint8_t* src_end = aligned ? src + (src_batched ? (int64_t)B : 1) * img_bytes : src;
  dim3 grid(ceil_div(W, kDilTW), ceil_div(H, kDilTH), B);
  warp_mask_composite_kernel<<<grid, kDilThreads, dil_smem_bytes(ksize >> 1), as_stream(stream)>>>(
      tabs.i16, src, src_end, base, flow, weight_map, 2, H, W, src_batched ? img_bytes : 0, thres, e, out, mask);
  SDOF_LAUNCH_CHECK("warp_mask_composite_kernel");
  return SDOF_OK;
}
