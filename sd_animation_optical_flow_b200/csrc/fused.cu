// Fused per-non-key-frame pass (SURVEY §8d "warp + confidence + mask + composite",
// 26 B/pixel algorithmic: flow 8 + logits 8 + src 3 + base 3 + out 3 + mask 1):
//   conf = softmax(weight_map)[0]                         (pdcnet_of.py:72-74)
//   mask = dilate_ellipse(255*(conf < thres), ksize)      (ofgen_pixel_inpaint.py:262-267)
//   out  = mask > 127 ? base : cubic_warp(src, x + flow)  (pdcnet_of.py:34-42, ofgen_pixel_inpaint.py:251-260 with ppw=1)
// One CTA per 64x16 tile: the low-confidence indicator of the tile + halo is staged in
// shared memory (logits are re-read only in the halo), dilated with packed byte max,
// and each thread then warps / selects 4 consecutive pixels.
#include "dilate.cuh"
#include "warp_tiled.cuh"

namespace sdof {

__device__ __forceinline__ float softmax0(const float* __restrict__ w, int K, int64_t HW) {
  float m = w[0];
  for (int k = 1; k < K; ++k) m = fmaxf(m, w[k * HW]);
  float s = 0.f;
  for (int k = 0; k < K; ++k) s = __fadd_rn(s, expf(__fsub_rn(w[k * HW], m)));
  return __fdiv_rn(expf(__fsub_rn(w[0], m)), s);
}

// ---- tiled version (warp_tiled.cuh): a 128-thread group owns a 32x16 tile ------------------------------------
// Per tile: (0) the low-confidence indicator of tile + halo goes to the group's shared indicator buffer (softmax from
// the logits), is dilated with packed byte max by 4-pixel threads and the mask is written out (global + a 32x16
// shared copy); (A-C) the warp phases of warp_tiled.cuh with lane = x; a masked pixel takes `base` instead.
constexpr int kFuRegionCap = 3200;                       // staged source entries per group (25 KB)
constexpr int kFuIndPitch = kWtTile + 2 * 16;            // indicator row pitch for any r <= 15 (r4 <= 16)
constexpr int kFuIndRows = kWtTileH + 2 * 15;
struct FuSmem {
  WtSmemT<kFuRegionCap> wt;
  unsigned char ind[kWtGroups][kFuIndRows * kFuIndPitch];  // indicator of tile + halo (pixel (0,0) at column r4, row r)
  unsigned char msk[kWtGroups][kWtTileH * kWtTile];        // dilated mask of the tile
};

__global__ void __launch_bounds__(kWtThreads, 1) warp_mask_composite_tiled_kernel(
    const int16_t* __restrict__ tab, const unsigned char* __restrict__ src, const unsigned char* __restrict__ old_src_end,
    const unsigned char* __restrict__ base, const float* __restrict__ flow, const float* __restrict__ wm, int K, WtTiling T,
    int H, int W, int64_t src_bstride, float thres, EllipseRows e, unsigned char* __restrict__ out,
    unsigned char* __restrict__ mask, int vec_ok) {
  extern __shared__ __align__(16) unsigned char fu_smem_raw[];
  FuSmem& S = *reinterpret_cast<FuSmem*>(fu_smem_raw);
  wt_load_table(S.wt, tab);
  __syncthreads();
  const int grp = threadIdx.x / kWtGroupThreads, gt = threadIdx.x % kWtGroupThreads, gw = gt >> 5, lane = gt & 31;
  uint2* region = S.wt.region[grp];
  unsigned char* ind = S.ind[grp];
  unsigned char* msk = S.msk[grp];
  const WtPack pk = wt_make_pack(lane);
  const int r = e.ksize >> 1, r4 = (r + 3) & ~3;
  const int ind_cols = kWtTile + 2 * r, ind_rows = kWtTileH + 2 * r;
  const float inv_cols = 1.0f / (float)ind_cols;
  const int64_t hw = (int64_t)H * W;
  const int tstride = gridDim.x * kWtGroups;
  const float2* flow2 = reinterpret_cast<const float2*>(flow);
  for (int t = blockIdx.x * kWtGroups + grp; t < T.ntiles; t += tstride) {
    int b, tyi, txi;
    wt_tile_coords(T, t, b, tyi, txi);
    const int tx0 = txi * kWtTile, ty0 = tyi * kWtTileH;
    // flow of this thread's 4 pixels (issued first: the loads fly during the mask phase)
    const int gx = min(tx0 + lane, W - 1), gy0 = ty0 + gw * 4;
    float2 f[4];
    {
      const float2* fp = flow2 + (int64_t)b * hw + gx;
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = __ldcs(fp + (unsigned)(min(gy0 + k, H - 1) * W));
    }
    // ---- 0: indicator of tile + halo, dilation, mask
    const float* wmb = wm + (int64_t)b * K * hw;
    for (int i = gt; i < ind_rows * ind_cols; i += kWtGroupThreads) {
      const int ly = (int)(((float)i + 0.5f) * inv_cols), lx = i - ly * ind_cols;
      const int y = ty0 - r + ly, x = tx0 - r + lx;
      unsigned char v = 0;
      if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) v = softmax0(wmb + (int64_t)y * W + x, K, hw) < thres ? 255 : 0;
      ind[ly * kFuIndPitch + (r4 - r) + lx] = v;
    }
    wt_group_barrier(grp);
    {
      const int lx4 = (gt & 7) * 4, ly = gt >> 3;  // 8 threads per tile row, 4 pixels each
      unsigned acc = 0;
      for (int i = 0; i < e.ksize; ++i) {
        const int half = e.half[i];
        const unsigned char* row = ind + (ly + i) * kFuIndPitch + r4 + lx4;  // column of dx = 0
        for (int dx = -half; dx <= half; ++dx) {
          const int a = dx & 3;
          const unsigned* q = reinterpret_cast<const unsigned*>(row + (dx - a));
          const unsigned lo = q[0];
          const unsigned hi = a ? q[1] : 0u;
          acc = __vmaxu4(acc, __funnelshift_r(lo, hi, a * 8));
        }
      }
      *reinterpret_cast<unsigned*>(msk + ly * kWtTile + lx4) = acc;
      const int y = ty0 + ly, x = tx0 + lx4;
      if (y < H && x < W) {
        unsigned char* mo = mask + (int64_t)b * hw + (int64_t)y * W + x;
        if (vec_ok && x + 4 <= W) {
          __stcs(reinterpret_cast<unsigned*>(mo), acc);
        } else {
          for (int i = 0; i < 4 && x + i < W; ++i) mo[i] = (unsigned char)((acc >> (8 * i)) & 0xff);
        }
      }
    }
    // ---- A: quantised coordinates, bounding box (its barrier also publishes msk)
    WtPixels px;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const FixedCoord fc = wt_fixed_coord(map_coord(gx, f[k].x, 1.f), map_coord(min(gy0 + k, H - 1), f[k].y, 1.f));
      px.sx[k] = fc.sx;
      px.sy[k] = fc.sy;
      px.fid[k] = fc.fidx;
    }
    const WtRegion R = wt_bbox(S.wt, kFuRegionCap, grp, gw, lane, px);
    // ---- B: stage the source rectangle
    const unsigned char* img = src + b * src_bstride;
    wt_stage(region, R, img, H, W, gt, grp);
    // ---- C: warp or take base, packed row stores
    const bool seg_full = vec_ok && (tx0 + kWtTile <= W);
    const bool lane_valid = tx0 + lane < W;
    const int64_t row0 = (int64_t)b * hw + (int64_t)gy0 * W + tx0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (gy0 + k >= H) break;  // warp-uniform
      unsigned v;
      if (msk[(gw * 4 + k) * kWtTile + lane] > 127) {
        const unsigned char* q = base + (row0 + (int64_t)k * W + (lane_valid ? lane : 0)) * 3;
        v = (unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16);
      } else if (R.staged) {
        v = wt_pixel(S.wt, region, R, px.sx[k], px.sy[k], px.fid[k]);
      } else {
        v = cubic_u8_c3_outlined(tab, img, old_src_end, H, W, px.sx[k], px.sy[k], px.fid[k]);
      }
      wt_store_row(out + (row0 + (int64_t)k * W) * 3, v, lane, pk, seg_full, lane_valid);
    }
    // the next tile's indicator writes must not overtake this tile's msk reads
    wt_group_barrier(grp);
  }
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
  const int64_t ntiles = (int64_t)ceil_div(W, kWtTile) * ceil_div(H, kWtTileH) * B;
  if (H <= kWtMaxDim && W <= kWtMaxDim && (int64_t)H * W * 3 < 0x7fffffffLL && ntiles < 0x7fffffffLL) {
    static bool attr_set[64] = {};
    int dev = 0;
    SDOF_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      SDOF_CUDA(cudaFuncSetAttribute(warp_mask_composite_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FuSmem)));
      attr_set[dev] = true;
    }
    const WtTiling T = wt_make_tiling(B, H, W);
    const int64_t want = ceil_div64(ntiles, kWtGroups);
    const int grid = (int)(want < sm_count() ? want : sm_count());
    // packed row stores / 4-byte mask stores need word-aligned 32-pixel segments
    const int vec_ok = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(mask)) & 3) == 0 && (W & 3) == 0;
    warp_mask_composite_tiled_kernel<<<grid, kWtThreads, sizeof(FuSmem), as_stream(stream)>>>(
        tabs.i16, src, src_end, base, flow, weight_map, 2, T, H, W, src_batched ? img_bytes : 0, thres, e, out, mask, vec_ok);
    SDOF_LAUNCH_CHECK("warp_mask_composite_tiled_kernel");
    return SDOF_OK;
  }
  dim3 grid(ceil_div(W, kDilTW), ceil_div(H, kDilTH), B);
  warp_mask_composite_kernel<<<grid, kDilThreads, dil_smem_bytes(ksize >> 1), as_stream(stream)>>>(
      tabs.i16, src, src_end, base, flow, weight_map, 2, H, W, src_batched ? img_bytes : 0, thres, e, out, mask);
  SDOF_LAUNCH_CHECK("warp_mask_composite_kernel");
  return SDOF_OK;
}
