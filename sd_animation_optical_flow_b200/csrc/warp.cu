// Backward warp kernels: bit-exact cv2.remap(INTER_CUBIC,
// BORDER_CONSTANT) for u8, fp32-weight cubic for float images, and a bilinear mode
// with grid_sample(align_corners=True, zeros) semantics.
//
// HBM-bound byte work: 14 B/pixel algorithmic (8 flow + 3 src + 3 dst).  One thread
// produces 4 consecutive pixels (two 16-byte flow loads, three 4-byte stores); source
// taps are fetched as aligned 32-bit words and multiplied with packed int16 weights by
// dp2a, so the LSU/ALU cost per pixel stays close to the HBM time.
#include <mutex>

#include "warp_tiled.cuh"

namespace sdof {

int get_cubic_tables(CubicTables* out) {
  static std::mutex mu;
  static CubicTables tabs[64] = {};
  int dev = 0;
  SDOF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SDOF_ERR_UNSUPPORTED, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(mu);
  if (!tabs[dev].i16) {
    void* p = nullptr;
    const size_t bi = sizeof(int16_t) * 1024 * 16, bf = sizeof(float) * 1024 * 16;
    SDOF_CUDA(cudaMalloc(&p, bi + bf));
    SDOF_CUDA(cudaMemcpy(p, cubic_table_i16_host(), bi, cudaMemcpyHostToDevice));
    SDOF_CUDA(cudaMemcpy(static_cast<char*>(p) + bi, cubic_table_f32_host(), bf, cudaMemcpyHostToDevice));
    tabs[dev].i16 = static_cast<const int16_t*>(p);
    tabs[dev].f32 = reinterpret_cast<const float*>(static_cast<char*>(p) + bi);
  }
  *out = tabs[dev];
  return SDOF_OK;
}

// Diagnostic counters of the tiled kernel: [0] tiles served from the staged shared-memory rectangle, [1] tiles that fell
// back to the per-pixel global-memory path (source rectangle larger than the group's region).  One atomic pair per group
// at kernel exit; read (and reset) by sdof_warp_tile_stats.
__device__ unsigned long long g_wt_stats[2];

// ---------------------------------------------------------------- cubic, u8, C = 3
// Persistent CTAs (one per SM, 3 groups of 256 threads); a group walks 32x32 output tiles
// (warp_tiled.cuh).  `old_src_end` is the bound of the per-pixel fallback path (cubic_u8_c3).
__global__ void __launch_bounds__(kWtThreads, 1) warp_cubic_u8c3_tiled_kernel(
    const int16_t* __restrict__ tab, const unsigned char* __restrict__ src, const float* __restrict__ flow,
    unsigned char* __restrict__ dst, WtTiling T, int Hs, int Ws, int H, int W, int64_t src_bstride, float sign,
    const unsigned char* __restrict__ old_src_end, int dst_vec_ok) {
  extern __shared__ __align__(16) unsigned char wt_smem_raw[];
  WtSmem& S = *reinterpret_cast<WtSmem*>(wt_smem_raw);
  wt_load_table(S, tab);
  __syncthreads();
  const int grp = threadIdx.x / kWtGroupThreads, gt = threadIdx.x % kWtGroupThreads, gw = gt >> 5, lane = gt & 31;
  uint2* region = S.region[grp];
  const WtPack pk = wt_make_pack(lane);
  const int tstride = gridDim.x * kWtGroups;
  const float2* flow2 = reinterpret_cast<const float2*>(flow);
  int t = blockIdx.x * kWtGroups + grp;
  if (t >= T.ntiles) return;  // group-uniform: the group's named barrier is never used by a partial group
  // Pixels of this thread in tile (b, tyi, txi): column min(tx0 + lane, W-1), rows min(gy0 + k, H-1).  The flow of
  // the NEXT tile is loaded before this tile's barriers and arithmetic (register double buffer), so the largest
  // HBM stream (8 of the 14 bytes per pixel) is always in flight.
  int b, tyi, txi;
  wt_tile_coords(T, t, b, tyi, txi);
  int n_tiles = 0, n_fallback = 0;
  float2 f[4];
  {
    const int gx = min(txi * kWtTile + lane, W - 1);
    const float2* fp = flow2 + (int64_t)b * H * W + gx;
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = __ldcs(fp + (unsigned)(min(tyi * kWtTileH + gw * 4 + k, H - 1) * W));
  }
  for (;;) {
    const int tx0 = txi * kWtTile, gy0 = tyi * kWtTileH + gw * 4;
    const int gx = min(tx0 + lane, W - 1);
    // ---- A: flow -> quantised coordinates
    WtPixels px;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const FixedCoord fc = wt_fixed_coord(map_coord(gx, f[k].x, sign), map_coord(min(gy0 + k, H - 1), f[k].y, sign));
      px.sx[k] = fc.sx;
      px.sy[k] = fc.sy;
      px.fid[k] = fc.fidx;
    }
    const unsigned char* img = src + b * src_bstride;
    unsigned char* orow = dst + (((int64_t)b * H + gy0) * W + tx0) * 3;
    const int t_next = t + tstride;
    int bn = b, tyn = tyi, txn = txi;
    if (t_next < T.ntiles) {
      wt_tile_coords(T, t_next, bn, tyn, txn);
      const int gxn = min(txn * kWtTile + lane, W - 1);
      const float2* fp = flow2 + (int64_t)bn * H * W + gxn;
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = __ldcs(fp + (unsigned)(min(tyn * kWtTileH + gw * 4 + k, H - 1) * W));
    }
    const WtRegion R = wt_bbox(S, kWtRegionCap, grp, gw, lane, px);
    // ---- B: stage the source rectangle
    wt_stage(region, R, img, Hs, Ws, gt, grp);
    // ---- C: taps from shared memory, packed row stores
    const bool seg_full = dst_vec_ok && (tx0 + kWtTile <= W);
    const bool lane_valid = tx0 + lane < W;
    ++n_tiles;
    n_fallback += R.staged ? 0 : 1;
    // all four pixels first (independent LDS -> dp2a chains the scheduler can interleave), then the row stores
    unsigned v[4];
    if (R.staged) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = wt_pixel(S, region, R, px.sx[k], px.sy[k], px.fid[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = cubic_u8_c3_outlined(tab, img, old_src_end, Hs, Ws, px.sx[k], px.sy[k], px.fid[k]);
    }
    if (seg_full && gy0 + 4 <= H) {
      // common case, straight-line: 4 full rows, each 24 words (W % 4 == 0, so a row is W*3/4 words)
      unsigned* op = reinterpret_cast<unsigned*>(orow) + lane;
      const unsigned row_words = (unsigned)(W * 3) >> 2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned a = __shfl_sync(0xffffffffu, v[k], pk.p);
        const unsigned b2 = __shfl_sync(0xffffffffu, v[k], pk.p + 1);
        if (lane < 24) __stcs(op + k * row_words, (a >> pk.sh) | (b2 << (24u - pk.sh)));
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (gy0 + k >= H) break;  // warp-uniform
        wt_store_row(orow + (unsigned)(k * W * 3), v[k], lane, pk, seg_full, lane_valid);
      }
    }
    if (t_next >= T.ntiles) break;
    t = t_next;
    b = bn;
    tyi = tyn;
    txi = txn;
  }
  if (gt == 0) {
    atomicAdd(&g_wt_stats[0], (unsigned long long)(n_tiles - n_fallback));
    if (n_fallback) atomicAdd(&g_wt_stats[1], (unsigned long long)n_fallback);
  }
}

// ---------------------------------------------------------------- generic paths
template <typename T, bool kCubic>
__global__ void __launch_bounds__(256) warp_generic_kernel(CubicTables tabs, const T* __restrict__ src,
                                                           const float* __restrict__ flow,
                                                           T* __restrict__ dst, int B, int Hs, int Ws, int C, int H,
                                                           int W, int64_t src_bstride, float sign) {
  const int64_t npix = (int64_t)B * H * W;
  const int64_t hw = (int64_t)H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    int b, y, x;
    decompose_pixel(p, hw, W, b, y, x);
    const float2 f = *reinterpret_cast<const float2*>(flow + p * 2);
    const T* img = src + b * src_bstride;
    if (kCubic) {
      const FixedCoord fc = fixed_coord(map_coord(x, f.x, sign), map_coord(y, f.y, sign));
      if (sizeof(T) == 1) {
        int acc[4] = {0, 0, 0, 0};
        const int16_t* wt = tabs.i16 + fc.fidx * 16;
        for (int ky = 0; ky < 4; ++ky) {
          const int yy = fc.sy + ky;
          if ((unsigned)yy >= (unsigned)Hs) continue;
          for (int kx = 0; kx < 4; ++kx) {
            const int xx = fc.sx + kx;
            if ((unsigned)xx >= (unsigned)Ws) continue;
            const int w = wt[ky * 4 + kx];
            const T* q = img + ((int64_t)yy * Ws + xx) * C;
            for (int c = 0; c < C; ++c) acc[c] += w * (int)q[c];
          }
        }
        for (int c = 0; c < C; ++c) dst[p * C + c] = (T)cast_q15_u8(acc[c]);
      } else {
        // float image: same 1/32-pixel quantisation, fp32 weights, taps accumulated in
        // OpenCV's order (ky outer, kx inner), no FMA contraction.
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* wt = tabs.f32 + fc.fidx * 16;
        for (int ky = 0; ky < 4; ++ky) {
          const int yy = fc.sy + ky;
          for (int kx = 0; kx < 4; ++kx) {
            const int xx = fc.sx + kx;
            const bool in = (unsigned)yy < (unsigned)Hs && (unsigned)xx < (unsigned)Ws;
            const float w = wt[ky * 4 + kx];
            for (int c = 0; c < C; ++c) {
              const float v = in ? (float)img[((int64_t)yy * Ws + xx) * C + c] : 0.f;
              acc[c] = __fadd_rn(acc[c], __fmul_rn(v, w));
            }
          }
        }
        for (int c = 0; c < C; ++c) dst[p * C + c] = (T)acc[c];
      }
    } else {
      // bilinear, zeros padding, align_corners=True pixel coordinates
      const float xs = __fadd_rn((float)x, __fmul_rn(sign, f.x));
      const float ys = __fadd_rn((float)y, __fmul_rn(sign, f.y));
      const float x0f = floorf(xs), y0f = floorf(ys);
      // NaN / huge coordinates land outside the image (all taps zero, finite weights)
      const bool okx = x0f >= -2.f && x0f <= (float)Ws + 1.f, oky = y0f >= -2.f && y0f <= (float)Hs + 1.f;
      const float ax = okx ? __fsub_rn(xs, x0f) : 0.f, ay = oky ? __fsub_rn(ys, y0f) : 0.f;
      const int x0 = okx ? (int)x0f : -2;
      const int y0 = oky ? (int)y0f : -2;
      const float w00 = __fmul_rn(__fsub_rn(1.f, ax), __fsub_rn(1.f, ay));
      const float w01 = __fmul_rn(ax, __fsub_rn(1.f, ay));
      const float w10 = __fmul_rn(__fsub_rn(1.f, ax), ay);
      const float w11 = __fmul_rn(ax, ay);
      const bool vx0 = (unsigned)x0 < (unsigned)Ws, vx1 = (unsigned)(x0 + 1) < (unsigned)Ws;
      const bool vy0 = (unsigned)y0 < (unsigned)Hs, vy1 = (unsigned)(y0 + 1) < (unsigned)Hs;
      for (int c = 0; c < C; ++c) {
        const float v00 = (vx0 && vy0) ? (float)img[((int64_t)y0 * Ws + x0) * C + c] : 0.f;
        const float v01 = (vx1 && vy0) ? (float)img[((int64_t)y0 * Ws + x0 + 1) * C + c] : 0.f;
        const float v10 = (vx0 && vy1) ? (float)img[((int64_t)(y0 + 1) * Ws + x0) * C + c] : 0.f;
        const float v11 = (vx1 && vy1) ? (float)img[((int64_t)(y0 + 1) * Ws + x0 + 1) * C + c] : 0.f;
        float r = __fmul_rn(v00, w00);
        r = __fadd_rn(r, __fmul_rn(v01, w01));
        r = __fadd_rn(r, __fmul_rn(v10, w10));
        r = __fadd_rn(r, __fmul_rn(v11, w11));
        if (sizeof(T) == 1) {
          float q = rintf(r);
          q = q < 0.f ? 0.f : (q > 255.f ? 255.f : q);
          dst[p * C + c] = (T)q;
        } else {
          dst[p * C + c] = (T)r;
        }
      }
    }
  }
}

template <typename T>
static int check_warp_args(const char* name, const T* src, const float* flow, int B, int Hs, int Ws, int C, int H, int W,
                           float sign, T* dst) {
  SDOF_REQUIRE(src && flow && dst, "%s: NULL pointer", name);
  SDOF_REQUIRE(B >= 0 && Hs >= 1 && Ws >= 1 && H >= 1 && W >= 1, "%s: bad sizes B=%d Hs=%d Ws=%d H=%d W=%d", name, B, Hs,
               Ws, H, W);
  SDOF_REQUIRE(C >= 1 && C <= 4, "%s: C must be in 1..4, got %d", name, C);
  SDOF_REQUIRE(sign == 1.0f || sign == -1.0f, "%s: sign must be +1 or -1", name);
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(flow) & 7) == 0, "%s: flow must be 8-byte aligned", name);
  SDOF_REQUIRE(Ws <= 32767 && Hs <= 32767, "%s: source larger than 32767 (cv2.remap limit)", name);
  return SDOF_OK;
}

template <typename T, bool kCubic>
static int launch_generic(const char* name, const T* src, const float* flow, int B, int src_batched, int Hs, int Ws,
                          int C, int H, int W, float sign, T* dst, sdof_stream_t stream) {
  int rc = check_warp_args(name, src, flow, B, Hs, Ws, C, H, W, sign, dst);
  if (rc) return rc;
  if (B == 0) return SDOF_OK;
  CubicTables tabs = {nullptr, nullptr};
  if (kCubic && (rc = get_cubic_tables(&tabs))) return rc;
  const int64_t npix = (int64_t)B * H * W;
  const int64_t bstride = src_batched ? (int64_t)Hs * Ws * C : 0;
  warp_generic_kernel<T, kCubic><<<grid_for(npix, 256, 8), 256, 0, as_stream(stream)>>>(tabs, src, flow, dst, B, Hs, Ws, C,
                                                                                         H, W, bstride, sign);
  SDOF_LAUNCH_CHECK(name);
  return SDOF_OK;
}

}  // namespace sdof

extern "C" {

int sdof_warp_cubic_u8(const uint8_t* src, const float* flow, int B, int src_batched, int Hs, int Ws, int C, int H,
                       int W, float sign, uint8_t* dst, sdof_stream_t stream) {
  using namespace sdof;
  if (C != 3)
    return launch_generic<unsigned char, true>("sdof_warp_cubic_u8", src, flow, B, src_batched, Hs, Ws, C, H, W, sign,
                                               dst, stream);
  int rc = check_warp_args("sdof_warp_cubic_u8", src, flow, B, Hs, Ws, C, H, W, sign, dst);
  if (rc) return rc;
  if (B == 0) return SDOF_OK;
  CubicTables tabs;
  if ((rc = get_cubic_tables(&tabs))) return rc;
  // sources larger than kWtMaxDim (or > 2^31 output bytes per image) take the generic kernel (same results)
  if (Hs > kWtMaxDim || Ws > kWtMaxDim || (int64_t)H * W * 3 >= 0x7fffffffLL)
    return launch_generic<unsigned char, true>("sdof_warp_cubic_u8", src, flow, B, src_batched, Hs, Ws, C, H, W, sign, dst, stream);
  const int64_t img_bytes = (int64_t)Hs * Ws * 3;
  const int64_t bstride = src_batched ? img_bytes : 0;
  // per-pixel fallback path: its word loads assume a 4-byte aligned source; otherwise it takes the tap loop
  const bool src_aligned = (reinterpret_cast<uintptr_t>(src) & 3) == 0;
  const uint8_t* old_src_end = src_aligned ? src + (src_batched ? (int64_t)B : 1) * img_bytes : src;
  // packed 96-byte row stores need every 32-pixel segment word-aligned
  const int dst_vec = (reinterpret_cast<uintptr_t>(dst) & 3) == 0 && (W & 3) == 0;
  static bool attr_set[64] = {};
  int dev = 0;
  SDOF_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    SDOF_CUDA(cudaFuncSetAttribute(warp_cubic_u8c3_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(WtSmem)));
    attr_set[dev] = true;
  }
  const int64_t ntiles = (int64_t)ceil_div(W, kWtTile) * ceil_div(H, kWtTileH) * B;
  SDOF_REQUIRE(ntiles < 0x7fffffffLL, "sdof_warp_cubic_u8: too many tiles (B*H*W too large for one launch)");
  const WtTiling T = wt_make_tiling(B, H, W);
  const int64_t want = ceil_div64(ntiles, kWtGroups);
  const int grid = (int)(want < sm_count() ? want : sm_count());
  warp_cubic_u8c3_tiled_kernel<<<grid, kWtThreads, sizeof(WtSmem), as_stream(stream)>>>(
      tabs.i16, src, flow, dst, T, Hs, Ws, H, W, bstride, sign, old_src_end, dst_vec);
  SDOF_LAUNCH_CHECK("warp_cubic_u8c3_tiled_kernel");
  return SDOF_OK;
}

// Diagnostic (host-synchronising): tiles of sdof_warp_cubic_u8's tiled kernel served from the staged shared-memory
// rectangle (out[0]) and by the per-pixel fallback (out[1]) on the current device since the last reset.
int sdof_warp_tile_stats(int64_t* out /* host, 2 */, int reset) {
  using namespace sdof;
  SDOF_REQUIRE(out != nullptr, "sdof_warp_tile_stats: out is NULL");
  unsigned long long h[2] = {0, 0};
  SDOF_CUDA(cudaDeviceSynchronize());
  SDOF_CUDA(cudaMemcpyFromSymbol(h, g_wt_stats, sizeof(h)));
  out[0] = (int64_t)h[0];
  out[1] = (int64_t)h[1];
  if (reset) {
    const unsigned long long z[2] = {0, 0};
    SDOF_CUDA(cudaMemcpyToSymbol(g_wt_stats, z, sizeof(z)));
  }
  return SDOF_OK;
}

// Host evaluation of the multiply-high division the tiled kernels use for tile coordinates (warp_tiled.cuh::wt_make_div /
// wt_div), exported so the CPU test-suite can check the magic numbers against n / d without a GPU.
uint32_t sdof_fastdiv_u31(uint32_t n, uint32_t d) {
  const sdof::WtDiv m = sdof::wt_make_div(d);
  return m.mul ? (uint32_t)(((uint64_t)n * m.mul) >> 32) >> m.shr : n;
}

int sdof_warp_cubic_f32(const float* src, const float* flow, int B, int src_batched, int Hs, int Ws, int C, int H,
                        int W, float sign, float* dst, sdof_stream_t stream) {
  return sdof::launch_generic<float, true>("sdof_warp_cubic_f32", src, flow, B, src_batched, Hs, Ws, C, H, W, sign, dst,
                                           stream);
}

int sdof_warp_bilinear_u8(const uint8_t* src, const float* flow, int B, int src_batched, int Hs, int Ws, int C, int H,
                          int W, float sign, uint8_t* dst, sdof_stream_t stream) {
  return sdof::launch_generic<unsigned char, false>("sdof_warp_bilinear_u8", src, flow, B, src_batched, Hs, Ws, C, H, W,
                                                    sign, dst, stream);
}

int sdof_warp_bilinear_f32(const float* src, const float* flow, int B, int src_batched, int Hs, int Ws, int C, int H,
                           int W, float sign, float* dst, sdof_stream_t stream) {
  return sdof::launch_generic<float, false>("sdof_warp_bilinear_f32", src, flow, B, src_batched, Hs, Ws, C, H, W, sign,
                                            dst, stream);
}

}  // extern "C"
