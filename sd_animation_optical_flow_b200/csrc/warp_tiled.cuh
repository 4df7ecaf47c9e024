// Tile machinery of the 3-channel u8 bicubic warp (shared by warp.cu and fused.cu).
//
// The per-pixel gather of cv2.remap (16 taps x 3 interleaved channels at an arbitrary byte
// alignment + a 32-byte weight row) is LSU/issue-bound when every tap comes from global memory
// (round-1 ncu: L1TEX 66-82 %, 330 instructions per pixel, 11 % of the HBM roofline).  Here a
// 256-thread GROUP owns a 32x32 output tile:
//
//   A. every lane loads the flow of its 4 pixels (lane = x, 4 rows per warp; the NEXT tile's flow is
//      prefetched into registers), quantises the sampling coordinates exactly like OpenCV (1/32 px)
//      and the group reduces the bounding box of all taps (redux.sync + one named barrier);
//   B. the group stages that source rectangle ONCE, with coalesced word loads, into shared
//      memory, converted to the operand layout of dp2a: entry x of a row is 8 bytes
//         { p[x].c0, p[x+1].c0, p[x].c1, p[x+1].c1 | p[x].c2, p[x+1].c2, -, - }
//      so a tap row is two conflict-free LDS.64 (entries sx and sx+2) and six dp2a, with no
//      per-pixel byte shuffling; pixels outside the source image are staged as 0
//      (BORDER_CONSTANT), which removes the border special case;
//   C. each lane accumulates its pixels from shared memory (weight rows come from a shared-memory
//      copy of the 32 KB table, split in two 16-byte halves so consecutive fractions are
//      conflict-free), and a warp packs its 32 pixels x 3 bytes into 24 words with two shuffles
//      for one coalesced 96-byte store.
//
// Three groups share one CTA (one CTA per SM, persistent over tiles) so the weight table is
// loaded into shared memory once per SM and the groups' load phases overlap each other's
// arithmetic.  Measured group shapes (32 frames 768x512, zero / smooth / sheared flow, us): 12 x 64 threads
// 122/156/208, 6 x 128: 108/145/175, 4 x 256 (64 registers): 116/152/177, 3 x 256: 104/141/164, 2 x 384: 123/159/181.  A tile whose bounding box does not fit the group's region (non-smooth flow,
// NaN / far out-of-image samples) falls back to the per-pixel global-memory path (warp.cuh).
#pragma once

#include "warp.cuh"

namespace sdof {

constexpr int kWtGroups = 3;            // independent groups per CTA (named barriers 1..3); 768 threads leave 85 registers per thread
constexpr int kWtGroupWarps = 8;        // warp w owns tile rows 4w..4w+3, lane = x
constexpr int kWtGroupThreads = 32 * kWtGroupWarps;
constexpr int kWtThreads = kWtGroups * kWtGroupThreads;
constexpr int kWtTile = 32;             // tile width
constexpr int kWtTileH = 4 * kWtGroupWarps;
constexpr int kWtRegionCap = 7936;      // 8-byte entries per group (62 KB): e.g. 88 x 88 source pixels (plain warp kernel)
constexpr int kWtMaxDim = 32766;        // largest source width / height of the tiled kernel (see wt_fixed_coord)

template <int CAP>
struct WtSmemT {
  uint4 tabA[1024];                     // weight rows ky = 0,1 of every (fy,fx)
  uint4 tabB[1024];                     // weight rows ky = 2,3
  int red[kWtGroups][kWtGroupWarps][4];  // per-warp bounding boxes
  uint2 region[kWtGroups][CAP];
};
using WtSmem = WtSmemT<kWtRegionCap>;

// n / d for 0 <= n < 2^31 as one multiply-high and a shift (host-built; mul == 0 means d == 1).
struct WtDiv {
  unsigned mul, shr;
};
inline WtDiv wt_make_div(unsigned d) {
  WtDiv r = {0u, 0u};
  if (d <= 1) return r;
  unsigned l = 0;
  while ((1ull << l) < d) ++l;                      // 2^(l-1) < d <= 2^l
  const unsigned long long k = 31ull + l;
  r.mul = (unsigned)(((1ull << k) + d - 1) / d);    // ceil(2^k / d) < 2^32
  r.shr = l - 1;
  return r;
}
__device__ __forceinline__ unsigned wt_div(unsigned n, WtDiv d) { return d.mul ? (__umulhi(n, d.mul) >> d.shr) : n; }

struct WtTiling {
  int tilesX, tiles_per_img, ntiles;
  WtDiv div_tpi, div_tx;
};
inline WtTiling wt_make_tiling(int B, int H, int W) {
  WtTiling T;
  T.tilesX = (W + kWtTile - 1) / kWtTile;
  T.tiles_per_img = T.tilesX * ((H + kWtTileH - 1) / kWtTileH);
  T.ntiles = T.tiles_per_img * B;
  T.div_tpi = wt_make_div((unsigned)T.tiles_per_img);
  T.div_tx = wt_make_div((unsigned)T.tilesX);
  return T;
}
__device__ __forceinline__ void wt_tile_coords(const WtTiling& T, int t, int& b, int& tyi, int& txi) {
  b = (int)wt_div((unsigned)t, T.div_tpi);
  const int rem = t - b * T.tiles_per_img;
  tyi = (int)wt_div((unsigned)rem, T.div_tx);
  txi = rem - tyi * T.tilesX;
}

__device__ __forceinline__ void wt_group_barrier(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kWtGroupThreads) : "memory");
}

template <class SMEM>
__device__ __forceinline__ void wt_load_table(SMEM& S, const int16_t* __restrict__ tab) {
  const uint4* t4 = reinterpret_cast<const uint4*>(tab);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    const uint4 v = __ldg(t4 + i);
    if (i & 1)
      S.tabB[i >> 1] = v;
    else
      S.tabA[i >> 1] = v;
  }
}

// fixed_coord() (warp.cuh) without the two range compares per coordinate: cvt.rni.s32.f32 saturates, and
// fmaxf sends NaN to -2^31, so only +huge differs from x86's cvtss2si (INT_MAX instead of INT_MIN).  Both
// land outside a source of at most kWtMaxDim pixels (sx = 32766 vs -32769), where every tap is the constant
// border 0 and the weights do not matter -- the launcher sends larger sources to the generic kernel.
__device__ __forceinline__ FixedCoord wt_fixed_coord(float mx, float my) {
  const int qx = __float2int_rn(fmaxf(mx * 32.0f, -2147483648.0f));
  const int qy = __float2int_rn(fmaxf(my * 32.0f, -2147483648.0f));
  FixedCoord fc;
  fc.sx = sat_s16(qx >> 5) - 1;
  fc.sy = sat_s16(qy >> 5) - 1;
  fc.fidx = ((qy & 31) << 5) | (qx & 31);
  return fc;
}

// Per-thread state of one tile: the quantised coordinates of the thread's 4 pixels.  Lanes / rows outside the
// output image hold a copy of the nearest inside pixel (their stores are masked), so no validity flags exist.
struct WtPixels {
  int sx[4], sy[4], fid[4];
};

struct WtRegion {
  int rx0, ry0;   // source coordinates of entry (0,0)
  int ngr;        // 4-entry groups staged per row
  int pitch;      // entries per staged row = 4*ngr + 2: rows start 16 bytes apart modulo 128, so lanes that sit
                  // in different rows at similar x (sheared flow) do not collide on a bank
  int rows;
  bool staged;
};

// Phase A (second half): bounding box of the group's taps -> region geometry.  Contains the
// group barrier that also orders the previous tile's reads of `region` before this tile's writes.
template <class SMEM>
__device__ __forceinline__ WtRegion wt_bbox(SMEM& S, int cap, int grp, int gw, int lane, const WtPixels& px) {
  int mnx = __vimin3_s32(px.sx[0], px.sx[1], min(px.sx[2], px.sx[3]));
  int mxx = __vimax3_s32(px.sx[0], px.sx[1], max(px.sx[2], px.sx[3]));
  int mny = __vimin3_s32(px.sy[0], px.sy[1], min(px.sy[2], px.sy[3]));
  int mxy = __vimax3_s32(px.sy[0], px.sy[1], max(px.sy[2], px.sy[3]));
  mnx = __reduce_min_sync(0xffffffffu, mnx);
  mny = __reduce_min_sync(0xffffffffu, mny);
  mxx = __reduce_max_sync(0xffffffffu, mxx);
  mxy = __reduce_max_sync(0xffffffffu, mxy);
  if (lane == 0) *reinterpret_cast<int4*>(S.red[grp][gw]) = make_int4(mnx, mny, mxx, mxy);
  wt_group_barrier(grp);
#pragma unroll
  for (int w = 0; w < kWtGroupWarps; ++w) {
    const int4 v = *reinterpret_cast<const int4*>(S.red[grp][w]);
    mnx = min(mnx, v.x);
    mny = min(mny, v.y);
    mxx = max(mxx, v.z);
    mxy = max(mxy, v.w);
  }
  WtRegion R;
  R.rx0 = mnx;
  R.ry0 = mny;
  // entries sx .. sx+2 of rows sy .. sy+3 are read; |sx|,|sy| <= 32769 so the differences fit easily
  const int ew = mxx - mnx + 3, eh = mxy - mny + 4;
  R.ngr = (ew + 3) >> 2;
  R.pitch = 4 * R.ngr + 2;
  R.rows = eh;
  R.staged = ew <= 2048 && eh <= 2048 && R.pitch * eh <= cap;
  return R;
}

// 15 bytes of pixels xs..xs+4 of one source row (a = address of pixel xs, possibly outside the row), 0 outside
// [0, Ws).  Out of line: it runs only at image borders and must not shape the main path's registers.
static __device__ __noinline__ uint4 wt_stage_bytes(const unsigned char* __restrict__ a, int xs, int Ws) {
  unsigned r[4] = {0, 0, 0, 0};
#pragma unroll
  for (int n = 0; n < 15; ++n) {
    const int xx = xs + n / 3;
    if ((unsigned)xx < (unsigned)Ws) r[n >> 2] |= (unsigned)a[n] << (8 * (n & 3));
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}

// Phase B: stage the source rectangle as dp2a pair entries.  Ends with the group barrier.
__device__ __forceinline__ void wt_stage(uint2* __restrict__ region, const WtRegion& R,
                                         const unsigned char* __restrict__ img, int Hs, int Ws, int gt, int grp) {
  if (R.staged) {
    const int items = R.ngr * R.rows;
    const float inv = 1.0f / (float)R.ngr;
    for (int i = gt; i < items; i += kWtGroupThreads) {
      // i / ngr: (i + 0.5) / ngr is >= 0.5 / ngr >= 1e-3 away from an integer, the float error is < 1e-4
      const int ry = (int)(((float)i + 0.5f) * inv);
      const int j = i - ry * R.ngr;
      const int y = R.ry0 + ry, xs = R.rx0 + 4 * j;
      unsigned r0 = 0, r1 = 0, r2 = 0, r3 = 0;  // bytes b0..b15 of pixels xs..xs+4 (15 used)
      if ((unsigned)y < (unsigned)Hs && xs > -5 && xs < Ws) {
        // pixel index fits int32 (32766^2 < 2^31) and may be slightly negative (xs >= -4 in row 0): signed 64-bit offset
        const unsigned char* a = img + (int64_t)(y * Ws + xs) * 3;
        if (xs >= 1 && xs + 7 <= Ws) {
          // aligned word loads may touch up to 3 bytes before and 5 bytes after the 15 wanted ones: with
          // xs >= 1 and xs + 7 <= Ws they stay inside this image row, so no buffer-bound checks are needed
          const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(a) & 3);
          const unsigned* q = reinterpret_cast<const unsigned*>(a - mis);
          const unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3);
          const unsigned w4 = mis >= 2 ? __ldg(q + 4) : 0u;
          const unsigned sh = mis * 8;
          r0 = __funnelshift_r(w0, w1, sh);
          r1 = __funnelshift_r(w1, w2, sh);
          r2 = __funnelshift_r(w2, w3, sh);
          r3 = __funnelshift_r(w3, w4, sh);
        } else {
          const uint4 r = wt_stage_bytes(a, xs, Ws);  // row ends: byte loads, pixels outside the row are 0
          r0 = r.x; r1 = r.y; r2 = r.z; r3 = r.w;
        }
      }
      uint4 o0, o1;
      o0.x = __byte_perm(r0, r1, 0x4130);  // b0 b3 b1 b4
      o0.y = __byte_perm(r0, r1, 0x0052);  // b2 b5
      o0.z = __byte_perm(r0, r1, 0x7463);  // b3 b6 b4 b7
      o0.w = __byte_perm(r1, r2, 0x0041);  // b5 b8
      o1.x = __byte_perm(r1, r2, 0x6352);  // b6 b9 b7 b10
      o1.y = __byte_perm(r1, r2, 0x0074);  // b8 b11
      o1.z = __byte_perm(r2, r3, 0x5241);  // b9 b12 b10 b13
      o1.w = __byte_perm(r2, r3, 0x0063);  // b11 b14
      uint4* d = reinterpret_cast<uint4*>(region + ry * R.pitch + 4 * j);
      d[0] = o0;
      d[1] = o1;
    }
  }
  wt_group_barrier(grp);
}

// Phase C: one pixel from the staged region; returns the 3 channels in the low 24 bits.
template <class SMEM>
__device__ __forceinline__ unsigned wt_pixel(const SMEM& S, const uint2* __restrict__ region, const WtRegion& R,
                                             int sx, int sy, int fid) {
  const uint2* e = region + (sy - R.ry0) * R.pitch + (sx - R.rx0);
  const uint4 wa = S.tabA[fid];
  const uint4 wb = S.tabB[fid];
  const unsigned wlo[4] = {wa.x, wa.z, wb.x, wb.z};
  const unsigned whi[4] = {wa.y, wa.w, wb.y, wb.w};
  int a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const uint2 e0 = e[ky * R.pitch];      // taps kx = 0,1
    const uint2 e1 = e[ky * R.pitch + 2];  // taps kx = 2,3
    a0 = dp2a_lo_s16u8(wlo[ky], e0.x, a0);
    a1 = dp2a_hi_s16u8(wlo[ky], e0.x, a1);
    a2 = dp2a_lo_s16u8(wlo[ky], e0.y, a2);
    a0 = dp2a_lo_s16u8(whi[ky], e1.x, a0);
    a1 = dp2a_hi_s16u8(whi[ky], e1.x, a1);
    a2 = dp2a_lo_s16u8(whi[ky], e1.y, a2);
  }
  return (unsigned)cast_q15_u8(a0) | ((unsigned)cast_q15_u8(a1) << 8) | ((unsigned)cast_q15_u8(a2) << 16);
}

// per-pixel global-memory path for tiles whose source rectangle does not fit shared memory; kept out of
// line so the staged path's register allocation is not shaped by it
static __device__ __noinline__ unsigned cubic_u8_c3_outlined(const int16_t* __restrict__ tab, const unsigned char* __restrict__ img,
                                                      const unsigned char* __restrict__ buf_end, int Hs, int Ws, int sx,
                                                      int sy, int fidx) {
  FixedCoord fc;
  fc.sx = sx;
  fc.sy = sy;
  fc.fidx = fidx;
  return cubic_u8_c3(tab, img, buf_end, Hs, Ws, fc);
}

// Lane constants of the row packer: word j of a 96-byte row segment starts in pixel j + j/3 at channel j % 3.
struct WtPack {
  int p;        // first source lane
  unsigned sh;  // 8 * channel offset
};
__device__ __forceinline__ WtPack wt_make_pack(int lane) {
  const int j = lane < 24 ? lane : 23;
  WtPack k;
  k.p = j + j / 3;
  k.sh = 8u * (unsigned)(j - (j / 3) * 3);
  return k;
}

// Store one 32-pixel row segment (`v` = this lane's pixel, 3 bytes).  All 32 lanes must call.
// `full` = all 32 pixels inside the image and the segment word-aligned: 24 lanes store one word each.
__device__ __forceinline__ void wt_store_row(unsigned char* __restrict__ o, unsigned v, int lane, const WtPack& pk, bool full,
                                             bool lane_valid) {
  const unsigned a = __shfl_sync(0xffffffffu, v, pk.p);
  const unsigned b = __shfl_sync(0xffffffffu, v, pk.p + 1);
  if (full) {
    if (lane < 24) __stcs(reinterpret_cast<unsigned*>(o) + lane, (a >> pk.sh) | (b << (24u - pk.sh)));
  } else if (lane_valid) {
    o[3 * lane] = (unsigned char)(v & 0xff);
    o[3 * lane + 1] = (unsigned char)((v >> 8) & 0xff);
    o[3 * lane + 2] = (unsigned char)((v >> 16) & 0xff);
  }
}

}  // namespace sdof
