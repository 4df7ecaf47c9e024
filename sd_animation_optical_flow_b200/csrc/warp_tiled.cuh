// Tile machinery of the 3-channel u8 bicubic warp (shared by warp.cu and fused.cu).
//
// The per-pixel gather of cv2.remap (16 taps x 3 interleaved channels at an arbitrary byte
// alignment + a 32-byte weight row) is LSU/issue-bound when every tap comes from global memory
// (round-1 ncu: L1TEX 66-82 %, 330 instructions per pixel, 11 % of the HBM roofline).  Here a
// 256-thread GROUP owns a 32x32 output tile:
//
//   A. every lane loads the flow of its 4 pixels (lane = x, 4 rows per warp), quantises the
//      sampling coordinates exactly like OpenCV (1/32 px) and the group reduces the bounding
//      box of all taps (redux.sync + one named barrier);
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
// Four groups share one CTA (one CTA per SM, persistent over tiles) so the weight table is
// loaded into shared memory once per SM and the groups' load phases overlap each other's
// arithmetic.  A tile whose bounding box does not fit the group's region (non-smooth flow,
// NaN / far out-of-image samples) falls back to the per-pixel global-memory path (warp.cuh).
#pragma once

#include "warp.cuh"

namespace sdof {

constexpr int kWtGroups = 4;            // independent groups per CTA (named barriers 1..4)
constexpr int kWtGroupThreads = 256;    // 8 warps: warp w owns tile rows 4w..4w+3, lane = x
constexpr int kWtThreads = kWtGroups * kWtGroupThreads;
constexpr int kWtTile = 32;
constexpr int kWtRegionCap = 5632;      // 8-byte entries per group (44 KB): e.g. 75 x 75 source pixels

struct WtSmem {
  uint4 tabA[1024];                     // weight rows ky = 0,1 of every (fy,fx)
  uint4 tabB[1024];                     // weight rows ky = 2,3
  int red[kWtGroups][8][4];             // per-warp bounding boxes
  uint2 region[kWtGroups][kWtRegionCap];
};

__device__ __forceinline__ void wt_group_barrier(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kWtGroupThreads) : "memory");
}

__device__ __forceinline__ void wt_load_table(WtSmem& S, const int16_t* __restrict__ tab) {
  const uint4* t4 = reinterpret_cast<const uint4*>(tab);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    const uint4 v = __ldg(t4 + i);
    if (i & 1)
      S.tabB[i >> 1] = v;
    else
      S.tabA[i >> 1] = v;
  }
}

// Per-thread state of one tile: the quantised coordinates of the thread's 4 pixels.
struct WtPixels {
  int sx[4], sy[4];
  int fid[4];  // weight-table row, -1 = pixel outside the output image
};

struct WtRegion {
  int rx0, ry0;  // source coordinates of entry (0,0)
  int pitch;     // entries per staged row (multiple of 4)
  int rows;
  bool staged;
};

// Phase A (second half): bounding box of the group's taps -> region geometry.  Contains the
// group barrier that also orders the previous tile's reads of `region` before this tile's writes.
__device__ __forceinline__ WtRegion wt_bbox(WtSmem& S, int grp, int gw, int lane, const WtPixels& px) {
  int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = (int)0x80000000, mxy = (int)0x80000000;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (px.fid[k] >= 0) {
      mnx = min(mnx, px.sx[k]);
      mxx = max(mxx, px.sx[k]);
      mny = min(mny, px.sy[k]);
      mxy = max(mxy, px.sy[k]);
    }
  mnx = __reduce_min_sync(0xffffffffu, mnx);
  mny = __reduce_min_sync(0xffffffffu, mny);
  mxx = __reduce_max_sync(0xffffffffu, mxx);
  mxy = __reduce_max_sync(0xffffffffu, mxy);
  if (lane == 0) *reinterpret_cast<int4*>(S.red[grp][gw]) = make_int4(mnx, mny, mxx, mxy);
  wt_group_barrier(grp);
#pragma unroll
  for (int w = 0; w < 8; ++w) {
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
  R.pitch = (ew + 3) & ~3;
  R.rows = eh;
  R.staged = mnx <= mxx && ew <= 4096 && eh <= 4096 && R.pitch * eh <= kWtRegionCap;
  return R;
}

// Phase B: stage the source rectangle as dp2a pair entries.  src_lo / src_hi bound the bytes the
// word loads may touch (the whole source buffer).  Ends with the group barrier.
__device__ __forceinline__ void wt_stage(uint2* __restrict__ region, const WtRegion& R,
                                         const unsigned char* __restrict__ img, int Hs, int Ws,
                                         const unsigned char* __restrict__ src_lo,
                                         const unsigned char* __restrict__ src_hi, int gt, int grp) {
  if (R.staged) {
    const int ngr = R.pitch >> 2;
    const int items = ngr * R.rows;
    for (int i = gt; i < items; i += kWtGroupThreads) {
      const int ry = i / ngr, j = i - ry * ngr;
      const int y = R.ry0 + ry, xs = R.rx0 + 4 * j;
      unsigned r0 = 0, r1 = 0, r2 = 0, r3 = 0;  // bytes b0..b15 of pixels xs..xs+4 (15 used)
      if ((unsigned)y < (unsigned)Hs && xs > -5 && xs < Ws) {
        const unsigned char* a = img + ((int64_t)y * Ws + xs) * 3;
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(a) & 3);
        const unsigned* q = reinterpret_cast<const unsigned*>(a - mis);
        if (xs >= 0 && xs + 5 <= Ws && reinterpret_cast<const unsigned char*>(q) >= src_lo &&
            reinterpret_cast<const unsigned char*>(q + 5) <= src_hi) {
          const unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3);
          const unsigned w4 = mis >= 2 ? __ldg(q + 4) : 0u;
          const unsigned sh = mis * 8;
          r0 = __funnelshift_r(w0, w1, sh);
          r1 = __funnelshift_r(w1, w2, sh);
          r2 = __funnelshift_r(w2, w3, sh);
          r3 = __funnelshift_r(w3, w4, sh);
        } else {
          // row ends / buffer ends: byte loads, pixels outside the row are 0
          unsigned r[4] = {0, 0, 0, 0};
#pragma unroll
          for (int n = 0; n < 15; ++n) {
            const int xx = xs + n / 3;
            if ((unsigned)xx < (unsigned)Ws) r[n >> 2] |= (unsigned)a[n] << (8 * (n & 3));
          }
          r0 = r[0]; r1 = r[1]; r2 = r[2]; r3 = r[3];
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
__device__ __forceinline__ unsigned wt_pixel(const WtSmem& S, const uint2* __restrict__ region, const WtRegion& R,
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

// A warp's 32 pixels (3 bytes each, `v` = lane's pixel) -> 24 aligned words of the output row.
// All 32 lanes must call; lanes 0..23 return their word.
__device__ __forceinline__ unsigned wt_pack_row(unsigned v, int lane) {
  const int j = lane < 24 ? lane : 23;
  const int p = j + j / 3, o = j - (j / 3) * 3;  // first pixel and channel offset of word j
  const unsigned a = __shfl_sync(0xffffffffu, v, p);
  const unsigned b = __shfl_sync(0xffffffffu, v, p + 1);
  return (a >> (8 * o)) | (b << (24 - 8 * o));
}

// Store one 32-pixel row segment.  `full` = all 32 pixels inside the image and the segment word-aligned.
__device__ __forceinline__ void wt_store_row(unsigned char* __restrict__ o, unsigned v, int lane, bool full,
                                             bool lane_valid) {
  const unsigned word = wt_pack_row(v, lane);
  if (full) {
    if (lane < 24) __stcs(reinterpret_cast<unsigned*>(o) + lane, word);
  } else if (lane_valid) {
    o[3 * lane] = (unsigned char)(v & 0xff);
    o[3 * lane + 1] = (unsigned char)((v >> 8) & 0xff);
    o[3 * lane + 2] = (unsigned char)((v >> 16) & 0xff);
  }
}

}  // namespace sdof
