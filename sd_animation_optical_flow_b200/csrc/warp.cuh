// Device-side building blocks of the backward warp (shared by warp.cu and fused.cu).
#pragma once

#include "sdof_common.cuh"

namespace sdof {

// weight tables live in global memory (L1-resident; a warp's pixels mostly share a row).
// Uploaded once per device; kernels receive the device pointers as arguments.
struct CubicTables {
  const int16_t* i16;  // [1024][16]
  const float* f32;    // [1024][16]
};
int get_cubic_tables(CubicTables* out);  // returns sdof_status

// cvRound(v * 32): round half to even; NaN / out-of-int-range -> INT_MIN (what
// cvtss2si returns, which is what OpenCV's cvRound / v_round do on x86).
__device__ __forceinline__ int cv_round_x32(float v) {
  float s = v * 32.0f;
  if (!(s >= -2147483648.0f && s < 2147483648.0f)) return (int)0x80000000;
  return __float2int_rn(s);
}

__device__ __forceinline__ int sat_s16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

// Sampling coordinate exactly as the reference builds its remap maps:
//   float32(float64 grid + flow) (pdcnet_of.py:35-40) / float32(-flow + arange) (ofgen.py:39-41).
// A single fp32 add gives the same bits: the grid value is an integer < 2^24 and flow has a 24-bit
// significand, so whenever the exact sum needs more than the 53 bits of a double, flow is smaller than
// 2^-5 float-ulps of the grid value and cannot reach a rounding midpoint -- no double-rounding case exists.
__device__ __forceinline__ float map_coord(int grid, float flow, float sign) {
  return __fadd_rn((float)grid, sign * flow);
}

// flat pixel index -> (batch, y, x); 32-bit divisions whenever the sizes allow
__device__ __forceinline__ void decompose_pixel(int64_t p, int64_t hw, int W, int& b, int& y, int& x) {
  if (p < 0x7fffffffLL && hw < 0x7fffffffLL) {
    const unsigned up = (unsigned)p, uhw = (unsigned)hw;
    const unsigned ub = up / uhw;
    const unsigned rem = up - ub * uhw;
    y = (int)(rem / (unsigned)W);
    x = (int)(rem - (unsigned)y * (unsigned)W);
    b = (int)ub;
  } else {
    const int64_t lb = p / hw;
    const int64_t rem = p - lb * hw;
    y = (int)(rem / W);
    x = (int)(rem - (int64_t)y * W);
    b = (int)lb;
  }
}

struct FixedCoord {
  int sx, sy;  // top-left tap (integer part - 1)
  int fidx;    // (fy*32 + fx) row of the weight table
};

__device__ __forceinline__ FixedCoord fixed_coord(float mx, float my) {
  int qx = cv_round_x32(mx), qy = cv_round_x32(my);
  FixedCoord fc;
  fc.sx = sat_s16(qx >> 5) - 1;
  fc.sy = sat_s16(qy >> 5) - 1;
  fc.fidx = ((qy & 31) << 5) | (qx & 31);
  return fc;
}

__device__ __forceinline__ int dp2a_lo_s16u8(unsigned w, unsigned px, int acc) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(acc));
  return d;
}
__device__ __forceinline__ int dp2a_hi_s16u8(unsigned w, unsigned px, int acc) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(acc));
  return d;
}

__device__ __forceinline__ unsigned char cast_q15_u8(int sum) {
  int v = (sum + (1 << 14)) >> 15;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One output pixel of the 3-channel u8 bicubic remap.  img = image base of this batch
// item, buf_end = one past the last byte of the whole source buffer.  Returns the three
// channels packed in the low 24 bits.
__device__ __forceinline__ unsigned cubic_u8_c3(const int16_t* __restrict__ tab, const unsigned char* __restrict__ img,
                                                const unsigned char* __restrict__ buf_end, int Hs, int Ws,
                                                FixedCoord fc) {
  const uint4* wrow = reinterpret_cast<const uint4*>(tab + fc.fidx * 16);
  const uint4 wa = __ldg(wrow);      // rows ky=0,1 : (w00,w01)(w02,w03)(w10,w11)(w12,w13)
  const uint4 wb = __ldg(wrow + 1);  // rows ky=2,3
  const unsigned wlo[4] = {wa.x, wa.z, wb.x, wb.z};
  const unsigned whi[4] = {wa.y, wa.w, wb.y, wb.w};
  int a0 = 0, a1 = 0, a2 = 0;
  const bool interior = Ws > 3 && Hs > 3 && (unsigned)fc.sx < (unsigned)(Ws - 3) && (unsigned)fc.sy < (unsigned)(Hs - 3);
  const int row_bytes = Ws * 3;
  // a single image is < 2 GB (H, W <= 32767), so byte offsets inside it fit in 32 bits
  const unsigned char* p0 = img + (fc.sy * Ws + fc.sx) * 3;
  // the aligned 16-byte window of the last tap row must stay inside the buffer
  const bool fast = interior && (p0 + 3 * row_bytes + 16 <= buf_end);
  if (fast) {
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(p0) & 3);
    const unsigned char* q0 = p0 - mis;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      // rows are row_bytes apart: the misalignment of row ky is (mis + ky*row_bytes) & 3
      const unsigned off = mis + (unsigned)(ky * row_bytes);
      const unsigned m = off & 3;
      const unsigned* q = reinterpret_cast<const unsigned*>(q0 + (off - m));
      const unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
      const unsigned w3 = m ? __ldg(q + 3) : 0u;
      const unsigned sh = m * 8;
      const unsigned r0 = __funnelshift_r(w0, w1, sh);  // bytes b0..b3  (pixel k channel c = b[3k+c])
      const unsigned r1 = __funnelshift_r(w1, w2, sh);  // b4..b7
      const unsigned r2 = __funnelshift_r(w2, w3, sh);  // b8..b11
      const unsigned x0 = __byte_perm(__byte_perm(r0, r1, 0x0630), r2, 0x5210);  // b0 b3 b6 b9
      const unsigned x1 = __byte_perm(__byte_perm(r0, r1, 0x0741), r2, 0x6210);  // b1 b4 b7 b10
      const unsigned x2 = __byte_perm(__byte_perm(r0, r1, 0x0052), r2, 0x7410);  // b2 b5 b8 b11
      a0 = dp2a_hi_s16u8(whi[ky], x0, dp2a_lo_s16u8(wlo[ky], x0, a0));
      a1 = dp2a_hi_s16u8(whi[ky], x1, dp2a_lo_s16u8(wlo[ky], x1, a1));
      a2 = dp2a_hi_s16u8(whi[ky], x2, dp2a_lo_s16u8(wlo[ky], x2, a2));
    }
  } else {
    // border / out-of-image: taps outside contribute cval = 0 (BORDER_CONSTANT)
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int yy = fc.sy + ky;
      if ((unsigned)yy >= (unsigned)Hs) continue;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int xx = fc.sx + kx;
        if ((unsigned)xx >= (unsigned)Ws) continue;
        const unsigned pair = (kx < 2) ? wlo[ky] : whi[ky];
        const int w = (int)(short)((kx & 1) ? (pair >> 16) : (pair & 0xffffu));
        const unsigned char* p = img + (yy * Ws + xx) * 3;
        a0 += w * (int)p[0];
        a1 += w * (int)p[1];
        a2 += w * (int)p[2];
      }
    }
  }
  return (unsigned)cast_q15_u8(a0) | ((unsigned)cast_q15_u8(a1) << 8) | ((unsigned)cast_q15_u8(a2) << 16);
}

}  // namespace sdof
