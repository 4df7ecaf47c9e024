// Tap-shifted implicit-GEMM convolution on tcgen05 for RAFT's update block (RAFT/core/update.py:32-60
// SepConvGRU), with the GRU gate arithmetic fused into the epilogue.
//
// Why: at batch 1 the update block runs on M = h*w = 6144 pixels.  cuDNN's TF32 kernels spend 12 us on each 1x5 / 5x1
// gate convolution (6 GFLOP, < 1 wave, bound by L2 -> SM operand traffic in fp32) and every gate needs two more
// element-wise launches around it (sigmoid / r*h, tanh / blend).  Here one GRU pass is two kernels:
//
//   gru_zr : [z | r | q_x] = conv_taps([h | motion | flow])      K = taps*256, N = 3 x 128
//            epilogue  z -> Z (fp32),  r -> RH16 = fp16(sigmoid(r + map) * h),  q_x -> QX (fp32)
//   gru_q  : q = conv_taps(RH16)                                  K = taps*128, N = 128
//            epilogue  h' = (1 - z) h + z tanh(q + q_x + map)  -> H (fp32, in place) and HX16[:, 0:128] (fp16)
//
// * operands in fp16 (11-bit significand = what TF32 keeps of an fp32 operand), fp32 accumulation in TMEM; the hidden
//   state's master copy stays fp32, only the copy the next convolution reads is fp16;
// * implicit GEMM by TMA: activations are NHWC fp16 [B,h,w,C]; the A tile of tap (dx,dy) is the 4-D box
//   (64 channels, px, py, 1) of the 128-pixel output patch shifted by (dx,dy) -- out-of-image rows/columns are zero-filled
//   by TMA, which IS the convolution's zero padding; weights are K-major [Cout][tap][Cin] fp16, the B tile a 2-D box
//   (64 k, 128 filters); both 128-byte rows with SWIZZLE_128B, four tcgen05.mma.kind::f16 (M=128, N=128, K=16) per slab;
// * one CTA = one 128-pixel x 128-filter tile (grid = tiles x N-tiles: 48 x 3 = 144 CTAs for gru_zr at 768x512: one
//   wave), warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue (TMEM lane quarter x column half); a 6-stage
//   mbarrier ring of 32 KB (A + B) stages.
#include <cuda_fp16.h>

#include <string.h>

#include "tc_ptx.cuh"

namespace sdof {

constexpr int kCtM = 128;                 // pixels per tile (TMEM lanes)
constexpr int kCtN = 128;                 // filters per tile (accumulator columns)
constexpr int kCtSlabK = 64;              // fp16 elements per K-slab = one 128-byte swizzle row
constexpr int kCtStages = 6;
constexpr int kCtStageA = kCtM * 128;     // 16 KB
constexpr int kCtStageB = kCtN * 128;     // 16 KB
constexpr int kCtStage = kCtStageA + kCtStageB;
constexpr int kCtSmem = kCtStages * kCtStage + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kCtEpiWarps = 8;
constexpr int kCtThreads = (2 + kCtEpiWarps) * 32;
constexpr int kCtMaxTaps = 9;

enum { kEpiGruZR = 0, kEpiGruQ = 1 };

struct ConvTcMaps {
  CUtensorMap act;   // fp16 activations, dims (C, w, h, B)
  CUtensorMap wgt;   // fp16 weights, dims (K_total, Cout)
};

struct ConvTcArgs {
  int B, h, w;
  int pxs;                 // log2 patch width; patch = 2^pxs x (128 >> pxs) pixels
  int tiles_x, tiles_y;
  int taps, slabs_per_tap; // K = taps * slabs_per_tap * 64
  int tap_dx[kCtMaxTaps], tap_dy[kCtMaxTaps];
  int epi;
  // gru_zr
  const float* zrmap;      // [npix][256]  bias + conv(context) of z | r
  const float* hid;        // [npix][128]  hidden state (fp32 master)
  float* z;                // [npix][128]  sigmoid(z)
  __half* rh16;            // [npix][128]  r * h, the q convolution's input
  float* qx;               // [npix][128]  the r-independent share of q
  // gru_q
  const float* qmap;       // [npix][128]  bias + conv(context) of q
  float* h_io;             // [npix][128]  hidden state, updated in place
  __half* hx16;            // [npix][hx16_stride] fp16 GRU input; channels [0,128) = h
  int hx16_stride;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

// UMMA shared-memory descriptor: K-major operand, 128-byte rows, SWIZZLE_128B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float sigmoid_fast(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(kCtThreads, 1) conv_taps_tc_kernel(const __grid_constant__ ConvTcMaps maps,
                                                                     const __grid_constant__ ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kCtStages * kCtStage;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kCtStages, bar_tfull = bars + 16 * kCtStages, tmem_slot = bar_tfull + 8;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int t = blockIdx.x;
  const int tx = t % a.tiles_x;
  t /= a.tiles_x;
  const int ty = t % a.tiles_y;
  const int b = t / a.tiles_y;
  const int nt = blockIdx.y;
  const int px = 1 << a.pxs, py = kCtM >> a.pxs;
  const int x0 = tx * px, y0 = ty * py;
  const int total_slabs = a.taps * a.slabs_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCtStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&maps.act);
    prefetch_tmap(&maps.wgt);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kCtN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int ks = 0; ks < total_slabs; ++ks) {
        const int tap = ks / a.slabs_per_tap, slab = ks - tap * a.slabs_per_tap;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        mbar_expect_tx(bar_full + 8 * stage, kCtStage);
        const uint32_t sa = base + stage * kCtStage;
        tma_load_4d(sa, &maps.act, bar_full + 8 * stage, slab * kCtSlabK, x0 + a.tap_dx[tap], y0 + a.tap_dy[tap], b);
        tma_load_2d(sa + kCtStageA, &maps.wgt, bar_full + 8 * stage, ks * kCtSlabK, nt * kCtN);
        if (++stage == kCtStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_fmt(0u, kCtM, kCtN);   // F16 x F16 -> F32
      uint32_t stage = 0, phase = 0;
      for (int ks = 0; ks < total_slabs; ++ks) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = base + stage * kCtStage;
        const uint64_t adesc = make_smem_desc_128(sa), bdesc = make_smem_desc_128(sa + kCtStageA);
#pragma unroll
        for (int j = 0; j < kCtSlabK / 16; ++j) tc_mma<true>(tmem_base, adesc + 2 * j, bdesc + 2 * j, idesc, (ks > 0 || j > 0) ? 1u : 0u);
        tc_commit(bar_empty + 8 * stage);
        if (++stage == kCtStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit(bar_tfull);
    }
  } else {
    // ===================================================================== epilogue
    // Phase 1: TMEM -> shared memory.  warp -> TMEM lane quarter q (= warp id % 4, a hardware rule of tcgen05.ld) and column
    // half; lane = pixel of the patch, register j = filter.  The accumulator tile [128 px][128 filters] is parked in the
    // (now idle) operand ring with a row pitch of 132 floats.
    // Phase 2: the 256 epilogue threads walk the tile as [pixel][filter quad]: a warp covers the 32 quads = 512 contiguous
    // bytes of ONE pixel in every global array it touches (the gate inputs, the outputs), so all global traffic is
    // coalesced.  (A lane-per-pixel epilogue costs one L1 wavefront per lane and instruction: measured 26 us for gru_q.)
    constexpr int kPitch = kCtN + 4;
    float* tile = reinterpret_cast<float*>(gen_base);
    const int q = warp & 3, chalf = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;          // 0..255
    mbar_wait(bar_tfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int chunk = 0; chunk < 2; ++chunk) {
      const int c0 = chalf * 64 + chunk * 32;
      uint32_t u[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, u);
      tmem_ld_wait(u);
      float4* dst = reinterpret_cast<float4*>(tile + (32 * q + lane) * kPitch + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dst[j] = make_float4(__uint_as_float(u[4 * j + 0]), __uint_as_float(u[4 * j + 1]), __uint_as_float(u[4 * j + 2]),
                             __uint_as_float(u[4 * j + 3]));
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps only
    const int c4 = et & 31;                   // filter quad of this thread
#pragma unroll 4
    for (int it = 0; it < kCtM / 8; ++it) {
      const int mrow = it * 8 + (et >> 5);    // one pixel row of the tile per warp and iteration
      const int yy = y0 + (mrow >> a.pxs), xx = x0 + (mrow & (px - 1));
      if (yy >= a.h || xx >= a.w) continue;   // warp-uniform
      const long long p = ((long long)b * a.h + yy) * a.w + xx;
      const float4 acc = *reinterpret_cast<const float4*>(tile + mrow * kPitch + 4 * c4);
      if (a.epi == kEpiGruZR) {
        if (nt == 0) {          // z
          const float4 mv = __ldg(reinterpret_cast<const float4*>(a.zrmap + p * 256) + c4);
          reinterpret_cast<float4*>(a.z + p * 128)[c4] = make_float4(sigmoid_fast(acc.x + mv.x), sigmoid_fast(acc.y + mv.y),
                                                                      sigmoid_fast(acc.z + mv.z), sigmoid_fast(acc.w + mv.w));
        } else if (nt == 1) {   // r -> r * h (fp16, the q convolution's operand)
          const float4 mv = __ldg(reinterpret_cast<const float4*>(a.zrmap + p * 256 + 128) + c4);
          const float4 h4 = reinterpret_cast<const float4*>(a.hid + p * 128)[c4];
          const __half2 lo = __floats2half2_rn(sigmoid_fast(acc.x + mv.x) * h4.x, sigmoid_fast(acc.y + mv.y) * h4.y);
          const __half2 hi = __floats2half2_rn(sigmoid_fast(acc.z + mv.z) * h4.z, sigmoid_fast(acc.w + mv.w) * h4.w);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&lo);
          pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          reinterpret_cast<uint2*>(a.rh16 + p * 128)[c4] = pk;
        } else {                // q_x: kept in fp32 for the q kernel's epilogue
          reinterpret_cast<float4*>(a.qx + p * 128)[c4] = acc;
        }
      } else {                  // kEpiGruQ: h' = (1 - z) h + z tanh(q + q_x + map)
        const float4 mv = __ldg(reinterpret_cast<const float4*>(a.qmap + p * 128) + c4);
        const float4 qv = reinterpret_cast<const float4*>(a.qx + p * 128)[c4];
        const float4 zv = reinterpret_cast<const float4*>(a.z + p * 128)[c4];
        float4* hp = reinterpret_cast<float4*>(a.h_io + p * 128) + c4;
        const float4 h4 = *hp;
        float4 r;
        r.x = (1.f - zv.x) * h4.x + zv.x * tanhf(acc.x + qv.x + mv.x);
        r.y = (1.f - zv.y) * h4.y + zv.y * tanhf(acc.y + qv.y + mv.y);
        r.z = (1.f - zv.z) * h4.z + zv.z * tanhf(acc.z + qv.z + mv.z);
        r.w = (1.f - zv.w) * h4.w + zv.w * tanhf(acc.w + qv.w + mv.w);
        *hp = r;
        const __half2 lo = __floats2half2_rn(r.x, r.y), hi = __floats2half2_rn(r.z, r.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(a.hx16 + p * a.hx16_stride)[c4] = pk;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCtN) : "memory");
  }
}

// motion-encoder tail in fp16: HX16[:, 128:254] = relu(mc + mf + bias), HX16[:, 254:256] = flow  (update.py:95-96;
// replaces relu_scatter on the tensor-core GRU path: the GRU input only exists as the fp16 operand)
__global__ void __launch_bounds__(256) motion_tail16_kernel(const float4* __restrict__ mc, const float4* __restrict__ mf,
                                                            const float4* __restrict__ bias, const float2* __restrict__ flow,
                                                            __half* __restrict__ hx16, int hx16_stride, int64_t npix) {
  const int64_t total = npix * 32;   // 128 channels = 32 quads per pixel (mc / mf carry 126 + 2 zero-filter channels)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i >> 5;
    const int c4 = (int)(i & 31);
    float4 v = mc[i];
    const float4 u = mf[i], bv = __ldg(bias + c4);
    v.x = fmaxf(v.x + u.x + bv.x, 0.f);
    v.y = fmaxf(v.y + u.y + bv.y, 0.f);
    v.z = fmaxf(v.z + u.z + bv.z, 0.f);
    v.w = fmaxf(v.w + u.w + bv.w, 0.f);
    if (c4 == 31) {                  // channels 124, 125 = motion features; 126, 127 of the GRU input = the flow
      const float2 f = flow[p];
      v.z = f.x;
      v.w = f.y;
    }
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(hx16 + p * hx16_stride + 128 + 4 * c4) = pk;
  }
}

static int pick_patch_shift_ct(int h, int w) {
  int best = 5, best_tiles = 1 << 30;
  for (int s = 5; s >= 3; --s) {
    const int tiles = ceil_div(w, 1 << s) * ceil_div(h, kCtM >> s);
    if (tiles < best_tiles) {
      best_tiles = tiles;
      best = s;
    }
  }
  return best;
}

static int launch_conv_taps(const __half* act, int cin, const __half* wgt, int cout, int B, int h, int w, int horizontal,
                            ConvTcArgs& a, cudaStream_t st, const char* name) {
  if (cin % kCtSlabK != 0 || cout % kCtN != 0) return fail(SDOF_ERR_UNSUPPORTED, "%s: Cin %% 64 and Cout %% 128 required", name);
  if (((reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(wgt)) & 127) != 0)
    return fail(SDOF_ERR_INVALID, "%s: operands must be 128-byte aligned", name);
  ConvTcMaps maps;
  memset(&maps, 0, sizeof(maps));
  a.B = B;
  a.h = h;
  a.w = w;
  a.pxs = pick_patch_shift_ct(h, w);
  const int px = 1 << a.pxs, py = kCtM >> a.pxs;
  a.tiles_x = ceil_div(w, px);
  a.tiles_y = ceil_div(h, py);
  a.taps = 5;
  a.slabs_per_tap = cin / kCtSlabK;
  for (int t = 0; t < 5; ++t) {
    a.tap_dx[t] = horizontal ? t - 2 : 0;
    a.tap_dy[t] = horizontal ? 0 : t - 2;
  }
  int rc;
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)w * cin * 2, (cuuint64_t)h * w * cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)kCtSlabK, (cuuint32_t)px, (cuuint32_t)py, 1};
    if ((rc = encode_map(&maps.act, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, act, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "activations (fp16)")))
      return rc;
  }
  {
    const int K = 5 * cin;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kCtSlabK, (cuuint32_t)kCtN};
    if ((rc = encode_map(&maps.wgt, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wgt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "weights (fp16)")))
      return rc;
  }
  const long long tiles = (long long)a.tiles_x * a.tiles_y * B;
  if (tiles > 0x7fffffffLL) return fail(SDOF_ERR_UNSUPPORTED, "%s: too many tiles", name);
  static bool attr_set[64] = {};
  int dev = 0;
  SDOF_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    SDOF_CUDA(cudaFuncSetAttribute(conv_taps_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtSmem));
    attr_set[dev] = true;
  }
  conv_taps_tc_kernel<<<dim3((unsigned)tiles, (unsigned)(cout / kCtN)), kCtThreads, kCtSmem, st>>>(maps, a);
  SDOF_LAUNCH_CHECK(name);
  return SDOF_OK;
}

}  // namespace sdof

extern "C" {

int sdof_gru_zr_tc(const void* hx16, const void* w_zr16, const float* zrmap, const float* h, int B, int hh, int ww, int horizontal,
                   float* z, void* rh16, float* qx, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(hx16 && w_zr16 && zrmap && h && z && rh16 && qx, "sdof_gru_zr_tc: NULL pointer");
  SDOF_REQUIRE(B >= 0 && hh >= 1 && ww >= 1, "sdof_gru_zr_tc: bad sizes");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(zrmap) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(z) |
                 reinterpret_cast<uintptr_t>(rh16) | reinterpret_cast<uintptr_t>(qx)) & 15) == 0, "sdof_gru_zr_tc: pointers must be 16-byte aligned");
  if (B == 0) return SDOF_OK;
  ConvTcArgs a;
  memset(&a, 0, sizeof(a));
  a.epi = kEpiGruZR;
  a.zrmap = zrmap;
  a.hid = h;
  a.z = z;
  a.rh16 = reinterpret_cast<__half*>(rh16);
  a.qx = qx;
  return launch_conv_taps(reinterpret_cast<const __half*>(hx16), 256, reinterpret_cast<const __half*>(w_zr16), 384, B, hh, ww, horizontal, a,
                          as_stream(stream), "sdof_gru_zr_tc");
}

int sdof_gru_q_tc(const void* rh16, const void* w_q16, const float* qmap, const float* qx, const float* z, int B, int hh, int ww,
                  int horizontal, float* h, void* hx16, int hx16_stride, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(rh16 && w_q16 && qmap && qx && z && h && hx16, "sdof_gru_q_tc: NULL pointer");
  SDOF_REQUIRE(B >= 0 && hh >= 1 && ww >= 1 && hx16_stride >= 128 && hx16_stride % 8 == 0, "sdof_gru_q_tc: bad sizes");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(qmap) | reinterpret_cast<uintptr_t>(qx) | reinterpret_cast<uintptr_t>(z) |
                 reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(hx16)) & 15) == 0, "sdof_gru_q_tc: pointers must be 16-byte aligned");
  if (B == 0) return SDOF_OK;
  ConvTcArgs a;
  memset(&a, 0, sizeof(a));
  a.epi = kEpiGruQ;
  a.qmap = qmap;
  a.qx = const_cast<float*>(qx);
  a.z = const_cast<float*>(z);
  a.h_io = h;
  a.hx16 = reinterpret_cast<__half*>(hx16);
  a.hx16_stride = hx16_stride;
  return launch_conv_taps(reinterpret_cast<const __half*>(rh16), 128, reinterpret_cast<const __half*>(w_q16), 128, B, hh, ww, horizontal, a,
                          as_stream(stream), "sdof_gru_q_tc");
}

int sdof_motion_tail16(const float* mc, const float* mf, const float* bias, const float* flow, int64_t npix, void* hx16, int hx16_stride,
                       sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(mc && mf && bias && flow && hx16, "sdof_motion_tail16: NULL pointer");
  SDOF_REQUIRE(hx16_stride >= 256 && hx16_stride % 8 == 0, "sdof_motion_tail16: hx16_stride must be >= 256 and a multiple of 8");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(mc) | reinterpret_cast<uintptr_t>(mf) | reinterpret_cast<uintptr_t>(bias) |
                 reinterpret_cast<uintptr_t>(hx16)) & 15) == 0 && (reinterpret_cast<uintptr_t>(flow) & 7) == 0,
               "sdof_motion_tail16: pointers must be 16-byte aligned (flow: 8)");
  if (npix <= 0) return SDOF_OK;
  motion_tail16_kernel<<<grid_for(npix * 32, 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(mc), reinterpret_cast<const float4*>(mf), reinterpret_cast<const float4*>(bias),
      reinterpret_cast<const float2*>(flow), reinterpret_cast<__half*>(hx16), hx16_stride, npix);
  SDOF_LAUNCH_CHECK("motion_tail16_kernel");
  return SDOF_OK;
}

}  // extern "C"
