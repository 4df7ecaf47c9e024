// CUDA-core fp32 correlation volume (SDOF_PREC_FP32), 2x2 pyramid pooling, and the
// channels-last feature pooling used by AlternateCorrBlock.
//
// The fp32 FMA volume is the on-device exact-arithmetic checker for the tcgen05 path
// (corr_tc.cu) and the fallback for shapes the TMA path rejects; it is not the fast path.
#include "corr.cuh"

namespace sdof {

// C[m][n] = scale * sum_k A[m][k] * B[n][k];  A [M][K], B [N][K] row-major (K contiguous).
// 64x64 tile, BK = 16, 256 threads, 4x4 micro-tile.  Output goes to pyramid level 0:
// element (m, n) -> out[m*pitch + (n / w2)*wp + n % w2].
constexpr int kSB = 64, kSK = 16;

__global__ void __launch_bounds__(256) corr_volume_fp32_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                               int M, int N, int K, int64_t a_bstride,
                                                               int64_t b_bstride, float scale, bool scale_is_div,
                                                               float divisor, float* __restrict__ out, int64_t pitch,
                                                               int w2, int wp, int64_t out_bstride) {
  __shared__ float As[kSK][kSB + 4];
  __shared__ float Bs[kSK][kSB + 4];
  const int b = blockIdx.z;
  A += b * a_bstride;
  Bm += b * b_bstride;
  out += b * out_bstride;
  const int m0 = blockIdx.y * kSB, n0 = blockIdx.x * kSB;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  // each thread loads one float4 of A and one of B per K-slab: row = tid/4, k4 = (tid%4)*4
  const int lr = threadIdx.x / 4, lk = (threadIdx.x % 4) * 4;
  for (int k0 = 0; k0 < K; k0 += kSK) {
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (m0 + lr < M && k0 + lk < K) va = *reinterpret_cast<const float4*>(A + (int64_t)(m0 + lr) * K + k0 + lk);
    if (n0 + lr < N && k0 + lk < K) vb = *reinterpret_cast<const float4*>(Bm + (int64_t)(n0 + lr) * K + k0 + lk);
    As[lk + 0][lr] = va.x; As[lk + 1][lr] = va.y; As[lk + 2][lr] = va.z; As[lk + 3][lr] = va.w;
    Bs[lk + 0][lr] = vb.x; Bs[lk + 1][lr] = vb.y; Bs[lk + 2][lr] = vb.z; Bs[lk + 3][lr] = vb.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float v = scale_is_div ? __fdiv_rn(acc[i][j], divisor) : acc[i][j] * scale;
      out[(int64_t)m * pitch + (int64_t)(n / w2) * wp + (n % w2)] = v;
    }
  }
}

// level l -> level l+1: avg_pool2d(2,2) floor, ATen's sum order ((a00+a01)+a10)+a11.
__global__ void __launch_bounds__(256) pool_level_kernel(const float* __restrict__ in, int64_t in_pitch, int in_wp,
                                                         float* __restrict__ out, int64_t out_pitch, int out_wp, int ho,
                                                         int wo, int64_t rows) {
  const int64_t per_row = (int64_t)ho * wo;
  const int64_t total = rows * per_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / per_row;
    const int rem = (int)(i - row * per_row);
    const int y = rem / wo, x = rem - y * wo;
    const float* p = in + row * in_pitch + (int64_t)(2 * y) * in_wp + 2 * x;
    float s = __fadd_rn(p[0], p[1]);
    s = __fadd_rn(s, p[in_wp]);
    s = __fadd_rn(s, p[in_wp + 1]);
    out[row * out_pitch + (int64_t)y * out_wp + x] = s * 0.25f;
  }
}

int launch_pool_levels(float* pyramid, const sdof_pyramid_layout& lay, int64_t rows, int first_level, cudaStream_t st) {
  for (int l = first_level; l + 1 < lay.levels; ++l) {
    const int ho = lay.h[l + 1], wo = lay.w[l + 1];
    const int64_t total = rows * ho * wo;
    if (total == 0) continue;
    pool_level_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(pyramid + lay.offset[l], lay.pitch[l], lay.wp[l],
                                                               pyramid + lay.offset[l + 1], lay.pitch[l + 1],
                                                               lay.wp[l + 1], ho, wo, rows);
    SDOF_LAUNCH_CHECK("pool_level_kernel");
  }
  return SDOF_OK;
}

int launch_corr_volume_fp32(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, float* pyramid,
                            const sdof_pyramid_layout& lay, cudaStream_t st) {
  const int n2 = h2 * w2;
  const float sq = sqrtf((float)C);
  const float inv = 1.0f / sq;
  // 1/sqrt(C) is exact (and so is the multiply) when C is a power of 4; otherwise divide like the reference
  const bool pow4 = (C & (C - 1)) == 0 && (__builtin_ctz(C) % 2 == 0);
  SDOF_REQUIRE(B <= 65535, "corr volume: B > 65535 not supported");
  dim3 grid(ceil_div(n2, kSB), ceil_div(n1, kSB), B);
  corr_volume_fp32_kernel<<<grid, 256, 0, st>>>(fmap1, fmap2, n1, n2, C, (int64_t)n1 * C, (int64_t)n2 * C, inv, !pow4, sq,
                                                pyramid + lay.offset[0], lay.pitch[0], w2, lay.wp[0],
                                                (int64_t)n1 * lay.pitch[0]);
  SDOF_LAUNCH_CHECK("corr_volume_fp32_kernel");
  return launch_pool_levels(pyramid, lay, (int64_t)B * n1, 0, st);
}

__global__ void __launch_bounds__(256) avgpool2_nhwc_kernel(const float4* __restrict__ in, int H, int W, int C4, int Ho,
                                                            int Wo, int64_t total, float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    int64_t t = i / C4;
    const int x = (int)(t % Wo);
    t /= Wo;
    const int y = (int)(t % Ho);
    const int64_t b = t / Ho;
    const float4* p = in + ((b * H + 2 * y) * (int64_t)W + 2 * x) * C4 + c;
    const float4 a = p[0], bq = p[C4], cq = p[(int64_t)W * C4], d = p[(int64_t)W * C4 + C4];
    float4 r;
    r.x = __fadd_rn(__fadd_rn(__fadd_rn(a.x, bq.x), cq.x), d.x) * 0.25f;
    r.y = __fadd_rn(__fadd_rn(__fadd_rn(a.y, bq.y), cq.y), d.y) * 0.25f;
    r.z = __fadd_rn(__fadd_rn(__fadd_rn(a.z, bq.z), cq.z), d.z) * 0.25f;
    r.w = __fadd_rn(__fadd_rn(__fadd_rn(a.w, bq.w), cq.w), d.w) * 0.25f;
    out[i] = r;
  }
}

}  // namespace sdof

extern "C" int sdof_avgpool2_nhwc(const float* in, int B, int H, int W, int C, float* out, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(in && out, "sdof_avgpool2_nhwc: NULL pointer");
  SDOF_REQUIRE(B >= 0 && H >= 1 && W >= 1 && C >= 4 && C % 4 == 0, "sdof_avgpool2_nhwc: bad sizes (C must be a multiple of 4)");
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "sdof_avgpool2_nhwc: pointers must be 16-byte aligned");
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const int64_t total = (int64_t)B * Ho * Wo * C4;
  if (total == 0) return SDOF_OK;
  avgpool2_nhwc_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(in), H, W, C4, Ho, Wo, total, reinterpret_cast<float4*>(out));
  SDOF_LAUNCH_CHECK("avgpool2_nhwc_kernel");
  return SDOF_OK;
}
