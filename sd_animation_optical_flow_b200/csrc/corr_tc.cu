// All-pairs correlation volume + 4-level pyramid on the 5th-gen tensor cores (SURVEY §8a C1+C2).
//
//   level0[b, m, n] = <fmap1[b,m,:], fmap2[b,n,:]> / sqrt(C)      m: source pixel, n: target pixel
//   level(l+1)      = avg_pool2d(level l, 2, 2)  over the target dims
//
// Mapping to the hardware (round-1 redesign after profiling, see profiles/README.md):
//   * The GEMM is issued TRANSPOSED: D[n, m] with an 8x16 SPATIAL patch of TARGET pixels on the 128
//     TMEM lanes (MMA M) and 256 consecutive SOURCE pixels on the accumulator columns (MMA N).
//     Both operands are K-major in HBM (channels-last features), so TMA loads 64-byte-swizzled
//     K-slabs straight from the feature tensors: A = 4-D box (16 ch-floats, 16 x, 8 y, 1 b) of fmap2,
//     B = 3-D box (16 ch-floats, 256 px, 1 b) of fmap1; 5-stage mbarrier ring, OOB zero-filled.
//   * One elected thread issues tcgen05.mma (M=128, N=256, kind::tf32 or kind::f16) into one of two
//     256-column TMEM accumulators; tcgen05.commit releases smem stages / publishes the accumulator.
//   * TWO epilogue warpgroups alternate tiles (each owns one accumulator), so a tile's epilogue may
//     take two MMA periods.  A warp reads 32 lanes x 32 columns with tcgen05.ld: lane = target pixel
//     (2 rows x 16 cols of the patch), register = source pixel.  Level 0 goes straight from
//     registers to HBM -- for each source pixel the warp writes two 64-byte row segments of that
//     pixel's map (no shared-memory staging, no TMA store, no proxy fences).  Level 1 is pooled
//     with three warp shuffles in ATen's summation order ((a00+a01)+a10)+a11; the level-1 tile
//     (32 KB) is exchanged through shared memory and levels 2-3 are finished by one thread per
//     source pixel.  Floor-pooled sizes / partial tiles are handled by predicates.
//   * Persistent grid (one CTA per SM), tiles ordered target-patch-fastest so concurrently running
//     CTAs complete each other's 128-byte lines in L2 and share the L2-resident operands.
//
// Roofline: 2*N1*N2*C flops on the tensor pipe vs. 4*N1*sum_l(h_l*w_l) bytes of stores to HBM
// (S config: 19.33 GFLOP vs 200.5 MB -> the kernel is HBM-store-bound).
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "corr.cuh"
#include "tc_ptx.cuh"

namespace sdof {

constexpr int kPatchY = 8;             // target patch rows   } 128 target pixels = TMEM lanes (MMA M)
constexpr int kPatchX = 16;            // target patch cols   }
constexpr int kBM = kPatchY * kPatchX;  // 128
constexpr int kBN = 256;               // source pixels per tile = accumulator columns (MMA N)
constexpr int kStages = 5;
constexpr int kAStage = kBM * kSlabBytes;  // 8192
constexpr int kBStage = kBN * kSlabBytes;  // 16384
constexpr int kEpiGroups = 2;          // epilogue warpgroups, one per TMEM accumulator
constexpr int kEpiWarps = 4 * kEpiGroups;
constexpr int kXchgBytes = 4 * kBN * 8 * 4;  // per group: level-1 tile [4 row-pairs][256 src px][8 x] fp32 = 32 KB
constexpr int kSmemOperands = kStages * (kAStage + kBStage);  // 122880
constexpr int kL2TileBytes = 2 * kBN * 4 * 4;  // per group: level-2 tile [2 rows][256 src px][4 x] fp32 = 8 KB
constexpr int kSmemEpi = kEpiGroups * (kXchgBytes + kL2TileBytes);  // 81920
constexpr int kSmemBars = 128;
constexpr int kSmemTotal = kSmemOperands + kSmemEpi + kSmemBars + 1024;  // + alignment slack
constexpr int kThreads = (kEpiWarps + 2) * 32;                // 8 epilogue warps + TMA + MMA
constexpr int kMaxTerms = 3;
constexpr int kTcLevels = 4;

struct TcMaps {
  CUtensorMap a[kMaxTerms];  // fmap2 (target patch rows)
  CUtensorMap b[kMaxTerms];  // fmap1 (source pixel rows)
  CUtensorMap out1, out2;    // pyramid levels 1 and 2, dims (x, y, source pixel, b)
};

struct TcArgs {
  int B, n1, h2, w2;
  int m_tiles, ty_tiles, tx_tiles;
  int nterms, kslabs, slab_elems;
  int levels;          // pyramid levels written by this kernel (1..4)
  float divisor;       // sqrt(C) (used when 1/sqrt(C) is not a power of two)
  int use_div;
  int debug;           // SDOF_TC_DEBUG bit mask (profiling experiments only): 1 skip L0 stores, 2 skip pooled levels,
                       // 4 skip MMA issue, 8 skip operand loads
  float* out[kTcLevels];
  long long pitch[kTcLevels];
  int wp[kTcLevels], lh[kTcLevels], lw[kTcLevels];
};

// One accumulator column (= one source pixel) of the epilogue: level-0 store, 2x2 pooling by shuffles in
// ATen's order ((a00+a01)+a10)+a11, level-1 store + hand-over to the level-2/3 stage.
template <bool kAllIn0, bool kPool>
__device__ __forceinline__ void epi_column(float v, float*& p0, long long pitch0, uint32_t in0, float*& xr, uint32_t own1) {
  if (kAllIn0)
    *p0 = v;  // 2 x 64-byte row segments per warp
  else
    st_global_pred(p0, v, in0);
  p0 += pitch0;
  if (kPool) {
    const float a01 = __shfl_xor_sync(0xffffffffu, v, 1);
    const float a10 = __shfl_xor_sync(0xffffffffu, v, 16);
    const float a11 = __shfl_xor_sync(0xffffffffu, v, 17);
    const float l1 = __fadd_rn(__fadd_rn(__fadd_rn(v, a01), a10), a11) * 0.25f;
    st_shared_pred(xr, l1, own1);  // level-1 tile [row pair][source pixel][8 x] -> TMA store + levels 2/3
    xr += 8;
  }
}

template <bool kAllIn0, bool kPool>
__device__ __forceinline__ void epi_chunk(const uint32_t (&u)[32], int ncols, bool use_div, float divisor, float*& p0,
                                          long long pitch0, uint32_t in0, float*& xr, uint32_t own1) {
  if (ncols >= 32 && !use_div) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj)
      epi_column<kAllIn0, kPool>(__uint_as_float(u[jj]), p0, pitch0, in0, xr, own1);
  } else {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      if (jj < ncols) {  // warp-uniform
        float v = __uint_as_float(u[jj]);
        if (use_div) v = __fdiv_rn(v, divisor);  // exact power-of-two scales are folded into the operand instead
        epi_column<kAllIn0, kPool>(v, p0, pitch0, in0, xr, own1);
      }
    }
  }
}

// ----------------------------------------------------------------------------- the kernel
template <bool kBf16>
__global__ void __launch_bounds__(kThreads, 1) corr_volume_tc_kernel(const __grid_constant__ TcMaps maps,
                                                                     const TcArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + kStages * kAStage;
  const uint32_t bars = base + kSmemOperands + kSmemEpi;
  const uint32_t bar_full = bars;                  // [kStages]
  const uint32_t bar_empty = bars + 8 * kStages;   // [kStages]
  const uint32_t bar_tfull = bars + 16 * kStages;  // [2]
  const uint32_t bar_tempty = bar_tfull + 16;      // [2]
  const uint32_t tmem_slot = bar_tempty + 16;      // u32
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = args.ty_tiles * args.tx_tiles;
  const int total_tiles = args.B * args.m_tiles * n_tiles;
  const int kk_total = args.nterms * args.kslabs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);  // the 4 warps of the epilogue group that owns the accumulator
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps && lane == 0) {
    for (int t = 0; t < args.nterms; ++t) {
      prefetch_tmap(&maps.a[t]);
      prefetch_tmap(&maps.b[t]);
    }
    if (args.levels > 1) prefetch_tmap(&maps.out1);
    if (args.levels > 2) prefetch_tmap(&maps.out2);
  }
  if (warp == kEpiWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  if (warp == kEpiWarps) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles;
        const int mt = (tile / n_tiles) % args.m_tiles;
        const int b = tile / (n_tiles * args.m_tiles);
        const int ty = nt / args.tx_tiles, tx = nt - ty * args.tx_tiles;
        for (int term = 0; term < args.nterms; ++term) {
          for (int k = 0; k < args.kslabs; ++k) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (args.debug & 8) {
              mbar_arrive(bar_full + 8 * stage);
            } else {
              mbar_expect_tx(bar_full + 8 * stage, kAStage + kBStage);
              tma_load_4d(smem_a + stage * kAStage, &maps.a[term], bar_full + 8 * stage, k * args.slab_elems,
                          tx * kPatchX, ty * kPatchY, b);
              tma_load_3d(smem_b + stage * kBStage, &maps.b[term], bar_full + 8 * stage, k * args.slab_elems, mt * kBN, b);
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_fmt(kBf16 ? 1u : 2u, kBM, kBN);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t ab = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * ab, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + ab * kBN;
        for (int kk = 0; kk < kk_total; ++kk) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_a + stage * kAStage);
          const uint64_t bdesc = make_smem_desc(smem_b + stage * kBStage);
#pragma unroll
          for (int j = 0; j < kMmaPerSlab; ++j)  // 32 bytes of K per MMA (UMMA_K = 8 tf32 / 16 bf16)
            if (!(args.debug & 4)) tc_mma<kBf16>(tmem_d, adesc + 2 * j, bdesc + 2 * j, idesc, (kk > 0 || j > 0) ? 1u : 0u);
          tc_commit(bar_empty + 8 * stage);  // frees the smem stage when these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(bar_tfull + 8 * ab);  // accumulator complete
      }
    }
  } else {
    // ===================================================================== epilogue: 2 groups x 4 warps
    const int grp = warp >> 2, q = warp & 3;       // q selects TMEM lanes 32q..32q+31 = patch rows 2q, 2q+1
    const int yl = lane >> 4, xl = lane & 15;
    float* xchg = reinterpret_cast<float*>(gen_base + kSmemOperands + grp * (kXchgBytes + kL2TileBytes));  // [4][kBN][8]
    float* l2tile = xchg + kXchgBytes / 4;                                                                   // [2][kBN][4]
    const uint32_t xchg_s = smem_u32(xchg), l2tile_s = smem_u32(l2tile);
    const int levels = args.levels;
    const bool do_l0 = !(args.debug & 1);
    const bool do_pool = levels > 1 && !(args.debug & 2);
    const long long pitch0 = args.pitch[0];
    const int h0 = args.lh[0], w0 = args.lw[0], wp0 = args.wp[0];
    float* const out0 = args.out[0];
    const bool use_div = args.use_div != 0;
    const float divisor = args.divisor;
    const int n1 = args.n1, m_tiles = args.m_tiles, tx_tiles = args.tx_tiles;
    const uint32_t own1 = (lane & 17) == 0;  // even column, upper row of the pair: owns a 2x2 window
    const bool issuer = q == 0 && lane == 0;  // the group's TMA-store thread
    const uint32_t bar_id = 1 + grp;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int nt = tile % n_tiles;
      const int mt = (tile / n_tiles) % m_tiles;
      const int b = tile / (n_tiles * m_tiles);
      const int ty = nt / tx_tiles, tx = nt - ty * tx_tiles;
      const int m0 = mt * kBN;
      const int mcount = min(kBN, n1 - m0);   // valid source pixels (columns) of this tile
      const long long row0 = (long long)b * n1 + m0;
      const int y = ty * kPatchY + 2 * q + yl, x = tx * kPatchX + xl;
      const uint32_t in0 = do_l0 && y < h0 && x < w0;
      float* p0 = out0 + row0 * pitch0 + (long long)y * wp0 + x;
      float* xr = xchg + (size_t)q * kBN * 8 + (xl >> 1);
      const bool all_in0 = __all_sync(0xffffffffu, in0);

      if (do_pool) {
        // the previous tile's TMA stores must have finished reading the exchange tiles before they are rewritten
        if (issuer) tma_store_wait_read0();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
      mbar_wait(bar_tfull + 8 * grp, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * kBN;
#pragma unroll 1
      for (int chunk = 0; chunk < kBN / 32; ++chunk) {
        uint32_t u[32];
        tmem_ld32(taddr + chunk * 32, u);
        tmem_ld_wait(u);
        if (chunk == kBN / 32 - 1) {
          // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * grp);
        }
        const int ncols = mcount - chunk * 32;
        if (ncols > 0) {
          if (do_pool) {
            if (all_in0)
              epi_chunk<true, true>(u, ncols, use_div, divisor, p0, pitch0, in0, xr, own1);
            else
              epi_chunk<false, true>(u, ncols, use_div, divisor, p0, pitch0, in0, xr, own1);
          } else {
            epi_chunk<false, false>(u, ncols, use_div, divisor, p0, pitch0, in0, xr, own1);
          }
        }
      }
      if (do_pool) {
        // level 1: the exchange tile [r][source pixel][8 x] goes out as 4 TMA boxes (one per level-1 row);
        // the tensor map clips columns / rows / source pixels beyond the (floor-pooled) level size
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (issuer) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            tma_store_4d(&maps.out1, xchg_s + r * (kBN * 32), tx * (kPatchX / 2), ty * (kPatchY / 2) + r, m0, b);
          tma_store_commit();
        }
        if (levels > 2) {
          // levels 2 and 3: one thread per source pixel finishes the 4x8 level-1 tile of that pixel
          const int t = q * 32 + lane;
#pragma unroll
          for (int rep = 0; rep < kBN / 128; ++rep) {
            const int j = t + rep * 128;
            float l1v[4][8];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float4 lo = *reinterpret_cast<const float4*>(xchg + ((size_t)r * kBN + j) * 8);
              const float4 hi = *reinterpret_cast<const float4*>(xchg + ((size_t)r * kBN + j) * 8 + 4);
              l1v[r][0] = lo.x; l1v[r][1] = lo.y; l1v[r][2] = lo.z; l1v[r][3] = lo.w;
              l1v[r][4] = hi.x; l1v[r][5] = hi.y; l1v[r][6] = hi.z; l1v[r][7] = hi.w;
            }
            float l2v[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                l2v[r][i] = pool4(l1v[2 * r][2 * i], l1v[2 * r][2 * i + 1], l1v[2 * r + 1][2 * i], l1v[2 * r + 1][2 * i + 1]);
              *reinterpret_cast<float4*>(l2tile + ((size_t)r * kBN + j) * 4) = make_float4(l2v[r][0], l2v[r][1], l2v[r][2], l2v[r][3]);
            }
            if (levels > 3 && j < mcount && ty < args.lh[3]) {
              float* p3 = args.out[3] + (row0 + j) * args.pitch[3] + (long long)ty * args.wp[3];
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int x3 = tx * 2 + i;
                if (x3 < args.lw[3]) p3[x3] = pool4(l2v[0][2 * i], l2v[0][2 * i + 1], l2v[1][2 * i], l2v[1][2 * i + 1]);
              }
            }
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (issuer) {
            tma_store_4d(&maps.out2, l2tile_s, tx * (kPatchX / 4), ty * (kPatchY / 4), m0, b);
            tma_store_4d(&maps.out2, l2tile_s + kBN * 16, tx * (kPatchX / 4), ty * (kPatchY / 4) + 1, m0, b);
            tma_store_commit();
          }
        }
      }
    }
    if (issuer) tma_store_wait_all();  // global writes complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ----------------------------------------------------------------------------- operand preparation
// 3xTF32: x = hi + lo with hi = tf32(x), lo = tf32(x - hi); products hi*hi + hi*lo + lo*hi keep
// ~21 mantissa bits.  Both parts are exactly representable in tf32, so the tensor core's own
// fp32->tf32 conversion cannot change them.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ x, int64_t n4, float scale,
                                                         float4* __restrict__ hi, float4* __restrict__ lo) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const float in[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};  // scale is a power of two: exact
    float h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[k]));
      h[k] = __uint_as_float(hb);
      const float rem = __fsub_rn(in[k], h[k]);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
      l[k] = __uint_as_float(lb);
    }
    hi[i] = make_float4(h[0], h[1], h[2], h[3]);
    lo[i] = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// Plain TF32 mode: round the features to tf32 (nearest, ties away) once, so the result does not
// depend on how the tensor core would truncate raw fp32 bits (truncation costs ~10x in flow error).
__global__ void __launch_bounds__(256) round_tf32_kernel(const float4* __restrict__ x, int64_t n4, float scale,
                                                         float4* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    uint32_t a, b, c, d;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(v.x * scale));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.y * scale));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(v.z * scale));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(v.w * scale));
    y[i] = make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
  }
}

__global__ void __launch_bounds__(256) to_bf16_kernel(const float4* __restrict__ x, int64_t n4, float scale,
                                                      uint2* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x * scale, v.y * scale), p1 = __floats2bfloat162_rn(v.z * scale, v.w * scale);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&p0);
    o.y = *reinterpret_cast<const uint32_t*>(&p1);
    y[i] = o;
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* ptr, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle sw, const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SDOF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SDOF_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return SDOF_OK;
}

int64_t corr_tc_workspace_bytes(int B, int n1, int n2, int C, int precision) {
  const int64_t e1 = (int64_t)B * n1 * C, e2 = (int64_t)B * n2 * C;
  auto al = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
  if (precision == SDOF_PREC_3XTF32) return 2 * al(e1 * 4) + 2 * al(e2 * 4);
  if (precision == SDOF_PREC_BF16) return al(e1 * 2) + al(e2 * 2);
  if (precision == SDOF_PREC_TF32) return al(e1 * 4) + al(e2 * 4);
  return 0;
}

int launch_corr_volume_tc(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, int precision,
                          float* pyramid, const sdof_pyramid_layout& lay, void* workspace, int64_t workspace_bytes,
                          cudaStream_t st) {
  const bool bf16 = precision == SDOF_PREC_BF16;
  const int n2 = h2 * w2;
  if (C % (bf16 ? 8 : 4) != 0) return SDOF_ERR_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(fmap1) | reinterpret_cast<uintptr_t>(fmap2) | reinterpret_cast<uintptr_t>(pyramid)) & 15) != 0)
    return SDOF_ERR_UNSUPPORTED;
  if (B > 65535) return SDOF_ERR_UNSUPPORTED;
  const int64_t need = corr_tc_workspace_bytes(B, n1, n2, C, precision);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need))
    return fail(SDOF_ERR_INVALID, "sdof_corr_volume_pyramid: workspace of %lld bytes required, got %lld", (long long)need,
                (long long)workspace_bytes);
  if (need > 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return fail(SDOF_ERR_INVALID, "sdof_corr_volume_pyramid: workspace must be 256-byte aligned");

  // levels this kernel writes: at most 4 and only those with a non-empty map
  int tc_levels = 0;
  while (tc_levels < lay.levels && tc_levels < kTcLevels && lay.h[tc_levels] >= 1 && lay.w[tc_levels] >= 1) ++tc_levels;
  if (tc_levels < 1) return SDOF_ERR_UNSUPPORTED;

  const int64_t e1 = (int64_t)B * n1 * C, e2 = (int64_t)B * n2 * C;
  auto al = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
  // 1/sqrt(C) is a power of two when C is a power of 4: fold it (exactly) into the A operand during the
  // rounding pre-pass; otherwise the epilogue divides by sqrt(C) like the reference
  const bool pow4 = (C & (C - 1)) == 0 && (__builtin_ctz(C) % 2 == 0);
  const float a_scale = pow4 ? 1.0f / sqrtf((float)C) : 1.0f;
  const void* a_ptr[kMaxTerms];
  const void* b_ptr[kMaxTerms];
  int nterms = 1;
  if (precision == SDOF_PREC_3XTF32) {
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    float* a_hi = reinterpret_cast<float*>(w);
    float* a_lo = reinterpret_cast<float*>(w + al(e1 * 4));
    float* b_hi = reinterpret_cast<float*>(w + 2 * al(e1 * 4));
    float* b_lo = reinterpret_cast<float*>(w + 2 * al(e1 * 4) + al(e2 * 4));
    split_tf32_kernel<<<grid_for(e1 / 4, 256, 8), 256, 0, st>>>(reinterpret_cast<const float4*>(fmap1), e1 / 4, a_scale,
                                                               reinterpret_cast<float4*>(a_hi), reinterpret_cast<float4*>(a_lo));
    SDOF_LAUNCH_CHECK("split_tf32_kernel");
    split_tf32_kernel<<<grid_for(e2 / 4, 256, 8), 256, 0, st>>>(reinterpret_cast<const float4*>(fmap2), e2 / 4, 1.0f,
                                                               reinterpret_cast<float4*>(b_hi), reinterpret_cast<float4*>(b_lo));
    SDOF_LAUNCH_CHECK("split_tf32_kernel");
    // small terms first, the dominant hi*hi product last
    a_ptr[0] = a_lo; b_ptr[0] = b_hi;
    a_ptr[1] = a_hi; b_ptr[1] = b_lo;
    a_ptr[2] = a_hi; b_ptr[2] = b_hi;
    nterms = 3;
  } else if (bf16) {
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    void* a16 = w;
    void* b16 = w + al(e1 * 2);
    to_bf16_kernel<<<grid_for(e1 / 4, 256, 8), 256, 0, st>>>(reinterpret_cast<const float4*>(fmap1), e1 / 4, a_scale,
                                                            reinterpret_cast<uint2*>(a16));
    SDOF_LAUNCH_CHECK("to_bf16_kernel");
    to_bf16_kernel<<<grid_for(e2 / 4, 256, 8), 256, 0, st>>>(reinterpret_cast<const float4*>(fmap2), e2 / 4, 1.0f,
                                                            reinterpret_cast<uint2*>(b16));
    SDOF_LAUNCH_CHECK("to_bf16_kernel");
    a_ptr[0] = a16;
    b_ptr[0] = b16;
  } else {
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    float* a_r = reinterpret_cast<float*>(w);
    float* b_r = reinterpret_cast<float*>(w + al(e1 * 4));
    round_tf32_kernel<<<grid_for(e1 / 4, 256, 8), 256, 0, st>>>(reinterpret_cast<const float4*>(fmap1), e1 / 4, a_scale,
                                                               reinterpret_cast<float4*>(a_r));
    SDOF_LAUNCH_CHECK("round_tf32_kernel");
    round_tf32_kernel<<<grid_for(e2 / 4, 256, 8), 256, 0, st>>>(reinterpret_cast<const float4*>(fmap2), e2 / 4, 1.0f,
                                                               reinterpret_cast<float4*>(b_r));
    SDOF_LAUNCH_CHECK("round_tf32_kernel");
    a_ptr[0] = a_r;
    b_ptr[0] = b_r;
  }

  const int es = bf16 ? 2 : 4;
  const int slab_elems = kSlabBytes / es;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = kSlabBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  for (int t = 0; t < nterms; ++t) {
    {  // MMA A operand (TMEM lanes): an 8x16 spatial patch of fmap2
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w2, (cuuint64_t)h2, (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)w2 * C * es, (cuuint64_t)n2 * C * es};
      cuuint32_t box[4] = {(cuuint32_t)slab_elems, (cuuint32_t)kPatchX, (cuuint32_t)kPatchY, 1};
      if ((rc = encode_map(&maps.a[t], dt, 4, b_ptr[t], dims, strides, box, sw, "fmap2"))) return rc;
    }
    {  // MMA B operand (accumulator columns): 256 consecutive source pixels of fmap1
      cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)n1, (cuuint64_t)B};
      cuuint64_t strides[2] = {(cuuint64_t)C * es, (cuuint64_t)n1 * C * es};
      cuuint32_t box[3] = {(cuuint32_t)slab_elems, (cuuint32_t)kBN, 1};
      if ((rc = encode_map(&maps.b[t], dt, 3, a_ptr[t], dims, strides, box, sw, "fmap1"))) return rc;
    }
  }

  // pooled levels 1 and 2 leave through TMA: dims (x, y, source pixel, b), one box = one level row of 256 source pixels
  for (int l = 1; l < tc_levels && l <= 2; ++l) {
    cuuint64_t dims[4] = {(cuuint64_t)lay.w[l], (cuuint64_t)lay.h[l], (cuuint64_t)n1, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)lay.wp[l] * 4, (cuuint64_t)lay.pitch[l] * 4, (cuuint64_t)n1 * lay.pitch[l] * 4};
    cuuint32_t box[4] = {(cuuint32_t)(l == 1 ? kPatchX / 2 : kPatchX / 4), 1, (cuuint32_t)kBN, 1};
    if ((rc = encode_map(l == 1 ? &maps.out1 : &maps.out2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, pyramid + lay.offset[l], dims,
                         strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, "pyramid level")))
      return rc;
  }

  TcArgs args;
  args.B = B;
  args.n1 = n1;
  args.h2 = h2;
  args.w2 = w2;
  args.m_tiles = ceil_div(n1, kBN);
  args.ty_tiles = ceil_div(h2, kPatchY);
  args.tx_tiles = ceil_div(w2, kPatchX);
  args.nterms = nterms;
  args.slab_elems = slab_elems;
  args.kslabs = ceil_div(C, slab_elems);
  args.levels = tc_levels;
  args.divisor = sqrtf((float)C);
  for (int i = 0; i < kTcLevels; ++i) {
    const int li = i < tc_levels ? i : 0;
    args.out[i] = pyramid + lay.offset[li];
    args.pitch[i] = lay.pitch[li];
    args.wp[i] = lay.wp[li];
    args.lh[i] = lay.h[li];
    args.lw[i] = lay.w[li];
  }
  args.use_div = !pow4;
  {
    const char* dbg = getenv("SDOF_TC_DEBUG");
    args.debug = dbg ? atoi(dbg) : 0;
  }
  const int64_t total_tiles = (int64_t)B * args.m_tiles * args.ty_tiles * args.tx_tiles;
  if (total_tiles > 0x7fffffff) return SDOF_ERR_UNSUPPORTED;
  const int grid = (int)(total_tiles < sm_count() ? total_tiles : sm_count());
  if (bf16) {
    SDOF_CUDA(cudaFuncSetAttribute(corr_volume_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    corr_volume_tc_kernel<true><<<grid, kThreads, kSmemTotal, st>>>(maps, args);
  } else {
    SDOF_CUDA(cudaFuncSetAttribute(corr_volume_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    corr_volume_tc_kernel<false><<<grid, kThreads, kSmemTotal, st>>>(maps, args);
  }
  SDOF_LAUNCH_CHECK("corr_volume_tc_kernel");
  // pyramid levels beyond the four fused ones (or beyond an empty level) come from the pooling kernel
  if (tc_levels < lay.levels) return launch_pool_levels(pyramid, lay, (int64_t)B * n1, tc_levels - 1, st);
  return SDOF_OK;
}

}  // namespace sdof

extern "C" {

int64_t sdof_corr_volume_workspace_bytes(int B, int h1, int w1, int h2, int w2, int C, int levels, int precision) {
  using namespace sdof;
  if (precision == SDOF_PREC_FP16 || precision == SDOF_PREC_BF16) {
    // resident kernel; shapes it cannot take fall back to the streaming kernel (tf32 / bf16 copies)
    const int64_t res = corr_res_workspace_bytes(B, h1 * w1, h2, w2, C, levels);
    const int64_t str = corr_tc_workspace_bytes(B, h1 * w1, h2 * w2, C, precision == SDOF_PREC_FP16 ? SDOF_PREC_TF32 : SDOF_PREC_BF16);
    return res > str ? res : str;
  }
  return corr_tc_workspace_bytes(B, h1 * w1, h2 * w2, C, precision);
}

int sdof_corr_prepare_operands(const float* fmap1, const float* fmap2, int B, int h1, int w1, int h2, int w2, int C,
                               int levels, int precision, int parts, void* workspace, int64_t workspace_bytes,
                               sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(parts >= 1 && parts <= 3, "sdof_corr_prepare_operands: parts must be 1, 2 or 3");
  SDOF_REQUIRE((!(parts & 1) || fmap1) && (!(parts & 2) || fmap2) && workspace, "sdof_corr_prepare_operands: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h1 >= 1 && w1 >= 1 && h2 >= 1 && w2 >= 1 && C >= 4 && C % 4 == 0, "sdof_corr_prepare_operands: bad sizes");
  if (precision != SDOF_PREC_FP16 && precision != SDOF_PREC_BF16)
    return fail(SDOF_ERR_UNSUPPORTED, "sdof_corr_prepare_operands: only the FP16 / BF16 paths have prepared operands");
  sdof_pyramid_layout lay;
  int rc = sdof_corr_pyramid_layout((int64_t)B * h1 * w1, h2, w2, levels, &lay);
  if (rc) return rc;
  if (B == 0) return SDOF_OK;
  rc = launch_corr_prepare_resident(fmap1, fmap2, B, h1 * w1, h2, w2, C, precision == SDOF_PREC_FP16 ? 0 : 1, lay, workspace,
                                    workspace_bytes, parts, as_stream(stream));
  if (rc == SDOF_ERR_UNSUPPORTED) return fail(rc, "sdof_corr_prepare_operands: shape not supported by the resident kernel (C %% 8 == 0, C <= 256)");
  return rc;
}

int sdof_corr_pyramid_from_operands(int B, int h1, int w1, int h2, int w2, int C, int levels, int precision,
                                    float* pyramid, void* workspace, int64_t workspace_bytes, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(pyramid && workspace, "sdof_corr_pyramid_from_operands: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h1 >= 1 && w1 >= 1 && h2 >= 1 && w2 >= 1 && C >= 4 && C % 4 == 0, "sdof_corr_pyramid_from_operands: bad sizes");
  if (precision != SDOF_PREC_FP16 && precision != SDOF_PREC_BF16)
    return fail(SDOF_ERR_UNSUPPORTED, "sdof_corr_pyramid_from_operands: only the FP16 / BF16 paths have prepared operands");
  sdof_pyramid_layout lay;
  int rc = sdof_corr_pyramid_layout((int64_t)B * h1 * w1, h2, w2, levels, &lay);
  if (rc) return rc;
  if (B == 0) return SDOF_OK;
  rc = launch_corr_pyramid_prepared(B, h1 * w1, h2, w2, C, precision == SDOF_PREC_FP16 ? 0 : 1, pyramid, lay, workspace,
                                    workspace_bytes, as_stream(stream));
  if (rc == SDOF_ERR_UNSUPPORTED) return fail(rc, "sdof_corr_pyramid_from_operands: shape not supported by the resident kernel");
  return rc;
}

int sdof_corr_volume_pyramid(const float* fmap1, const float* fmap2, int B, int h1, int w1, int h2, int w2, int C,
                             int levels, int precision, float* pyramid, void* workspace, int64_t workspace_bytes,
                             sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(fmap1 && fmap2 && pyramid, "sdof_corr_volume_pyramid: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h1 >= 1 && w1 >= 1 && h2 >= 1 && w2 >= 1, "sdof_corr_volume_pyramid: bad sizes");
  SDOF_REQUIRE(C >= 4 && C % 4 == 0, "sdof_corr_volume_pyramid: C must be a positive multiple of 4, got %d", C);
  SDOF_REQUIRE(precision >= SDOF_PREC_TF32 && precision <= SDOF_PREC_FP16, "sdof_corr_volume_pyramid: unknown precision %d",
               precision);
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(fmap1) | reinterpret_cast<uintptr_t>(fmap2) | reinterpret_cast<uintptr_t>(pyramid)) & 15) == 0,
               "sdof_corr_volume_pyramid: pointers must be 16-byte aligned");
  const int n1 = h1 * w1;
  sdof_pyramid_layout lay;
  int rc = sdof_corr_pyramid_layout((int64_t)B * n1, h2, w2, levels, &lay);
  if (rc) return rc;
  if (B == 0) return SDOF_OK;
  cudaStream_t st = as_stream(stream);
  if (precision == SDOF_PREC_FP16 || precision == SDOF_PREC_BF16) {
    rc = launch_corr_pyramid_resident(fmap1, fmap2, B, n1, h2, w2, C, precision == SDOF_PREC_FP16 ? 0 : 1, pyramid, lay, workspace,
                                      workspace_bytes, st);
    if (rc != SDOF_ERR_UNSUPPORTED) return rc;
    precision = precision == SDOF_PREC_FP16 ? SDOF_PREC_TF32 : SDOF_PREC_BF16;  // e.g. C > 256: streaming kernel
  }
  if (precision != SDOF_PREC_FP32) {
    rc = launch_corr_volume_tc(fmap1, fmap2, B, n1, h2, w2, C, precision, pyramid, lay, workspace, workspace_bytes, st);
    if (rc != SDOF_ERR_UNSUPPORTED) return rc;
    // shapes the TMA path cannot express fall through to the CUDA-core kernel
  }
  return launch_corr_volume_fp32(fmap1, fmap2, B, n1, h2, w2, C, pyramid, lay, st);
}

}  // extern "C"
