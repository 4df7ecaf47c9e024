// Correlation pyramid on tcgen05 with a RESIDENT operand (the fast path; SURVEY §8a C1+C2).
//
// Why this shape (profiles/README.md, round 1): streaming both operands of every 128x256 tile
// through a shared-memory ring is latency-bound -- the HBM-store rate needs ~180 KB of operand
// bytes in flight per SM.  Here a CTA keeps 256 source pixels x C channels (16-bit, 128 KB for C=256)
// RESIDENT in shared memory and only streams the 8 KB K-slabs of the target patch, so the operand
// traffic per tile drops 3x and a whole tile's worth of slabs is in flight.
//
//   level l [b, m, n] = < fmap1[b,m,:] / sqrt(C) , pool_l(fmap2)[b,n,:] >      (16-bit operands, fp32 accumulate)
//
// * Pyramid from POOLED FEATURES: avg-pooling is linear, so level l of the reference's pyramid
//   (avg_pool2d of the volume, RAFT/core/corr.py:25-27) equals the correlation with the 2^l-pooled
//   fmap2 -- the identity the reference's own AlternateCorrBlock relies on (corr.py:68-72).  Each
//   level is therefore just another set of tiles of the same GEMM and the epilogue is a pure
//   register -> HBM store.  The pooled maps are built in fp32 in ATen's summation order by the
//   pre-pass, then rounded once to the operand type.
// * Operands: fp16 (default; 11-bit significand = TF32 precision at half the bytes, saturating
//   conversion) or bf16.  1/sqrt(C) is folded exactly into fmap1 when it is a power of two.
// * GEMM issued transposed: D[n, m], a 4x32 SPATIAL patch of target pixels on the 128 TMEM lanes,
//   the 256 resident source pixels on the accumulator columns.  tcgen05.mma.kind::f16, M=128, N=256.
// * Warp roles (576 threads): warps 0-15 = epilogue (all sixteen drain every tile: 4 lane quarters x 4
//   column quarters, TMEM accumulators double-buffered), warp 16 = TMA producer, warp 17 = MMA issuer.  Epilogue: tcgen05.ld 32x32b.x32,
//   warp = one patch row, lane = 32 consecutive target x, register = source pixel; every register is
//   stored as one full 128-byte line of that source pixel's level map.
// * Persistent grid: the flattened tile list (source block major, then level, then patch) is split
//   into equal contiguous ranges, so a CTA reloads its resident block at most twice.
// * Round 2: (a) the pyramid can be STORED in fp16 (`OUT_HALF`): half the bytes of the store-bound epilogue and of every
//   lookup, and the whole 768x512 pyramid (100 MB) fits the 126 MB L2 across the 20 lookups of a pair.  Adjacent lanes
//   exchange one register per column pair, so a lane stores a packed half2 (x, x+1) of ONE source pixel's map and a
//   warp instruction still writes whole 64-byte runs.  fp16 storage rounds the volume to an 11-bit significand, which
//   is what the TF32 convolution that consumes the lookup does to it anyway.
//   (b) AUTO-RANGED 16-bit operands: a per-tensor power-of-two scale (from an abs-max pass) puts max|f| at 2^13..2^14
//   before the fp16 conversion and is undone exactly in the epilogue, so the fp16 path neither saturates on large nor
//   loses bits on tiny features (fp32-SGEMM-like range).
//   (c) operands live in two separate buffers (source / target) with a small header; the TARGET buffer (the pooled
//   fmap2 levels) can be shared by all B pairs of a call (`B2 == 1`): in the key-frame scheme fmap2 is the key frame.
//   (d) per-level patch shape (32x4, 16x8 or 8x16 target pixels) so the small pooled levels waste fewer MMA rows.
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <stdlib.h>
#include <string.h>

#include "corr.cuh"
#include "tc_ptx.cuh"

namespace sdof {

// target patch = 128 pixels; shape chosen at launch (patch_x in {32, 16}; SDOF_RES_PATCHX overrides for experiments)
constexpr int kRM = 128;                   // target pixels per tile (TMEM lanes)
constexpr int kRN = 256;                   // resident source pixels (accumulator columns)
constexpr int kRMaxSlabs = 8;              // C <= 8 * 32 = 256 channels resident
constexpr int kRAStages = 8;               // ring of 8 KB target-patch slabs
constexpr int kRAStage = kRM * kSlabBytes;  // 8192
constexpr int kRBSlab = kRN * kSlabBytes;   // 16384
constexpr int kRSmemB = kRMaxSlabs * kRBSlab;     // 131072
constexpr int kRSmemA = kRAStages * kRAStage;     // 65536
constexpr int kRMaxTilesPerCta = 512;      // tile table in shared memory (8 bytes per tile)
constexpr int kRSmemTable = kRMaxTilesPerCta * 8;
constexpr int kRSmemTotal = kRSmemB + kRSmemA + 256 + kRSmemTable + 1024;
constexpr int kREpiWarps = 16;            // 4 TMEM lane quarters x 4 column quarters
constexpr int kRThreads = (kREpiWarps + 2) * 32;
constexpr int kRLevels = SDOF_MAX_LEVELS;

struct ResMaps {
  CUtensorMap src;             // fmap1, 16-bit, dims (C, n1, B)
  CUtensorMap tgt[kRLevels];   // pooled fmap2 level l, 16-bit, dims (C, w_l, h_l, B)
};

// Operand buffers start with a header: float[0..kAmaxBlocks) partial abs-maxima (corr_absmax_kernel), float[256] = 1/scale,
// float[257] = scale (corr_prep16_kernel); the 16-bit data follow at kOpHdrBytes.
constexpr int kOpHdrBytes = 2048;
constexpr int kAmaxBlocks = 128;
constexpr int kHdrInvScale = 256, kHdrScale = 257;

struct ResArgs {
  int B, n1, m_tiles, levels, kslabs, slab_elems;
  int tile_begin_level[kRLevels + 1];  // prefix sums of patches per level inside one source block
  int tx_tiles[kRLevels];
  int lh[kRLevels], lw[kRLevels], wp[kRLevels];
  int pxs[kRLevels];                   // log2 of the level's patch width (patch = 2^pxs x 128/2^pxs target pixels)
  long long pitch[kRLevels];
  void* out[kRLevels];                 // float* or __half* (OUT_HALF)
  const float* src_hdr;                // operand headers: the epilogue undoes the operand scales
  const float* tgt_hdr;
  int total_tiles;
  float divisor;                       // sqrt(C)
  float rsqrt_c;                       // 1/sqrt(C) when that is a power of two (use_div == 0)
  int use_div;
  int fmt;  // 0 = fp16, 1 = bf16
  int tgt_shared;                      // the target operand has batch 1 and serves every pair of the call
  int debug;             // SDOF_RES_DEBUG: 1 skip stores, 4 skip MMA, 8 skip A loads, 16 skip TMEM loads (experiments only)
};

struct TileInfo {
  int blk;  // b * m_tiles + mt
  int b, mt, level, ty, tx;
};

__device__ __forceinline__ TileInfo decode_tile(const ResArgs& a, int t) {
  TileInfo ti;
  const int per_block = a.tile_begin_level[a.levels];
  ti.blk = t / per_block;
  int r = t - ti.blk * per_block;
  ti.b = ti.blk / a.m_tiles;
  ti.mt = ti.blk - ti.b * a.m_tiles;
  int l = 0;
  while (l + 1 < a.levels && r >= a.tile_begin_level[l + 1]) ++l;
  r -= a.tile_begin_level[l];
  ti.level = l;
  ti.ty = r / a.tx_tiles[l];
  ti.tx = r - ti.ty * a.tx_tiles[l];
  return ti;
}

// packed tile descriptor kept in shared memory: decoded once per CTA instead of once per warp per tile
struct PackedTile {
  uint32_t blk;       // b * m_tiles + mt
  uint32_t lvl_ty_tx; // level << 28 | ty << 14 | tx
};
__device__ __forceinline__ TileInfo unpack_tile(const ResArgs& a, PackedTile pt) {
  TileInfo ti;
  ti.blk = (int)pt.blk;
  ti.b = ti.blk / a.m_tiles;
  ti.mt = ti.blk - ti.b * a.m_tiles;
  ti.level = (int)(pt.lvl_ty_tx >> 28);
  ti.ty = (int)((pt.lvl_ty_tx >> 14) & 0x3fff);
  ti.tx = (int)(pt.lvl_ty_tx & 0x3fff);
  return ti;
}

template <bool OUT_HALF>
__global__ void __launch_bounds__(kRThreads, 1) corr_pyramid_resident_kernel(const __grid_constant__ ResMaps maps,
                                                                             const __grid_constant__ ResArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = base;                // [kslabs][256 rows][64 B]
  const uint32_t smem_a = base + kRSmemB;      // [kRAStages][128 rows][64 B]
  const uint32_t bars = base + kRSmemB + kRSmemA;
  const uint32_t bar_afull = bars;                      // [kRAStages]
  const uint32_t bar_aempty = bars + 8 * kRAStages;     // [kRAStages]
  const uint32_t bar_bfull = bars + 16 * kRAStages;     // resident block landed
  const uint32_t bar_bempty = bar_bfull + 8;            // resident block no longer read by the tensor core
  const uint32_t bar_tfull = bar_bempty + 8;            // [2]
  const uint32_t bar_tempty = bar_tfull + 16;           // [2]
  const uint32_t tmem_slot = bar_tempty + 16;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous, equal share of the flattened tile list
  const int t_begin = (int)((long long)args.total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)args.total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRAStages; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    mbar_init(bar_bfull, 1);
    mbar_init(bar_bempty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, kREpiWarps);  // all 8 epilogue warps drain every tile
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kREpiWarps && lane == 0) {
    prefetch_tmap(&maps.src);
    for (int l = 0; l < args.levels; ++l) prefetch_tmap(&maps.tgt[l]);
  }
  if (warp == kREpiWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  PackedTile* table = reinterpret_cast<PackedTile*>(gen_base + kRSmemB + kRSmemA + 256);
  for (int i = threadIdx.x; i < t_end - t_begin; i += blockDim.x) {
    const TileInfo ti = decode_tile(args, t_begin + i);
    table[i].blk = (uint32_t)ti.blk;
    table[i].lvl_ty_tx = ((uint32_t)ti.level << 28) | ((uint32_t)ti.ty << 14) | (uint32_t)ti.tx;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  if (warp == kREpiWarps) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, bphase = 0;
      int cur_blk = -1;
      for (int t = t_begin; t < t_end; ++t) {
        const TileInfo ti = unpack_tile(args, table[t - t_begin]);
        if (ti.blk != cur_blk) {
          // (re)load the resident block once every tensor-core read of the previous one has retired
          mbar_wait(bar_bempty, bphase ^ 1);
          mbar_expect_tx(bar_bfull, (uint32_t)args.kslabs * kRBSlab);
          for (int k = 0; k < args.kslabs; ++k)
            tma_load_3d(smem_b + k * kRBSlab, &maps.src, bar_bfull, k * args.slab_elems, ti.mt * kRN, ti.b);
          bphase ^= 1;
          cur_blk = ti.blk;
        }
        for (int k = 0; k < args.kslabs; ++k) {
          mbar_wait(bar_aempty + 8 * stage, phase ^ 1);
          if (args.debug & 8) {
            mbar_arrive(bar_afull + 8 * stage);
          } else {
            mbar_expect_tx(bar_afull + 8 * stage, kRAStage);
            tma_load_4d(smem_a + stage * kRAStage, &maps.tgt[ti.level], bar_afull + 8 * stage, k * args.slab_elems,
                        ti.tx << args.pxs[ti.level], ti.ty << (7 - args.pxs[ti.level]), args.tgt_shared ? 0 : ti.b);
          }
          if (++stage == kRAStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kREpiWarps + 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_fmt((uint32_t)args.fmt, kRM, kRN);
      uint32_t stage = 0, phase = 0, bphase = 0;
      int cur_blk = -1;
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const TileInfo ti = unpack_tile(args, table[t - t_begin]);
        if (ti.blk != cur_blk) {
          mbar_wait(bar_bfull, bphase);
          bphase ^= 1;
          cur_blk = ti.blk;
        }
        const uint32_t ab = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * ab, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + ab * kRN;
        for (int k = 0; k < args.kslabs; ++k) {
          mbar_wait(bar_afull + 8 * stage, phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_a + stage * kRAStage);
          const uint64_t bdesc = make_smem_desc(smem_b + k * kRBSlab);
#pragma unroll
          for (int j = 0; j < kMmaPerSlab; ++j)
            if (!(args.debug & 4)) tc_mma<true>(tmem_d, adesc + 2 * j, bdesc + 2 * j, idesc, (k > 0 || j > 0) ? 1u : 0u);
          tc_commit(bar_aempty + 8 * stage);
          if (++stage == kRAStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(bar_tfull + 8 * ab);
        // last tile that reads this resident block: release it to the producer
        bool last_of_block = (t + 1 == t_end);
        if (!last_of_block) last_of_block = (int)table[t + 1 - t_begin].blk != cur_blk;
        if (last_of_block) tc_commit(bar_bempty);
      }
    }
  } else {
    // ===================================================================== epilogue: 16 warps on EVERY tile
    // warp (cq, q): TMEM lanes 32q..32q+31 (32 target pixels of the patch), accumulator columns
    // [64*cq, 64*cq+64) = 64 source pixels.  A warp retires one store every ~50-75 cycles (memory latency), so
    // the store rate scales with the number of warps; all warps drain the same tile so that one accumulator is
    // always being refilled by the tensor core while the other one drains.
    const int cq = warp >> 2, q = warp & 3;
    constexpr int kColsPerWarp = kRN / 4;
    const bool use_div = args.use_div != 0;
    const float divisor = args.divisor;
    // undo the operand scales (exact powers of two) and, when exact, fold in 1/sqrt(C)
    const float mul = __ldg(args.src_hdr + kHdrInvScale) * __ldg(args.tgt_hdr + kHdrInvScale) * (use_div ? 1.0f : args.rsqrt_c);
    const int mrow = 32 * q + lane;  // TMEM lane = row of the A tile = patch pixel (x fastest)
    const bool odd = (lane & 1) != 0;
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const TileInfo ti = unpack_tile(args, table[t - t_begin]);
      const int l = ti.level;
      const int pxs = args.pxs[l];
      const int py = mrow >> pxs, px = mrow & ((1 << pxs) - 1);
      const int m0 = ti.mt * kRN + cq * kColsPerWarp;
      const int mcount = min(kColsPerWarp, args.n1 - m0);  // may be <= 0 for a partial block
      const long long pitch = args.pitch[l];
      const int y = (ti.ty << (7 - pxs)) + py, x = (ti.tx << pxs) + px;
      const uint32_t ab = it & 1;

      mbar_wait(bar_tfull + 8 * ab, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * kRN + cq * kColsPerWarp;
      uint32_t ua[32], ub[32];
      if (!(args.debug & 16)) {
        tmem_ld32(taddr, ua);
        tmem_ld32(taddr + 32, ub);
        tmem_ld_wait(ua, ub);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) ua[i] = ub[i] = 0;
      }
      // this warp's share of the accumulator is in registers: hand the TMEM buffer back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * ab);

      if constexpr (OUT_HALF) {
        // Lanes (2i, 2i+1) hold x and x+1 of every column (source pixel).  Per column pair (j, j+1) they swap one
        // register: the even lane ends up with (x, x+1) of column j, the odd lane with (x-1, x) of column j+1, and each
        // stores ONE packed half2 into its column's map -- 16 lanes x 4 B = a 64-byte run per map and instruction.
        const int xe = x & ~1;
        const bool in = !(args.debug & 1) && y < args.lh[l] && xe < args.lw[l];
        const bool all_in = __all_sync(0xffffffffu, in);
        __half2* p = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(args.out[l]) +
                                                ((long long)ti.b * args.n1 + m0 + (odd ? 1 : 0)) * pitch + (long long)y * args.wp[l] + xe);
        const long long pstep = pitch;  // two columns further, in half2 units
        auto store_chunk = [&](const uint32_t (&u)[32], int chunk) {
          const int ncols = mcount - chunk * 32;
          if (ncols <= 0) return;  // warp-uniform
          const bool fast = ncols >= 32 && all_in && !use_div;
#pragma unroll
          for (int jj = 0; jj < 32; jj += 2) {
            if (!fast && jj >= ncols) break;  // warp-uniform
            float a = __uint_as_float(u[jj]) * mul, b = __uint_as_float(u[jj + 1]) * mul;
            if (use_div) {
              a = __fdiv_rn(a, divisor);
              b = __fdiv_rn(b, divisor);
            }
            const float recv = __shfl_xor_sync(0xffffffffu, odd ? a : b, 1);
            const float lo = odd ? recv : a, hi = odd ? b : recv;
            const __half2 v = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f), fminf(fmaxf(hi, -65504.f), 65504.f));
            if (fast || (in && jj + (odd ? 1 : 0) < ncols)) *p = v;
            p += pstep;
          }
        };
        store_chunk(ua, 0);
        store_chunk(ub, 1);
      } else {
        const bool in = !(args.debug & 1) && y < args.lh[l] && x < args.lw[l];
        const bool all_in = __all_sync(0xffffffffu, in);
        float* p = reinterpret_cast<float*>(args.out[l]) + ((long long)ti.b * args.n1 + m0) * pitch + (long long)y * args.wp[l] + x;
        auto store_chunk = [&](const uint32_t (&u)[32], int chunk) {
          const int ncols = mcount - chunk * 32;
          if (ncols >= 32 && all_in && !use_div) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              *p = __uint_as_float(u[jj]) * mul;
              p += pitch;
            }
          } else if (ncols > 0) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              if (jj < ncols) {  // warp-uniform
                float v = __uint_as_float(u[jj]) * mul;
                if (use_div) v = __fdiv_rn(v, divisor);
                if (in) *p = v;
                p += pitch;
              }
            }
          }
        };
        store_chunk(ua, 0);
        store_chunk(ub, 1);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kREpiWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ----------------------------------------------------------------------------- pre-pass
// One launch prepares every operand: fmap1 * scale -> 16-bit, and fmap2 avg-pooled to every level
// (fp32, recursive 2x2 pooling in ATen's order ((a+b)+c)+d)*0.25) -> 16-bit.  Block roles:
//   [0, nb_src)          fmap1 conversion, one float4 per thread
//   [nb_src, nb_l2)      one thread per 4x4 cell of fmap2 and channel quad: 16 independent loads, emits the
//                        cell's 16 level-0, 4 level-1 and 1 level-2 values
//   [nb_l2, ...)         one thread per level >= 3 output (recomputed from level 0; a few thousand threads)
struct PrepArgs {
  const float* fmap1;
  const float* fmap2;
  void* src16;
  void* tgt16[kRLevels];
  long long src_items;                  // float4 units of fmap1
  long long cell_items;                 // B * ceil(h2/2) * ceil(w2/2) * C4   (2x2 cells: levels 0 and 1)
  long long l2_items;                   // B * lh[2] * lw[2] * C4             (level 2 from 16 level-0 pixels)
  long long deep_begin[kRLevels + 1];   // prefix sums over levels >= 3 (scalar units), index l-3
  int nb_src, nb_cells, nb_l2;
  int B, n1, h2, w2, C4, levels;
  int lh[kRLevels], lw[kRLevels];
  float scale;
  int fmt;
};

__device__ __forceinline__ float4 pool4v(float4 a, float4 b, float4 d, float4 e) {
  float4 r;
  r.x = __fadd_rn(__fadd_rn(__fadd_rn(a.x, b.x), d.x), e.x) * 0.25f;
  r.y = __fadd_rn(__fadd_rn(__fadd_rn(a.y, b.y), d.y), e.y) * 0.25f;
  r.z = __fadd_rn(__fadd_rn(__fadd_rn(a.z, b.z), d.z), e.z) * 0.25f;
  r.w = __fadd_rn(__fadd_rn(__fadd_rn(a.w, b.w), d.w), e.w) * 0.25f;
  return r;
}

template <int L>
__device__ __forceinline__ float pooled_scalar(const float* __restrict__ f, int W, int C, int y, int x, int c) {
  // level-L value at (y, x), channel c: recursive 2x2 pooling of the 2^L x 2^L block of level 0
  if constexpr (L == 0) {
    return __ldg(f + ((long long)y * W + x) * C + c);
  } else {
    const float a = pooled_scalar<L - 1>(f, W, C, 2 * y, 2 * x, c), b = pooled_scalar<L - 1>(f, W, C, 2 * y, 2 * x + 1, c);
    const float d = pooled_scalar<L - 1>(f, W, C, 2 * y + 1, 2 * x, c), e = pooled_scalar<L - 1>(f, W, C, 2 * y + 1, 2 * x + 1, c);
    return __fadd_rn(__fadd_rn(__fadd_rn(a, b), d), e) * 0.25f;
  }
}

__device__ __forceinline__ unsigned short pack16_scalar(float v, int fmt) {
  if (fmt == 0) {
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    return *reinterpret_cast<const unsigned short*>(&h);
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<const unsigned short*>(&h);
}

__device__ __forceinline__ uint2 pack16(float4 v, int fmt) {
  uint2 o;
  if (fmt == 0) {
    // fp16, saturating (features beyond +-65504 would otherwise become inf)
    const float m = 65504.f;
    const __half2 p0 = __floats2half2_rn(fminf(fmaxf(v.x, -m), m), fminf(fmaxf(v.y, -m), m));
    const __half2 p1 = __floats2half2_rn(fminf(fmaxf(v.z, -m), m), fminf(fmaxf(v.w, -m), m));
    o.x = *reinterpret_cast<const uint32_t*>(&p0);
    o.y = *reinterpret_cast<const uint32_t*>(&p1);
  } else {
    const __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
    o.x = *reinterpret_cast<const uint32_t*>(&p0);
    o.y = *reinterpret_cast<const uint32_t*>(&p1);
  }
  return o;
}

__global__ void __launch_bounds__(256, 3) corr_prep16_kernel(const __grid_constant__ PrepArgs a) {
  const int C4 = a.C4;
  int blk = blockIdx.x;
  if (blk < a.nb_src) {
    const long long i = (long long)blk * 256 + threadIdx.x;
    if (i < a.src_items) {
      float4 v = __ldg(reinterpret_cast<const float4*>(a.fmap1) + i);
      v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;  // power of two: exact
      reinterpret_cast<uint2*>(a.src16)[i] = pack16(v, a.fmt);
    }
    return;
  }
  blk -= a.nb_src;
  if (blk < a.nb_cells) {
    // one thread per 2x2 cell of fmap2 and channel quad: 4 level-0 values and (if complete) 1 level-1 value
    const long long i = (long long)blk * 256 + threadIdx.x;
    if (i >= a.cell_items) return;
    const int cw = (a.w2 + 1) >> 1, ch = (a.h2 + 1) >> 1;
    const int c = (int)(i % C4);
    long long t = i / C4;
    const int cx = (int)(t % cw);
    t /= cw;
    const int cy = (int)(t % ch);
    const int b = (int)(t / ch);
    const float4* f = reinterpret_cast<const float4*>(a.fmap2) + (long long)b * a.h2 * a.w2 * C4;
    uint2* t0 = reinterpret_cast<uint2*>(a.tgt16[0]) + (long long)b * a.h2 * a.w2 * C4;
    float4 v[2][2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = 2 * cy + dy, x = 2 * cx + dx;
        v[dy][dx] = (y < a.h2 && x < a.w2) ? __ldg(f + ((long long)y * a.w2 + x) * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = 2 * cy + dy, x = 2 * cx + dx;
        if (y < a.h2 && x < a.w2) t0[((long long)y * a.w2 + x) * C4 + c] = pack16(v[dy][dx], a.fmt);
      }
    if (a.levels > 1 && cy < a.lh[1] && cx < a.lw[1]) {
      uint2* t1 = reinterpret_cast<uint2*>(a.tgt16[1]) + (long long)b * a.lh[1] * a.lw[1] * C4;
      t1[((long long)cy * a.lw[1] + cx) * C4 + c] = pack16(pool4v(v[0][0], v[0][1], v[1][0], v[1][1]), a.fmt);
    }
    return;
  }
  blk -= a.nb_cells;
  if (blk < a.nb_l2) {
    // level 2: one thread per output and channel quad, 16 independent level-0 loads
    const long long i = (long long)blk * 256 + threadIdx.x;
    if (i >= a.l2_items) return;
    const int c = (int)(i % C4);
    long long t = i / C4;
    const int x = (int)(t % a.lw[2]);
    t /= a.lw[2];
    const int y = (int)(t % a.lh[2]);
    const int b = (int)(t / a.lh[2]);
    const float4* f = reinterpret_cast<const float4*>(a.fmap2) + (long long)b * a.h2 * a.w2 * C4;
    float4 l1[2][2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const float4* q = f + ((long long)(4 * y + 2 * dy) * a.w2 + (4 * x + 2 * dx)) * C4 + c;
        l1[dy][dx] = pool4v(__ldg(q), __ldg(q + C4), __ldg(q + (long long)a.w2 * C4), __ldg(q + (long long)a.w2 * C4 + C4));
      }
    reinterpret_cast<uint2*>(a.tgt16[2])[i] = pack16(pool4v(l1[0][0], l1[0][1], l1[1][0], l1[1][1]), a.fmt);
    return;
  }
  blk -= a.nb_l2;
  {
    // levels >= 3: one thread per scalar output, recomputed from level 0 (a few thousand threads)
    const long long i = (long long)blk * 256 + threadIdx.x;
    if (i >= a.deep_begin[a.levels > 3 ? a.levels - 3 : 0]) return;
    int l = 3;
    while (l + 1 < a.levels && i >= a.deep_begin[l - 2]) ++l;
    const long long r = i - a.deep_begin[l - 3];
    const int C = 4 * C4;
    const int c = (int)(r % C);
    long long t = r / C;
    const int x = (int)(t % a.lw[l]);
    t /= a.lw[l];
    const int y = (int)(t % a.lh[l]);
    const int b = (int)(t / a.lh[l]);
    const float* f = a.fmap2 + (long long)b * a.h2 * a.w2 * C;
    // pool the four quadrants one after the other (bounds the number of loads in flight / live registers)
    float qv[4];
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int yy = 2 * y + (k >> 1), xx = 2 * x + (k & 1);
      switch (l) {
        case 3: qv[k] = pooled_scalar<2>(f, a.w2, C, yy, xx, c); break;
        case 4: qv[k] = pooled_scalar<3>(f, a.w2, C, yy, xx, c); break;
        default: qv[k] = pooled_scalar<4>(f, a.w2, C, yy, xx, c); break;
      }
    }
    const float v = __fadd_rn(__fadd_rn(__fadd_rn(qv[0], qv[1]), qv[2]), qv[3]) * 0.25f;
    reinterpret_cast<unsigned short*>(a.tgt16[l])[r] = pack16_scalar(v, a.fmt);
  }
}

// ----------------------------------------------------------------------------- host
static int64_t al256(int64_t v) { return (v + 255) & ~(int64_t)255; }

int64_t corr_res_workspace_bytes(int B, int n1, int h2, int w2, int C, int levels) {
  int64_t bytes = al256((int64_t)B * n1 * C * 2);
  for (int l = 0; l < levels; ++l) bytes += al256((int64_t)B * (h2 >> l) * (w2 >> l) * C * 2);
  return bytes;
}

bool corr_res_supported(int C, int levels) { return C % 8 == 0 && C <= kRMaxSlabs * 32 && levels >= 1 && levels <= 6; }

// Pre-pass: 16-bit operand copies of fmap1 (scaled) and of every pooled level of fmap2 into `workspace`.
// part: 1 = fmap1 only, 2 = fmap2 levels only, 3 = both (a key frame's fmap2 operands can be reused by every
// pair that shares it).
int launch_corr_prepare_resident(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, int fmt,
                                 const sdof_pyramid_layout& lay, void* workspace, int64_t workspace_bytes, int part,
                                 cudaStream_t st) {
  const int levels = lay.levels;
  if (!corr_res_supported(C, levels)) return SDOF_ERR_UNSUPPORTED;
  if (B > 65535) return SDOF_ERR_UNSUPPORTED;
  const int64_t need = corr_res_workspace_bytes(B, n1, h2, w2, C, levels);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(SDOF_ERR_INVALID, "correlation workspace of %lld bytes required, got %lld", (long long)need,
                (long long)workspace_bytes);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return fail(SDOF_ERR_INVALID, "correlation workspace must be 256-byte aligned");
  const bool pow4 = (C & (C - 1)) == 0 && (__builtin_ctz(C) % 2 == 0);
  PrepArgs pa;
  memset(&pa, 0, sizeof(pa));
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  pa.fmap1 = fmap1;
  pa.fmap2 = fmap2;
  pa.src16 = w;
  w += al256((int64_t)B * n1 * C * 2);
  pa.B = B; pa.n1 = n1; pa.h2 = h2; pa.w2 = w2; pa.C4 = C / 4; pa.levels = levels;
  pa.scale = pow4 ? 1.0f / sqrtf((float)C) : 1.0f;
  pa.fmt = fmt;
  pa.src_items = (part & 1) ? (int64_t)B * n1 * (C / 4) : 0;
  pa.cell_items = (part & 2) ? (int64_t)B * ((h2 + 1) / 2) * ((w2 + 1) / 2) * (C / 4) : 0;
  pa.l2_items = ((part & 2) && levels > 2) ? (int64_t)B * lay.h[2] * lay.w[2] * (C / 4) : 0;
  pa.deep_begin[0] = 0;
  for (int l = 0; l < levels; ++l) {
    pa.lh[l] = lay.h[l];
    pa.lw[l] = lay.w[l];
    pa.tgt16[l] = w;
    w += al256((int64_t)B * lay.h[l] * lay.w[l] * C * 2);
    if (l >= 3) pa.deep_begin[l - 2] = pa.deep_begin[l - 3] + ((part & 2) ? (int64_t)B * lay.h[l] * lay.w[l] * C : 0);
  }
  const long long deep_items = levels > 3 ? pa.deep_begin[levels - 3] : 0;
  pa.nb_src = (int)ceil_div64(pa.src_items, 256);
  pa.nb_cells = (int)ceil_div64(pa.cell_items, 256);
  pa.nb_l2 = (int)ceil_div64(pa.l2_items, 256);
  const int nb_deep = (int)ceil_div64(deep_items, 256);
  const int nb = pa.nb_src + pa.nb_cells + pa.nb_l2 + nb_deep;
  if (nb == 0) return SDOF_OK;
  corr_prep16_kernel<<<nb, 256, 0, st>>>(pa);
  SDOF_LAUNCH_CHECK("corr_prep16_kernel");
  return SDOF_OK;
}

// Main kernel on operands prepared by launch_corr_prepare_resident.
int launch_corr_pyramid_prepared(int B, int n1, int h2, int w2, int C, int fmt, float* pyramid, const sdof_pyramid_layout& lay,
                                 void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const int levels = lay.levels;
  if (!corr_res_supported(C, levels)) return SDOF_ERR_UNSUPPORTED;
  if (B > 65535) return SDOF_ERR_UNSUPPORTED;
  const int64_t need = corr_res_workspace_bytes(B, n1, h2, w2, C, levels);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(SDOF_ERR_INVALID, "correlation workspace of %lld bytes required, got %lld", (long long)need,
                (long long)workspace_bytes);
  const bool pow4 = (C & (C - 1)) == 0 && (__builtin_ctz(C) % 2 == 0);
  ResArgs ra;
  memset(&ra, 0, sizeof(ra));
  ResMaps maps;
  memset(&maps, 0, sizeof(maps));
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  void* src16 = w;
  w += al256((int64_t)B * n1 * C * 2);
  void* tgt16[kRLevels] = {};
  int used_levels = 0;
  for (int l = 0; l < levels; ++l) {
    tgt16[l] = w;
    w += al256((int64_t)B * lay.h[l] * lay.w[l] * C * 2);
    if (lay.h[l] >= 1 && lay.w[l] >= 1) used_levels = l + 1;
  }
  if (used_levels == 0) return SDOF_OK;

  const CUtensorMapDataType dt = fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int slab_elems = kSlabBytes / 2;
  int rc;
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)n1, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)n1 * C * 2};
    cuuint32_t box[3] = {(cuuint32_t)slab_elems, (cuuint32_t)kRN, 1};
    if ((rc = encode_map(&maps.src, dt, 3, src16, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B, "fmap1 (16-bit)"))) return rc;
  }
  ra.B = B;
  ra.n1 = n1;
  ra.m_tiles = ceil_div(n1, kRN);
  ra.levels = used_levels;
  ra.slab_elems = slab_elems;
  ra.kslabs = ceil_div(C, slab_elems);
  ra.divisor = sqrtf((float)C);
  ra.use_div = !pow4;
  ra.fmt = fmt;
  ra.patch_x = 32;
  {
    const char* e = getenv("SDOF_RES_PATCHX");
    if (e && (atoi(e) == 16 || atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 8)) ra.patch_x = atoi(e);
    const char* d = getenv("SDOF_RES_DEBUG");
    ra.debug = d ? atoi(d) : 0;
  }
  ra.patch_y = kRM / ra.patch_x;
  ra.tile_begin_level[0] = 0;
  for (int l = 0; l < used_levels; ++l) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)lay.w[l], (cuuint64_t)lay.h[l], (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)lay.w[l] * C * 2, (cuuint64_t)lay.h[l] * lay.w[l] * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)slab_elems, (cuuint32_t)ra.patch_x, (cuuint32_t)ra.patch_y, 1};
    if ((rc = encode_map(&maps.tgt[l], dt, 4, tgt16[l], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B, "pooled fmap2 (16-bit)")))
      return rc;
    const int tyt = ceil_div(lay.h[l], ra.patch_y), txt = ceil_div(lay.w[l], ra.patch_x);
    ra.tx_tiles[l] = txt;
    ra.tile_begin_level[l + 1] = ra.tile_begin_level[l] + tyt * txt;
    ra.lh[l] = lay.h[l];
    ra.lw[l] = lay.w[l];
    ra.wp[l] = lay.wp[l];
    ra.pitch[l] = lay.pitch[l];
    ra.out[l] = pyramid + lay.offset[l];
  }
  const long long total = (long long)B * ra.m_tiles * ra.tile_begin_level[used_levels];
  if (total > 0x7fffffff) return SDOF_ERR_UNSUPPORTED;
  ra.total_tiles = (int)total;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  if (ceil_div64(total, grid) + 1 > kRMaxTilesPerCta) return SDOF_ERR_UNSUPPORTED;  // tile table would not fit
  SDOF_CUDA(cudaFuncSetAttribute(corr_pyramid_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRSmemTotal));
  corr_pyramid_resident_kernel<<<grid, kRThreads, kRSmemTotal, st>>>(maps, ra);
  SDOF_LAUNCH_CHECK("corr_pyramid_resident_kernel");
  return SDOF_OK;
}


int launch_corr_pyramid_resident(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, int fmt,
                                 float* pyramid, const sdof_pyramid_layout& lay, void* workspace, int64_t workspace_bytes,
                                 cudaStream_t st) {
  int rc = launch_corr_prepare_resident(fmap1, fmap2, B, n1, h2, w2, C, fmt, lay, workspace, workspace_bytes, 3, st);
  if (rc) return rc;
  return launch_corr_pyramid_prepared(B, n1, h2, w2, C, fmt, pyramid, lay, workspace, workspace_bytes, st);
}

}  // namespace sdof
