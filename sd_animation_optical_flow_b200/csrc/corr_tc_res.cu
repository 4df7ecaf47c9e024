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
#include <vector>

#include "tc_ptx.cuh"

namespace sdof {

// target patch = 128 pixels; shape chosen at launch (patch_x in {32, 16}; SDOF_RES_PATCHX overrides for experiments)
constexpr int kRM = 128;                   // target pixels per tile (TMEM lanes)
constexpr int kRN = 256;                   // resident source pixels (accumulator columns)
// K-slab of THIS kernel: 128-byte rows (64 fp16/bf16 channels, SWIZZLE_128B), four UMMA instructions per slab.  Round 1 used
// 64-byte slabs: eight mbarrier round trips per tile made the single MMA-issuing thread (and the producer thread) the
// limiter once the fp16-stored epilogue stopped being one (ablation: 1.9 us per tile with loads, MMA and stores skipped).
constexpr int kRSlabBytes = 128;
constexpr int kRMmaPerSlab = kRSlabBytes / 32;
constexpr int kRMaxSlabs = 4;              // C <= 4 * 64 = 256 channels resident
constexpr int kRAStages = 5;               // ring of 16 KB target-patch slabs (a tile is 4): what fits beside the 128 KB resident block
constexpr int kRAStage = kRM * kRSlabBytes;  // 16384
constexpr int kRBSlab = kRN * kRSlabBytes;   // 32768
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
  float* out_factor;                   // fp16 pyramid header (float[0]): stored * factor = correlation
  int total_tiles;
  float divisor;                       // sqrt(C)
  float rsqrt_c;                       // 1/sqrt(C) when that is a power of two (use_div == 0)
  int use_div;
  int fmt;  // 0 = fp16, 1 = bf16
  int tgt_shared;                      // the target operand has batch 1 and serves every pair of the call
  int debug;             // SDOF_RES_DEBUG: 1 skip stores, 4 skip MMA, 8 skip A loads, 16 skip TMEM loads (experiments only)
  long long* trace;      // built with -DSDOF_RES_TRACE and run with SDOF_RES_TRACE=<file>: clock64 stamps [cta][tile of the cta < kRTraceTiles][kRTraceSlots] (experiments only)
};
constexpr int kRTraceTiles = 64, kRTraceSlots = 16;
// slots: 0 MMA warp owns the accumulator | 1 first slab landed | 2 tile's MMAs issued + committed | 3 epilogue warp 0 saw the
// accumulator full | 4 all 64 columns in registers, buffer handed back | 5 stores of columns 0-31 issued | 7 of columns 32-63 |
// 8 producer issued the tile's last slab load | 9 producer started (re)loading the resident block | 10 MMA thread saw it landed
#ifndef SDOF_RES_TRACE   // compiled out of the product build: the stamps cost ~1.8 us of a 32 us kernel even when switched off
#define RES_TRACE(it_, slot_) do { } while (0)
#else
#define RES_TRACE(it_, slot_)                                                                                          \
  do {                                                                                                                 \
    if (args.trace != nullptr && (it_) < kRTraceTiles)                                                                 \
      args.trace[((long long)blockIdx.x * kRTraceTiles + (it_)) * kRTraceSlots + (slot_)] = clock64();                 \
  } while (0)
#endif

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
  const uint32_t smem_b = base;                // [kslabs][256 rows][128 B]
  const uint32_t smem_a = base + kRSmemB;      // [kRAStages][128 rows][128 B]
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
  // everything above (barrier init, TMEM allocation, tile table, descriptor prefetch) touched nothing the operand pre-pass
  // writes: with a programmatic dependent launch it overlapped that kernel's tail
  pdl_wait();
  pdl_trigger();

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
          RES_TRACE(t - t_begin, 9);
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
        RES_TRACE(t - t_begin, 8);
      }
    }
  } else if (warp == kREpiWarps + 1) {
    // ===================================================================== MMA issuer
    // The WHOLE warp walks the loop and one elected lane issues: every operand of the tensor-core instructions is then
    // warp-uniform and lives in uniform registers.  (Round 2 finding, tools/corr_trace.py + tools/microbench/mma_rate.cu: with
    // the loop under `if (lane == 0)` the compiler moves each descriptor through R2UR and wraps each tcgen05 instruction in
    // an ELECT loop -- ~300 cycles of single-lane work per 128-byte slab, more than the tensor pipe's short queue hides:
    // 176 instead of 128 cycles per MMA and ~900 idle cycles per tile.)  No shared-memory reads on this path either: the
    // resident-block boundaries follow from the tile index.
    const uint32_t idesc = make_idesc_fmt((uint32_t)args.fmt, kRM, kRN);
    const int per_block = args.tile_begin_level[args.levels];
    const bool skip_mma = (args.debug & 4) != 0;
    const int kslabs = args.kslabs;
    uint32_t stage = 0, phase = 0, bphase = 0;
    int r = t_begin % per_block;      // position of the tile inside its source block
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      // (one lane polls: thirty-two polling lanes slow every other mbarrier operation of the CTA down)
      if (it == 0 || r == 0) {
        if (lane == 0) mbar_wait(bar_bfull, bphase);
        __syncwarp();
        if (lane == 0) RES_TRACE(it, 10);
        bphase ^= 1;
      }
      const uint32_t ab = it & 1;
      if (lane == 0) mbar_wait(bar_tempty + 8 * ab, ((it >> 1) & 1) ^ 1);
      __syncwarp();
      tc_fence_after();
      if (lane == 0) RES_TRACE(it, 0);
      const uint32_t tmem_d = tmem_base + ab * kRN;
      for (int k = 0; k < kslabs; ++k) {
        if (lane == 0) mbar_wait(bar_afull + 8 * stage, phase);
        __syncwarp();
        tc_fence_after();
        if (k == 0 && lane == 0) RES_TRACE(it, 1);
        const uint64_t adesc = make_smem_desc_sw128(smem_a + stage * kRAStage);
        const uint64_t bdesc = make_smem_desc_sw128(smem_b + k * kRBSlab);
        if (elect_one()) {
          if (!skip_mma) {
#pragma unroll
            for (int j = 0; j < kRMmaPerSlab; ++j) tc_mma<true>(tmem_d, adesc + 2 * j, bdesc + 2 * j, idesc, (k > 0 || j > 0) ? 1u : 0u);
          }
          tc_commit(bar_aempty + 8 * stage);
        }
        __syncwarp();
        if (++stage == kRAStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      // last tile that reads this resident block: release it to the producer
      const bool last_of_block = (t + 1 == t_end) || (r + 1 == per_block);
      if (elect_one()) {
        tc_commit(bar_tfull + 8 * ab);
        if (last_of_block) tc_commit(bar_bempty);
      }
      __syncwarp();
      if (lane == 0) {
        RES_TRACE(it, 2);
        if (args.trace != nullptr && it < kRTraceTiles) {   // wall clock beside the cycle counter: the SM clock of this run
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          args.trace[((long long)blockIdx.x * kRTraceTiles + it) * kRTraceSlots + 11] = (long long)gt;
        }
      }
      if (++r == per_block) r = 0;
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
    if (OUT_HALF && blockIdx.x == 0 && threadIdx.x == 0) {
      // fp16 pyramid: stored value = raw accumulator; true correlation = stored * factor (read by the lookup kernel)
      const float f = __ldg(args.src_hdr + kHdrInvScale) * __ldg(args.tgt_hdr + kHdrInvScale);
      *args.out_factor = use_div ? __fdiv_rn(f, divisor) : f * args.rsqrt_c;
    }
    // Drain, release, store: TMEM -> registers is fast (470-890 B/clk per SM, tools/microbench/tmem_ld.cu: ~150 ns for this
    // warp's 64 columns), so both halves are loaded at once and the accumulator goes back to the tensor core before the first
    // store.  (Measured and dropped in round 2: software-pipelining the two halves against the stores, with the hand-over in
    // between -- 33.5 vs 31.5 us; the tile time is the memory system accepting this store pattern, see profiles/README.md.)
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + cq * kColsPerWarp;
    uint32_t ua[32], ub[32];
    const bool skip_ld = (args.debug & 16) != 0;
    auto issue_ld = [&](uint32_t (&u)[32], uint32_t addr) {
      if (!skip_ld) {
        tmem_ld32(addr, u);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) u[i] = 0;
      }
    };
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
      // one lane per warp polls the barrier (512 polling threads slow every other mbarrier operation of the CTA down);
      // tcgen05.fence::after_thread_sync orders the TMEM loads after the warp-level hand-over
      if (lane == 0) mbar_wait(bar_tfull + 8 * ab, (it >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      if (threadIdx.x == 0) RES_TRACE(it, 3);
      issue_ld(ua, taddr0 + ab * kRN);
      issue_ld(ub, taddr0 + ab * kRN + 32);
      tmem_ld_wait(ua, ub);
      if (threadIdx.x == 0) RES_TRACE(it, 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * ab);

      if constexpr (OUT_HALF) {
        // The operands were scaled so that |accumulator| < 2^15 (corr_prep16_kernel), so the raw fp32 accumulator converts
        // to fp16 without overflow and WITHOUT a multiply; the pyramid header carries the factor the lookup applies.
        // Lanes (2i, 2i+1) hold x and x+1 of every column (source pixel).  Per column pair (j, j+1) they swap one
        // register: the even lane ends up with (x, x+1) of column j, the odd lane with (x-1, x) of column j+1, and each
        // stores ONE packed half2 into its column's map -- 16 lanes x 4 B = a 64-byte run per map and instruction.
        const int xe = x & ~1;
        const bool in = y < args.lh[l] && xe < args.lw[l];
        const bool all_in = __all_sync(0xffffffffu, in);
        const bool do_store = !(args.debug & 1);
        __half2* p = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(args.out[l]) +
                                                ((long long)ti.b * args.n1 + m0 + (odd ? 1 : 0)) * pitch + (long long)y * args.wp[l] + xe);
        // Instruction diet (the epilogue is ALU-pipe bound: 2 issue cycles per ALU instruction and quarter-SM): each lane
        // first packs ITS pixel's (column j, column j+1) into one half2, the pair of lanes swaps the packed words, and one
        // PRMT picks (x, x+1) of the lane's column; the address is base + constant * pitch in one IMAD.WIDE.  5
        // instructions per 4-byte store where the select-shuffle-select-convert-add64 sequence took 9.
        const uint32_t pstep = (uint32_t)pitch;                 // two columns further, in half2 units (pitch < 2^31 halves)
        const uint32_t sel = odd ? 0x3276u : 0x5410u;           // odd: (partner's hi, my hi) = column j+1; even: (my lo, partner's lo) = column j
        auto store_chunk = [&](const uint32_t (&u)[32], int chunk) {
          const int ncols = mcount - chunk * 32;
          __half2* pc = p + (size_t)(chunk * 16) * pstep;
          if (ncols >= 32 && all_in) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __half2 h = __floats2half2_rn(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1]));
              const uint32_t mine = *reinterpret_cast<const uint32_t*>(&h);
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, mine, 1);
              const uint32_t v = __byte_perm(mine, recv, sel);
              if (do_store) *reinterpret_cast<uint32_t*>(pc + (size_t)i * pstep) = v;
            }
          } else if (ncols > 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (2 * i < ncols) {  // warp-uniform
                const __half2 h = __floats2half2_rn(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1]));
                const uint32_t mine = *reinterpret_cast<const uint32_t*>(&h);
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, mine, 1);
                const uint32_t v = __byte_perm(mine, recv, sel);
                if (do_store && in && 2 * i + (odd ? 1 : 0) < ncols) *reinterpret_cast<uint32_t*>(pc + (size_t)i * pstep) = v;
              }
            }
          }
        };
        store_chunk(ua, 0);
        if (threadIdx.x == 0) RES_TRACE(it, 5);
        store_chunk(ub, 1);
        if (threadIdx.x == 0) RES_TRACE(it, 7);
      } else {
        const bool in = y < args.lh[l] && x < args.lw[l];
        const bool all_in = __all_sync(0xffffffffu, in);
        const bool do_store = !(args.debug & 1);
        float* p = reinterpret_cast<float*>(args.out[l]) + ((long long)ti.b * args.n1 + m0) * pitch + (long long)y * args.wp[l] + x;
        const uint32_t pstep = (uint32_t)pitch;
        auto store_chunk = [&](const uint32_t (&u)[32], int chunk) {
          const int ncols = mcount - chunk * 32;
          float* pc = p + (size_t)(chunk * 32) * pstep;
          if (ncols >= 32 && all_in && !use_div) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (do_store) pc[(size_t)jj * pstep] = __uint_as_float(u[jj]) * mul;
          } else if (ncols > 0) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              if (jj < ncols) {  // warp-uniform
                float v = __uint_as_float(u[jj]) * mul;
                if (use_div) v = __fdiv_rn(v, divisor);
                if (in && do_store) pc[(size_t)jj * pstep] = v;
              }
            }
          }
        };
        store_chunk(ua, 0);
        if (threadIdx.x == 0) RES_TRACE(it, 5);
        store_chunk(ub, 1);
        if (threadIdx.x == 0) RES_TRACE(it, 7);
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
  float* src_hdr;                       // header of the source operand buffer (partials in, scale out); NULL = part not requested
  float* tgt_hdr;
  void* src16;
  void* tgt16[kRLevels];
  long long src_items;                  // float4 units of fmap1
  long long cell_items;                 // B * ceil(h2/2) * ceil(w2/2) * C4   (2x2 cells: levels 0 and 1)
  long long l2_items;                   // B * lh[2] * lw[2] * C4             (level 2 from 16 level-0 pixels)
  long long deep_begin[kRLevels + 1];   // prefix sums over levels >= 3 (scalar units), index l-3
  int nb_src, nb_cells, nb_l2;
  int B, n1, h2, w2, C4, levels;
  int lh[kRLevels], lw[kRLevels];
  int fmt;
  int src_exp, tgt_exp;                 // see scale_from_partials
};

// Per-tensor abs-max, pass 1: kAmaxBlocks partial maxima per tensor into the operand header (no atomics, no reset needed;
// NaNs are ignored by fmaxf, an infinity switches the scaling off).  blockIdx.y selects the tensor.
__global__ void __launch_bounds__(256) corr_absmax_kernel(const float4* __restrict__ a, long long na4, float* __restrict__ pa,
                                                          const float4* __restrict__ b, long long nb4, float* __restrict__ pb) {
  pdl_wait();
  pdl_trigger();
  const float4* x = blockIdx.y == 0 ? a : b;
  const long long n4 = blockIdx.y == 0 ? na4 : nb4;
  float* part = blockIdx.y == 0 ? pa : pb;
  if (x == nullptr || part == nullptr) return;
  float m = 0.f;
  const long long stride = (long long)gridDim.x * 256;
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {   // four independent loads in flight
    const float4 v0 = __ldg(x + i), v1 = __ldg(x + i + stride), v2 = __ldg(x + i + 2 * stride), v3 = __ldg(x + i + 3 * stride);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w))));
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w))));
  }
  for (; i < n4; i += stride) {
    const float4 v = __ldg(x + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    part[blockIdx.x] = m;
  }
}

// Pass 2 (inside the conversion kernel): the power-of-two scale of a tensor from its abs-max partials; 1 for an all-zero /
// non-finite tensor.
// The scales of the two tensors are chosen so that C * max|f1 s1| * max|f2 s2| < 2^15: the fp32 accumulator of ANY pair of
// pixels then fits fp16, so the fp16-stored pyramid is written without a rescale or a clamp (target_exp: amax * s lands in
// [2^(target_exp-1), 2^target_exp); for C = 256: 2^4 for fmap1, 2^3 for fmap2).  fp16 keeps 11 significant bits down to
// 2^-14, i.e. for every element larger than 2^-17 of the tensor's maximum.
__device__ __forceinline__ float scale_from_partials(const float* __restrict__ part, int lane, int target_exp) {
  float m = 0.f;
  for (int i = lane; i < kAmaxBlocks; i += 32) m = fmaxf(m, part[i]);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (!(m > 0.f) || !(m < 3.0e38f)) return 1.0f;
  int e;
  frexpf(m, &e);                      // m = f * 2^e, f in [0.5, 1)
  e = target_exp - e;
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  return ldexpf(1.0f, e);
}

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

__device__ __forceinline__ float4 scaled4(float4 v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }

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

constexpr int kPrepSrcPerThread = 4;   // float4 items per thread in the fmap1 section (independent loads in flight)

__global__ void __launch_bounds__(256, 4) corr_prep16_kernel(const __grid_constant__ PrepArgs a) {
  pdl_wait();
  pdl_trigger();
  const int C4 = a.C4;
  int blk = blockIdx.x;
  // every CTA derives its tensor's scale from the abs-max partials (128 floats from L2); the first CTA of a tensor's
  // section publishes it for the correlation kernel's epilogue
  __shared__ float s_scale;
  const bool is_src = blk < a.nb_src;
  if (threadIdx.x < 32) {
    float* hdr = is_src ? a.src_hdr : a.tgt_hdr;
    const float sc = scale_from_partials(hdr, threadIdx.x, is_src ? a.src_exp : a.tgt_exp);
    if (threadIdx.x == 0) {
      s_scale = sc;
      if (blk == 0 || blk == a.nb_src) {
        hdr[kHdrInvScale] = 1.0f / sc;   // exact: power of two
        hdr[kHdrScale] = sc;
      }
    }
  }
  __syncthreads();
  const float scale = s_scale;
  if (is_src) {
    const long long i0 = (long long)blk * (256 * kPrepSrcPerThread) + threadIdx.x;
    float4 v[kPrepSrcPerThread];
#pragma unroll
    for (int k = 0; k < kPrepSrcPerThread; ++k) {
      const long long i = i0 + k * 256;
      if (i < a.src_items) v[k] = __ldg(reinterpret_cast<const float4*>(a.fmap1) + i);
    }
#pragma unroll
    for (int k = 0; k < kPrepSrcPerThread; ++k) {
      const long long i = i0 + k * 256;
      if (i < a.src_items) {
        v[k].x *= scale; v[k].y *= scale; v[k].z *= scale; v[k].w *= scale;  // power of two: exact
        reinterpret_cast<uint2*>(a.src16)[i] = pack16(v[k], a.fmt);
      }
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
        if (y < a.h2 && x < a.w2) t0[((long long)y * a.w2 + x) * C4 + c] = pack16(scaled4(v[dy][dx], scale), a.fmt);
      }
    if (a.levels > 1 && cy < a.lh[1] && cx < a.lw[1]) {
      uint2* t1 = reinterpret_cast<uint2*>(a.tgt16[1]) + (long long)b * a.lh[1] * a.lw[1] * C4;
      t1[((long long)cy * a.lw[1] + cx) * C4 + c] = pack16(scaled4(pool4v(v[0][0], v[0][1], v[1][0], v[1][1]), scale), a.fmt);
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
    reinterpret_cast<uint2*>(a.tgt16[2])[i] = pack16(scaled4(pool4v(l1[0][0], l1[0][1], l1[1][0], l1[1][1]), scale), a.fmt);
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
    reinterpret_cast<unsigned short*>(a.tgt16[l])[r] = pack16_scalar(v * scale, a.fmt);
  }
}

// ----------------------------------------------------------------------------- host
static int64_t al256(int64_t v) { return (v + 255) & ~(int64_t)255; }

int64_t corr_res_src_bytes(int B, int n1, int C) { return kOpHdrBytes + al256((int64_t)B * n1 * C * 2); }
int64_t corr_res_tgt_bytes(int B2, int h2, int w2, int C, int levels) {
  int64_t bytes = kOpHdrBytes;
  for (int l = 0; l < levels; ++l) bytes += al256((int64_t)B2 * (h2 >> l) * (w2 >> l) * C * 2);
  return bytes;
}
int64_t corr_res_workspace_bytes(int B, int n1, int h2, int w2, int C, int levels) {
  return corr_res_src_bytes(B, n1, C) + corr_res_tgt_bytes(B, h2, w2, C, levels);
}

bool corr_res_supported(int C, int levels) { return C % 8 == 0 && C <= kRMaxSlabs * (kRSlabBytes / 2) && levels >= 1 && levels <= 6; }

static int check_ops_ptr(const void* p, const char* what) {
  if (p == nullptr) return fail(SDOF_ERR_INVALID, "%s operand buffer is NULL", what);
  if ((reinterpret_cast<uintptr_t>(p) & 255) != 0) return fail(SDOF_ERR_INVALID, "%s operand buffer must be 256-byte aligned", what);
  return SDOF_OK;
}

// Pre-pass: abs-max + auto-ranged 16-bit copies.  fmap1 != NULL: fmap1 [B, n1, C] -> src_ops; fmap2 != NULL: every
// avg-pooled level of fmap2 [B2, h2, w2, C] -> tgt_ops (a key frame's target operands are built once and reused by every
// pair that shares it).  Two launches (partials, conversion) for whichever parts are requested.
int launch_corr_prepare_parts(const float* fmap1, int B, int n1, void* src_ops, const float* fmap2, int B2, int h2, int w2,
                              void* tgt_ops, int C, int levels, int fmt, cudaStream_t st) {
  if (!corr_res_supported(C, levels)) return SDOF_ERR_UNSUPPORTED;
  if (B > 65535 || B2 > 65535) return SDOF_ERR_UNSUPPORTED;
  int rc;
  if (fmap1 && (rc = check_ops_ptr(src_ops, "source"))) return rc;
  if (fmap2 && (rc = check_ops_ptr(tgt_ops, "target"))) return rc;
  if (!fmap1 && !fmap2) return SDOF_OK;
  PrepArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.fmap1 = fmap1;
  pa.fmap2 = fmap2;
  pa.B = fmap2 ? B2 : B; pa.n1 = n1; pa.h2 = h2; pa.w2 = w2; pa.C4 = C / 4; pa.levels = levels;
  pa.fmt = fmt;
  {
    int L = 0;
    while ((1 << L) < C) ++L;
    pa.src_exp = (15 - L + 1) / 2;   // ceil
    pa.tgt_exp = (15 - L) / 2;       // floor: src_exp + tgt_exp + L == 15
  }
  if (fmap1) {
    pa.src_hdr = reinterpret_cast<float*>(src_ops);
    pa.src16 = reinterpret_cast<uint8_t*>(src_ops) + kOpHdrBytes;
    pa.src_items = (int64_t)B * n1 * (C / 4);
  }
  pa.deep_begin[0] = 0;
  if (fmap2) {
    pa.tgt_hdr = reinterpret_cast<float*>(tgt_ops);
    uint8_t* w = reinterpret_cast<uint8_t*>(tgt_ops) + kOpHdrBytes;
    pa.cell_items = (int64_t)B2 * ((h2 + 1) / 2) * ((w2 + 1) / 2) * (C / 4);
    for (int l = 0; l < levels; ++l) {
      pa.lh[l] = h2 >> l;
      pa.lw[l] = w2 >> l;
      pa.tgt16[l] = w;
      w += al256((int64_t)B2 * pa.lh[l] * pa.lw[l] * C * 2);
      if (l >= 3) pa.deep_begin[l - 2] = pa.deep_begin[l - 3] + (int64_t)B2 * pa.lh[l] * pa.lw[l] * C;
    }
    pa.l2_items = levels > 2 ? (int64_t)B2 * pa.lh[2] * pa.lw[2] * (C / 4) : 0;
  }
  const long long deep_items = (fmap2 && levels > 3) ? pa.deep_begin[levels - 3] : 0;
  pa.nb_src = (int)ceil_div64(pa.src_items, 256 * kPrepSrcPerThread);
  pa.nb_cells = (int)ceil_div64(pa.cell_items, 256);
  pa.nb_l2 = (int)ceil_div64(pa.l2_items, 256);
  const int nb_deep = (int)ceil_div64(deep_items, 256);
  const int nb = pa.nb_src + pa.nb_cells + pa.nb_l2 + nb_deep;
  if (nb == 0) return SDOF_OK;
  SDOF_CUDA(launch_pdl(corr_absmax_kernel, dim3(kAmaxBlocks, 2), dim3(256), 0, st, reinterpret_cast<const float4*>(fmap1),
                       (long long)pa.src_items, pa.src_hdr, reinterpret_cast<const float4*>(fmap2),
                       (long long)(fmap2 ? (int64_t)B2 * h2 * w2 * (C / 4) : 0), pa.tgt_hdr));
  SDOF_LAUNCH_CHECK("corr_absmax_kernel");
  SDOF_CUDA(launch_pdl(corr_prep16_kernel, dim3(nb), dim3(256), 0, st, pa));
  SDOF_LAUNCH_CHECK("corr_prep16_kernel");
  return SDOF_OK;
}

// patch width (log2) for a level: the candidate among 32x4 / 16x8 / 8x16 target pixels with the fewest tiles (ties: wider)
static int pick_patch_shift(int h, int w) {
  int best = 5, best_tiles = 1 << 30;
  for (int s = 5; s >= 3; --s) {
    const int tiles = ceil_div(w, 1 << s) * ceil_div(h, kRM >> s);
    if (tiles < best_tiles) {
      best_tiles = tiles;
      best = s;
    }
  }
  return best;
}

// Main kernel on prepared operands.  out_half: pyramid stored as fp16 (layout in half elements) instead of fp32.
// B2 == B, or B2 == 1: one target (key-frame) operand shared by every pair.
int launch_corr_pyramid_parts(const void* src_ops, const void* tgt_ops, int B, int n1, int B2, int h2, int w2, int C, int fmt,
                              int out_half, void* pyramid, const sdof_pyramid_layout& lay, cudaStream_t st) {
  const int levels = lay.levels;
  if (!corr_res_supported(C, levels)) return SDOF_ERR_UNSUPPORTED;
  if (B > 65535) return SDOF_ERR_UNSUPPORTED;
  if (B2 != B && B2 != 1) return fail(SDOF_ERR_INVALID, "target operand batch must be B (%d) or 1, got %d", B, B2);
  int rc;
  if ((rc = check_ops_ptr(src_ops, "source")) || (rc = check_ops_ptr(tgt_ops, "target"))) return rc;
  const bool pow4 = (C & (C - 1)) == 0 && (__builtin_ctz(C) % 2 == 0);
  ResArgs ra;
  memset(&ra, 0, sizeof(ra));
  ResMaps maps;
  memset(&maps, 0, sizeof(maps));
  const void* src16 = reinterpret_cast<const uint8_t*>(src_ops) + kOpHdrBytes;
  const uint8_t* w = reinterpret_cast<const uint8_t*>(tgt_ops) + kOpHdrBytes;
  const void* tgt16[kRLevels] = {};
  int used_levels = 0;
  for (int l = 0; l < levels; ++l) {
    tgt16[l] = w;
    w += al256((int64_t)B2 * lay.h[l] * lay.w[l] * C * 2);
    if (lay.h[l] >= 1 && lay.w[l] >= 1) used_levels = l + 1;
  }
  if (used_levels == 0) return SDOF_OK;

  const CUtensorMapDataType dt = fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int slab_elems = kRSlabBytes / 2;
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)n1, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)n1 * C * 2};
    cuuint32_t box[3] = {(cuuint32_t)slab_elems, (cuuint32_t)kRN, 1};
    if ((rc = encode_map(&maps.src, dt, 3, src16, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "fmap1 (16-bit)"))) return rc;
  }
  ra.B = B;
  ra.n1 = n1;
  ra.m_tiles = ceil_div(n1, kRN);
  ra.levels = used_levels;
  ra.slab_elems = slab_elems;
  ra.kslabs = ceil_div(C, slab_elems);
  ra.divisor = sqrtf((float)C);
  ra.rsqrt_c = pow4 ? 1.0f / sqrtf((float)C) : 1.0f;
  ra.use_div = !pow4;
  ra.fmt = fmt;
  ra.tgt_shared = (B2 == 1 && B > 1) ? 1 : 0;
  ra.src_hdr = reinterpret_cast<const float*>(src_ops);
  ra.tgt_hdr = reinterpret_cast<const float*>(tgt_ops);
  {
    const char* d = getenv("SDOF_RES_DEBUG");
    ra.debug = d ? atoi(d) : 0;
  }
  ra.trace = nullptr;
#ifdef SDOF_RES_TRACE
  const char* trace_path = getenv("SDOF_RES_TRACE");
#else
  const char* trace_path = nullptr;
#endif
  const size_t trace_bytes = (size_t)sm_count() * kRTraceTiles * kRTraceSlots * sizeof(long long);
  if (trace_path && *trace_path) {
    SDOF_CUDA(cudaMalloc(reinterpret_cast<void**>(&ra.trace), trace_bytes));
    SDOF_CUDA(cudaMemsetAsync(ra.trace, 0, trace_bytes, st));
  }
  ra.tile_begin_level[0] = 0;
  for (int l = 0; l < used_levels; ++l) {
    const int pxs = pick_patch_shift(lay.h[l], lay.w[l]);
    const int patch_x = 1 << pxs, patch_y = kRM >> pxs;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)lay.w[l], (cuuint64_t)lay.h[l], (cuuint64_t)B2};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)lay.w[l] * C * 2, (cuuint64_t)lay.h[l] * lay.w[l] * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)slab_elems, (cuuint32_t)patch_x, (cuuint32_t)patch_y, 1};
    if ((rc = encode_map(&maps.tgt[l], dt, 4, tgt16[l], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "pooled fmap2 (16-bit)")))
      return rc;
    const int tyt = ceil_div(lay.h[l], patch_y), txt = ceil_div(lay.w[l], patch_x);
    if (tyt > 0x3fff || txt > 0x3fff) return SDOF_ERR_UNSUPPORTED;
    ra.pxs[l] = pxs;
    ra.tx_tiles[l] = txt;
    ra.tile_begin_level[l + 1] = ra.tile_begin_level[l] + tyt * txt;
    ra.lh[l] = lay.h[l];
    ra.lw[l] = lay.w[l];
    ra.wp[l] = lay.wp[l];
    ra.pitch[l] = lay.pitch[l];
    ra.out[l] = out_half ? static_cast<void*>(reinterpret_cast<__half*>(pyramid) + lay.offset[l])
                         : static_cast<void*>(reinterpret_cast<float*>(pyramid) + lay.offset[l]);
    ra.out_factor = reinterpret_cast<float*>(pyramid);   // fp16 layout: the first 128 bytes are the header
  }
  const long long total = (long long)B * ra.m_tiles * ra.tile_begin_level[used_levels];
  if (total > 0x7fffffff) return SDOF_ERR_UNSUPPORTED;
  ra.total_tiles = (int)total;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  if (ceil_div64(total, grid) + 1 > kRMaxTilesPerCta) return SDOF_ERR_UNSUPPORTED;  // tile table would not fit
  if (out_half) {
    SDOF_CUDA(cudaFuncSetAttribute(corr_pyramid_resident_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRSmemTotal));
    L2Window win;   // written once, read by 20 lookups: same window as sdof_corr_lookup_h
    win.ptr = pyramid;
    win.bytes = (size_t)lay.total_floats * 2;
    {
      const size_t cap = l2_persist_bytes();
      if (cap > 0 && win.bytes > cap) win.bytes = cap;
    }
    SDOF_CUDA(launch_pdl_win(corr_pyramid_resident_kernel<true>, dim3(grid), dim3(kRThreads), (size_t)kRSmemTotal, st, win, maps, ra));
  } else {
    SDOF_CUDA(cudaFuncSetAttribute(corr_pyramid_resident_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRSmemTotal));
    SDOF_CUDA(launch_pdl(corr_pyramid_resident_kernel<false>, dim3(grid), dim3(kRThreads), (size_t)kRSmemTotal, st, maps, ra));
  }
  SDOF_LAUNCH_CHECK("corr_pyramid_resident_kernel");
  if (ra.trace) {
    // experiments only: synchronises and dumps the stamps (header: ctas, tiles, slots, total tiles)
    std::vector<long long> host(trace_bytes / sizeof(long long));
    SDOF_CUDA(cudaStreamSynchronize(st));
    SDOF_CUDA(cudaMemcpy(host.data(), ra.trace, trace_bytes, cudaMemcpyDeviceToHost));
    SDOF_CUDA(cudaFree(ra.trace));
    if (FILE* f = fopen(trace_path, "wb")) {
      const long long hdr[4] = {grid, kRTraceTiles, kRTraceSlots, total};
      fwrite(hdr, sizeof(hdr), 1, f);
      fwrite(host.data(), 1, trace_bytes, f);
      fclose(f);
    }
  }
  return SDOF_OK;
}

// ---- the round-1 single-workspace entry points, on top of the split operand buffers (source part first)
static int split_workspace(void* workspace, int64_t workspace_bytes, int B, int n1, int h2, int w2, int C, int levels,
                           void** src_ops, void** tgt_ops) {
  const int64_t need = corr_res_workspace_bytes(B, n1, h2, w2, C, levels);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(SDOF_ERR_INVALID, "correlation workspace of %lld bytes required, got %lld", (long long)need,
                (long long)workspace_bytes);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return fail(SDOF_ERR_INVALID, "correlation workspace must be 256-byte aligned");
  *src_ops = workspace;
  *tgt_ops = reinterpret_cast<uint8_t*>(workspace) + corr_res_src_bytes(B, n1, C);
  return SDOF_OK;
}

int launch_corr_prepare_resident(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, int fmt,
                                 const sdof_pyramid_layout& lay, void* workspace, int64_t workspace_bytes, int part,
                                 cudaStream_t st) {
  if (!corr_res_supported(C, lay.levels)) return SDOF_ERR_UNSUPPORTED;
  void *so, *to;
  int rc = split_workspace(workspace, workspace_bytes, B, n1, h2, w2, C, lay.levels, &so, &to);
  if (rc) return rc;
  return launch_corr_prepare_parts((part & 1) ? fmap1 : nullptr, B, n1, so, (part & 2) ? fmap2 : nullptr, B, h2, w2, to, C,
                                   lay.levels, fmt, st);
}

int launch_corr_pyramid_prepared(int B, int n1, int h2, int w2, int C, int fmt, float* pyramid, const sdof_pyramid_layout& lay,
                                 void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (!corr_res_supported(C, lay.levels)) return SDOF_ERR_UNSUPPORTED;
  void *so, *to;
  int rc = split_workspace(workspace, workspace_bytes, B, n1, h2, w2, C, lay.levels, &so, &to);
  if (rc) return rc;
  return launch_corr_pyramid_parts(so, to, B, n1, B, h2, w2, C, fmt, 0, pyramid, lay, st);
}

int launch_corr_pyramid_resident(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, int fmt,
                                 float* pyramid, const sdof_pyramid_layout& lay, void* workspace, int64_t workspace_bytes,
                                 cudaStream_t st) {
  int rc = launch_corr_prepare_resident(fmap1, fmap2, B, n1, h2, w2, C, fmt, lay, workspace, workspace_bytes, 3, st);
  if (rc) return rc;
  return launch_corr_pyramid_prepared(B, n1, h2, w2, C, fmt, pyramid, lay, workspace, workspace_bytes, st);
}

}  // namespace sdof

// ----------------------------------------------------------------------------- C ABI (round 2): split operands, fp16 pyramid
extern "C" {

int sdof_corr_pyramid_layout_ex(int64_t rows, int h2, int w2, int levels, int elem_bytes, sdof_pyramid_layout* out) {
  using namespace sdof;
  if (elem_bytes == 4) return sdof_corr_pyramid_layout(rows, h2, w2, levels, out);
  SDOF_REQUIRE(elem_bytes == 2, "sdof_corr_pyramid_layout_ex: elem_bytes must be 4 (fp32) or 2 (fp16), got %d", elem_bytes);
  SDOF_REQUIRE(out != nullptr, "sdof_corr_pyramid_layout_ex: out is NULL");
  SDOF_REQUIRE(levels >= 1 && levels <= SDOF_MAX_LEVELS, "sdof_corr_pyramid_layout_ex: levels must be in [1,%d], got %d",
               SDOF_MAX_LEVELS, levels);
  SDOF_REQUIRE(rows >= 0 && h2 >= 1 && w2 >= 1, "sdof_corr_pyramid_layout_ex: bad sizes rows=%lld h2=%d w2=%d", (long long)rows, h2, w2);
  memset(out, 0, sizeof(*out));
  out->levels = levels;
  int64_t off = 64;  // 128-byte header: float[0] = factor (stored value * factor = correlation), written by the volume kernel
  for (int l = 0; l < levels; ++l) {
    const int h = h2 >> l, w = w2 >> l;
    const int wp = (w + 7) & ~7;  // rows start 16-byte aligned; an odd width always has a spare padding column
    out->h[l] = h;
    out->w[l] = w;
    out->wp[l] = wp;
    out->pitch[l] = (int64_t)h * wp;
    out->offset[l] = off;
    off += rows * out->pitch[l];
    off = (off + 63) & ~(int64_t)63;  // keep every level 128-byte aligned
  }
  out->total_floats = off;  // in ELEMENTS (halves here)
  return SDOF_OK;
}

int64_t sdof_corr_src_operand_bytes(int B, int h1, int w1, int C) { return sdof::corr_res_src_bytes(B, h1 * w1, C); }
int64_t sdof_corr_tgt_operand_bytes(int B2, int h2, int w2, int C, int levels) {
  return sdof::corr_res_tgt_bytes(B2, h2, w2, C, levels);
}

static int fmt_of(int precision) { return precision == SDOF_PREC_FP16 ? 0 : (precision == SDOF_PREC_BF16 ? 1 : -1); }

int sdof_corr_prepare_src(const float* fmap1, int B, int h1, int w1, int C, int precision, void* src_ops, int64_t src_bytes,
                          sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(fmap1 && src_ops, "sdof_corr_prepare_src: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h1 >= 1 && w1 >= 1 && C >= 8 && C % 8 == 0, "sdof_corr_prepare_src: bad sizes");
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(fmap1) & 15) == 0, "sdof_corr_prepare_src: fmap1 must be 16-byte aligned");
  if (fmt_of(precision) < 0) return fail(SDOF_ERR_UNSUPPORTED, "sdof_corr_prepare_src: precision must be FP16 or BF16");
  SDOF_REQUIRE(src_bytes >= corr_res_src_bytes(B, h1 * w1, C), "sdof_corr_prepare_src: operand buffer of %lld bytes required, got %lld",
               (long long)corr_res_src_bytes(B, h1 * w1, C), (long long)src_bytes);
  if (B == 0) return SDOF_OK;
  int rc = launch_corr_prepare_parts(fmap1, B, h1 * w1, src_ops, nullptr, 0, 1, 1, nullptr, C, 1, fmt_of(precision), as_stream(stream));
  if (rc == SDOF_ERR_UNSUPPORTED) return fail(rc, "sdof_corr_prepare_src: shape not supported by the resident kernel (C %% 8 == 0, C <= 256)");
  return rc;
}

int sdof_corr_prepare_tgt(const float* fmap2, int B2, int h2, int w2, int C, int levels, int precision, void* tgt_ops,
                          int64_t tgt_bytes, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(fmap2 && tgt_ops, "sdof_corr_prepare_tgt: NULL pointer");
  SDOF_REQUIRE(B2 >= 0 && h2 >= 1 && w2 >= 1 && C >= 8 && C % 8 == 0, "sdof_corr_prepare_tgt: bad sizes");
  SDOF_REQUIRE(levels >= 1 && levels <= 6, "sdof_corr_prepare_tgt: levels must be in [1,6], got %d", levels);
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(fmap2) & 15) == 0, "sdof_corr_prepare_tgt: fmap2 must be 16-byte aligned");
  if (fmt_of(precision) < 0) return fail(SDOF_ERR_UNSUPPORTED, "sdof_corr_prepare_tgt: precision must be FP16 or BF16");
  SDOF_REQUIRE(tgt_bytes >= corr_res_tgt_bytes(B2, h2, w2, C, levels), "sdof_corr_prepare_tgt: operand buffer of %lld bytes required, got %lld",
               (long long)corr_res_tgt_bytes(B2, h2, w2, C, levels), (long long)tgt_bytes);
  if (B2 == 0) return SDOF_OK;
  int rc = launch_corr_prepare_parts(nullptr, 0, 0, nullptr, fmap2, B2, h2, w2, tgt_ops, C, levels, fmt_of(precision), as_stream(stream));
  if (rc == SDOF_ERR_UNSUPPORTED) return fail(rc, "sdof_corr_prepare_tgt: shape not supported by the resident kernel (C %% 8 == 0, C <= 256)");
  return rc;
}

int sdof_corr_prepare_both(const float* fmap1, int B, int h1, int w1, void* src_ops, int64_t src_bytes, const float* fmap2, int B2,
                           int h2, int w2, int levels, void* tgt_ops, int64_t tgt_bytes, int C, int precision, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(fmap1 && fmap2 && src_ops && tgt_ops, "sdof_corr_prepare_both: NULL pointer");
  SDOF_REQUIRE(B >= 0 && B2 >= 0 && h1 >= 1 && w1 >= 1 && h2 >= 1 && w2 >= 1 && C >= 8 && C % 8 == 0, "sdof_corr_prepare_both: bad sizes");
  SDOF_REQUIRE(levels >= 1 && levels <= 6, "sdof_corr_prepare_both: levels must be in [1,6], got %d", levels);
  SDOF_REQUIRE(((reinterpret_cast<uintptr_t>(fmap1) | reinterpret_cast<uintptr_t>(fmap2)) & 15) == 0, "sdof_corr_prepare_both: feature maps must be 16-byte aligned");
  if (fmt_of(precision) < 0) return fail(SDOF_ERR_UNSUPPORTED, "sdof_corr_prepare_both: precision must be FP16 or BF16");
  SDOF_REQUIRE(src_bytes >= corr_res_src_bytes(B, h1 * w1, C) && tgt_bytes >= corr_res_tgt_bytes(B2, h2, w2, C, levels),
               "sdof_corr_prepare_both: operand buffers too small");
  if (B == 0 || B2 == 0) return SDOF_OK;
  int rc = launch_corr_prepare_parts(fmap1, B, h1 * w1, src_ops, fmap2, B2, h2, w2, tgt_ops, C, levels, fmt_of(precision), as_stream(stream));
  if (rc == SDOF_ERR_UNSUPPORTED) return fail(rc, "sdof_corr_prepare_both: shape not supported by the resident kernel (C %% 8 == 0, C <= 256)");
  return rc;
}

int sdof_corr_pyramid_from_parts(const void* src_ops, const void* tgt_ops, int B, int h1, int w1, int B2, int h2, int w2, int C,
                                 int levels, int precision, int elem_bytes, void* pyramid, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(src_ops && tgt_ops && pyramid, "sdof_corr_pyramid_from_parts: NULL pointer");
  SDOF_REQUIRE(B >= 0 && h1 >= 1 && w1 >= 1 && h2 >= 1 && w2 >= 1 && C >= 8 && C % 8 == 0, "sdof_corr_pyramid_from_parts: bad sizes");
  SDOF_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "sdof_corr_pyramid_from_parts: elem_bytes must be 2 or 4");
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(pyramid) & 127) == 0, "sdof_corr_pyramid_from_parts: pyramid must be 128-byte aligned");
  if (fmt_of(precision) < 0) return fail(SDOF_ERR_UNSUPPORTED, "sdof_corr_pyramid_from_parts: precision must be FP16 or BF16");
  sdof_pyramid_layout lay;
  int rc = sdof_corr_pyramid_layout_ex((int64_t)B * h1 * w1, h2, w2, levels, elem_bytes, &lay);
  if (rc) return rc;
  if (B == 0) return SDOF_OK;
  rc = launch_corr_pyramid_parts(src_ops, tgt_ops, B, h1 * w1, B2, h2, w2, C, fmt_of(precision), elem_bytes == 2, pyramid, lay,
                                 as_stream(stream));
  if (rc == SDOF_ERR_UNSUPPORTED) return fail(rc, "sdof_corr_pyramid_from_parts: shape not supported by the resident kernel");
  return rc;
}

}  // extern "C"
