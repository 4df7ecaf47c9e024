/*
 * sdof_b200.h — C ABI of libsdof_b200.so: the B200 (sm_100a) flow -> warp ->
 * mask/composite hot path of zyddnys/sd_animation_optical_flow.
 *
 * Plain C: device pointers, sizes and a CUDA stream handle; no torch types.
 * Every entry point is stream-ordered, performs no host synchronisation and
 * returns SDOF_OK (0) or an error code; sdof_last_error() gives the text.
 * All pointers are DEVICE pointers unless a parameter says "host".
 *
 * Each declaration cites the reference interface (file:line under the
 * reference checkout) that it replaces.  INTEGRATION.md shows the
 * reference-side bindings (ctypes stubs) a maintainer would add.
 */
#ifndef SDOF_B200_H
#define SDOF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDOF_ABI_VERSION 1
#define SDOF_MAX_LEVELS 8

typedef void* sdof_stream_t; /* cudaStream_t; NULL = legacy default stream */

enum sdof_status {
  SDOF_OK = 0,
  SDOF_ERR_INVALID = 1,     /* bad argument (null pointer, size, alignment) */
  SDOF_ERR_CUDA = 2,        /* a CUDA call or launch failed                  */
  SDOF_ERR_UNSUPPORTED = 3  /* shape/mode not supported by this build        */
};

/* arithmetic used for the all-pairs contraction (output is always fp32) */
enum sdof_precision {
  SDOF_PREC_TF32 = 0,   /* tcgen05 kind::tf32, one pass, features rounded to nearest */
  SDOF_PREC_3XTF32 = 1, /* tcgen05 kind::tf32, error-compensated (fp32-faithful) */
  SDOF_PREC_BF16 = 2,   /* tcgen05 kind::f16 on bf16 copies of the features     */
                        /* FP16/BF16 use the resident-operand kernel and build level l from the 2^l-pooled fmap2
                         * (avg-pooling is linear: RAFT/core/corr.py:68-72 does the same); TF32/3XTF32/FP32 pool the
                         * level-0 volume exactly like RAFT/core/corr.py:25-27 */
  SDOF_PREC_FP32 = 3,   /* CUDA-core fp32 FMA (exact-order checker on device)   */
  SDOF_PREC_FP16 = 4    /* tcgen05 kind::f16 on fp16 copies (11-bit significand = TF32 precision): the fast path */
};

int sdof_abi_version(void);
const char* sdof_last_error(void);

/* ------------------------------------------------------------------ C1 + C2
 * Layout of the correlation pyramid in HBM.  Level l holds, for each of the
 * `rows` = B*h1*w1 source pixels, an h[l] x w[l] map (h[l]=h2>>l, w[l]=w2>>l,
 * floor) stored row-major with row pitch wp[l] (w[l] rounded up to 4 floats so
 * that TMA stores are 16-byte aligned); pitch[l] = h[l]*wp[l].  Element (row, y, x) of level l is at
 *   pyramid[offset[l] + row*pitch[l] + y*wp[l] + x].
 * Replaces the list `CorrBlock.corr_pyramid` (RAFT/core/corr.py:15-27).       */
typedef struct sdof_pyramid_layout {
  int32_t levels;
  int32_t h[SDOF_MAX_LEVELS];
  int32_t w[SDOF_MAX_LEVELS];
  int32_t wp[SDOF_MAX_LEVELS];
  int64_t pitch[SDOF_MAX_LEVELS];
  int64_t offset[SDOF_MAX_LEVELS];
  int64_t total_floats;
} sdof_pyramid_layout;

int sdof_corr_pyramid_layout(int64_t rows, int h2, int w2, int levels, sdof_pyramid_layout* out /* host */);

/* Scratch bytes sdof_corr_volume_pyramid needs for `precision` (rounded / split / bf16 copies of
 * the features; 0 for FP32). */
int64_t sdof_corr_volume_workspace_bytes(int B, int h1, int w1, int h2, int w2, int C, int levels, int precision);

/* All-pairs correlation volume + its average-pooled pyramid in one pass:
 *   level0[b, i, j] = <fmap1[b,i,:], fmap2[b,j,:]> / sqrt(C);  level l+1 = avg_pool2d(level l, 2, 2)
 * fmap1 [B, h1*w1, C], fmap2 [B, h2*w2, C]: fp32, channels-last, contiguous, 16-byte aligned.
 * Replaces CorrBlock.corr + CorrBlock.__init__ (RAFT/core/corr.py:12-27, 52-60):
 * torch.matmul + divide + 3x avg_pool2d.                                      */
int sdof_corr_volume_pyramid(const float* fmap1, const float* fmap2, int B, int h1, int w1, int h2, int w2, int C,
                             int levels, int precision, float* pyramid, void* workspace, int64_t workspace_bytes,
                             sdof_stream_t stream);

/* The FP16 / BF16 path in two steps, so that operands can be reused: in the key-frame scheme every pair of a
 * key frame shares one feature map (RAFT is run as image1 = current frame, image2 = key frame, so fmap2 and
 * its pooled levels are constant across those pairs).
 *   sdof_corr_prepare_operands : 16-bit copies of fmap1 (pre-scaled by 1/sqrt(C) when that is exact) and of every
 *                                avg-pooled level of fmap2 into `workspace`; parts: 1 = fmap1, 2 = fmap2, 3 = both
 *                                (either pointer may be NULL when its part is not requested)
 *   sdof_corr_pyramid_from_operands : the tcgen05 kernel alone, on a prepared workspace.
 * workspace: sdof_corr_volume_workspace_bytes(..., precision) bytes, 256-byte aligned.  SDOF_ERR_UNSUPPORTED for
 * other precisions or C > 256.                                                                                  */
int sdof_corr_prepare_operands(const float* fmap1, const float* fmap2, int B, int h1, int w1, int h2, int w2, int C,
                               int levels, int precision, int parts, void* workspace, int64_t workspace_bytes,
                               sdof_stream_t stream);
int sdof_corr_pyramid_from_operands(int B, int h1, int w1, int h2, int w2, int C, int levels, int precision,
                                    float* pyramid, void* workspace, int64_t workspace_bytes, sdof_stream_t stream);

/* Round 2: the same path with (a) the pyramid optionally STORED IN FP16 (elem_bytes = 2: half the bytes of the store-bound
 * volume kernel and of every lookup; the 768x512 pyramid, 100 MB, then stays resident in the 126 MB L2 across the 20 lookups
 * of a pair; an 11-bit significand is what the TF32 convolution consuming the lookup keeps of it anyway), (b) the two
 * operands in SEPARATE buffers, so that one target (key-frame) operand serves every pair of a call (B2 = 1) and is built
 * once per key frame (ofgen_keyframe_inpaint.py:602-625 warps every frame against the same references), and (c) AUTO-RANGED
 * 16-bit operands: a per-tensor power-of-two scale from an abs-max pass, undone exactly in the epilogue, so FP16 neither
 * saturates on large nor loses bits on tiny feature magnitudes (the reference's SGEMM has fp32 range, RAFT/core/corr.py:58).
 *   sdof_corr_pyramid_layout_ex : layout for elem_bytes 4 (== sdof_corr_pyramid_layout) or 2 (row pitch rounded up to
 *                                 8 halves, offsets / pitches / total_floats counted in ELEMENTS)
 *   sdof_corr_prepare_src / _tgt : fmap1 [B,h1*w1,C] -> src_ops; fmap2 [B2,h2,w2,C] + its avg-pooled levels -> tgt_ops
 *                                 (buffers of sdof_corr_{src,tgt}_operand_bytes bytes, 256-byte aligned)
 *   sdof_corr_pyramid_from_parts: the tcgen05 kernel; B2 == B or 1; pyramid 128-byte aligned.
 * FP16 / BF16 only, C % 8 == 0, C <= 256 (SDOF_ERR_UNSUPPORTED otherwise).                                        */
int sdof_corr_pyramid_layout_ex(int64_t rows, int h2, int w2, int levels, int elem_bytes, sdof_pyramid_layout* out /* host */);
int64_t sdof_corr_src_operand_bytes(int B, int h1, int w1, int C);
int64_t sdof_corr_tgt_operand_bytes(int B2, int h2, int w2, int C, int levels);
int sdof_corr_prepare_src(const float* fmap1, int B, int h1, int w1, int C, int precision, void* src_ops, int64_t src_bytes,
                          sdof_stream_t stream);
int sdof_corr_prepare_tgt(const float* fmap2, int B2, int h2, int w2, int C, int levels, int precision, void* tgt_ops,
                          int64_t tgt_bytes, sdof_stream_t stream);
/* Both operands of a pair in ONE abs-max launch and ONE conversion launch (instead of two each). */
int sdof_corr_prepare_both(const float* fmap1, int B, int h1, int w1, void* src_ops, int64_t src_bytes, const float* fmap2, int B2,
                           int h2, int w2, int levels, void* tgt_ops, int64_t tgt_bytes, int C, int precision, sdof_stream_t stream);
int sdof_corr_pyramid_from_parts(const void* src_ops, const void* tgt_ops, int B, int h1, int w1, int B2, int h2, int w2, int C,
                                 int levels, int precision, int elem_bytes, void* pyramid, sdof_stream_t stream);

/* ------------------------------------------------------------------------ C3
 * Windowed bilinear lookup in all pyramid levels for one GRU iteration.
 * coords [B,2,h1,w1] (x then y, level-0 pixels of the target map) ->
 * out [B, levels*(2r+1)^2, h1, w1] fp32, channel = l*(2r+1)^2 + (2r+1)*ix + iy,
 * zeros outside the map.
 * Replaces CorrBlock.__call__ (RAFT/core/corr.py:29-50) + bilinear_sampler
 * (RAFT/core/utils/utils.py:57-71): 4x grid_sample + cat + permute + contiguous. */
int sdof_corr_lookup(const float* pyramid, const float* coords, int B, int h1, int w1, int h2, int w2, int levels,
                     int radius, float* out, sdof_stream_t stream);

/* Same lookup with channels-last tensors: coords [B,h1,w1,2], out [B,h1,w1,levels*(2r+1)^2] (same channel order).
 * Used by the NHWC update loop (raft_fast.py) so that the 1x1 convolution convc1 (RAFT/core/update.py:83) reads
 * it without a layout change.                                                                                  */
int sdof_corr_lookup_nhwc(const float* pyramid, const float* coords, int B, int h1, int w1, int h2, int w2, int levels,
                          int radius, float* out, sdof_stream_t stream);

/* Either lookup on a pyramid of elem_bytes 4 (fp32) or 2 (fp16, sdof_corr_pyramid_layout_ex); channels_last selects the
 * tensor layouts of sdof_corr_lookup_nhwc.  Output is always fp32. */
int sdof_corr_lookup_ex(const void* pyramid, int elem_bytes, const float* coords, int B, int h1, int w1, int h2, int w2, int levels,
                        int radius, float* out, int channels_last, sdof_stream_t stream);

/* ------------------------------------------------------------------------ K1
 * On-the-fly windowed correlation, no volume.  Same contract as the pybind op
 *   alt_cuda_corr.forward(fmap1, fmap2, coords, radius) -> [corr]
 * (RAFT/alt_cuda_corr/correlation.cpp:23-33,51-54; kernel
 * correlation_kernel.cu:18-119; launcher :260-286):
 * fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C] fp32 channels-last, coords [B,N,H1,W1,2]
 * (x,y in fmap2 pixels) -> corr [B,N,(2r+1)^2,H1,W1], UNNORMALISED, every element
 * written (no zero-initialisation needed).  C must be a multiple of 4.          */
int sdof_alt_corr_forward(const float* fmap1, const float* fmap2, const float* coords, int B, int H1, int W1, int H2,
                          int W2, int C, int N, int radius, float* corr, sdof_stream_t stream);

/* C4: one level of AlternateCorrBlock.__call__ (RAFT/core/corr.py:74-91) without the
 * permute/contiguous copies: coords [B,2,H1,W1] planar, scaled by coord_scale (=1/2^l)
 * inside the kernel; writes channels [chan_offset, chan_offset+(2r+1)^2) of
 * out [B, out_channels, H1, W1], multiplied by out_scale (=1/sqrt(C)).            */
int sdof_alt_corr_level(const float* fmap1, const float* fmap2_level, const float* coords, int B, int H1, int W1,
                        int H2, int W2, int C, int radius, float coord_scale, float out_scale, int chan_offset,
                        int out_channels, float* out, sdof_stream_t stream);

/* 2x2 average pool of a channels-last map [B,H,W,C] -> [B,H/2,W/2,C] (floor), used to
 * build AlternateCorrBlock's fmap2 pyramid (RAFT/core/corr.py:68-72).            */
int sdof_avgpool2_nhwc(const float* in, int B, int H, int W, int C, float* out, sdof_stream_t stream);

/* -------------------------------------------------------------------- W1 / W2
 * Backward warp.  dst[b,y,x,:] = sample(src[b or 0], x + sign*flow[b,y,x,0], y + sign*flow[b,y,x,1]).
 * The sampling map is float32(double(x) + double(sign*flow)), which is both
 * pdcnet_of.warp_frame's map (sign=+1, pdcnet_of.py:34-42) and ofgen.warp_frame's
 * (sign=-1, ofgen.py:37-43).  src [Bs,Hs,Ws,C] (Bs = B, or 1 to share one key frame),
 * flow [B,H,W,2], dst [B,H,W,C]; C in 1..4.
 * cubic: bit-exact cv2.remap(INTER_CUBIC, BORDER_CONSTANT 0) for u8 (1/32-pixel
 * coordinates, 2^15 fixed-point weight table); same quantisation, fp32 weights for f32.
 * bilinear: grid_sample(bilinear, zeros, align_corners=True) semantics at exact
 * pixel coordinates (RAFT/core/utils/utils.py:57-71); u8 output = round-half-even. */
int sdof_warp_cubic_u8(const uint8_t* src, const float* flow, int B, int src_batched, int Hs, int Ws, int C, int H,
                       int W, float sign, uint8_t* dst, sdof_stream_t stream);
int sdof_warp_cubic_f32(const float* src, const float* flow, int B, int src_batched, int Hs, int Ws, int C, int H,
                        int W, float sign, float* dst, sdof_stream_t stream);
int sdof_warp_bilinear_u8(const uint8_t* src, const float* flow, int B, int src_batched, int Hs, int Ws, int C, int H,
                          int W, float sign, uint8_t* dst, sdof_stream_t stream);
int sdof_warp_bilinear_f32(const float* src, const float* flow, int B, int src_batched, int Hs, int Ws, int C, int H,
                           int W, float sign, float* dst, sdof_stream_t stream);

/* W3: cv2.resize(src, (Wd, Hd), interpolation=cv2.INTER_CUBIC) for float32 images, the two resizes of warp_frame_latent
 * (pdcnet_of.py:24,30; dup ofgen_pixel_inpaint.py:92-103) around the cubic warp: src [B,Hs,Ws,C] -> dst [B,Hd,Wd,C], both
 * channels-last.  OpenCV's float path restated (1/scale in double, float32 a = -0.75 weights, border replication, no
 * antialiasing); agrees with cv2 to 2e-7 of the image range at the x8 / /8 ratios the reference uses. */
int sdof_resize_cubic_f32(const float* src, int B, int Hs, int Ws, int C, int Hd, int Wd, float* dst, sdof_stream_t stream);

/* Host-side copy of the 1024x16 int16 bicubic weight table the u8 kernel uses
 * (out: host int16[16384]); lets CPU tests pin it against OpenCV without a GPU. */
int sdof_cubic_table_i16(int16_t* out /* host */);
/* Host-side row half-widths of cv2.getStructuringElement(MORPH_ELLIPSE,(k,k)). */
int sdof_ellipse_half_widths(int ksize, int32_t* out /* host, ksize entries */);

/* ------------------------------------------------------------------------ M1
 * confidence = softmax(weight_map, dim=1)[:,0], log_confidence = log_softmax(...)[:,0]
 * (pdcnet_of.py:72-74).  weight_map [B,K,H,W] -> conf, logconf [B,H,W] (either may be NULL). */
int sdof_confidence_softmax(const float* weight_map, int B, int K, int H, int W, float* conf, float* logconf,
                            sdof_stream_t stream);

/* M2: travel distance of of_calc (ofgen_pixel_inpaint.py:105-118). flow [B,H,W,2], conf [B,H,W] -> v [B,H,W]. */
int sdof_travel_distance(const float* flow, const float* conf, int B, int H, int W, float conf_thres, float* v,
                         sdof_stream_t stream);

/* M3: generate_mask (ofgen_pixel_inpaint.py:262-267; ofgen_keyframe_inpaint.py:317-322):
 * mask = dilate_ellipse(255*(conf < thres), ksize); log_conf[conf < thres] = 0 in place
 * (log_conf may be NULL).  conf [B,H,W] -> mask u8 [B,H,W].                       */
int sdof_generate_mask(const float* conf, float* log_conf, int B, int H, int W, float thres, int ksize,
                       uint8_t* mask, sdof_stream_t stream);

/* cv2.dilate(src, getStructuringElement(MORPH_ELLIPSE,(ksize,ksize))) on u8 [B,H,W];
 * invert != 0 dilates (255 - src) (ofgen_keyframe_inpaint.py:772-774).  ksize odd, <= 31. */
int sdof_dilate_ellipse_u8(const uint8_t* src, int B, int H, int W, int ksize, int invert, uint8_t* dst,
                           sdof_stream_t stream);

/* M6: expand_mask (ofgen_keyframe_inpaint.py:968-973): out = mask | dilate(255*(gray(|laplacian(img)| mod 256) > 20)).
 * mask u8 [B,H,W], image u8 [B,H,W,3], scratch u8 [B,H,W].                        */
int sdof_expand_mask(const uint8_t* mask, const uint8_t* image, int B, int H, int W, int ksize, uint8_t* scratch,
                     uint8_t* out, sdof_stream_t stream);

/* M4: mix_propagated_ai_frame (ofgen_pixel_inpaint.py:251-260). raw, warped, out u8 [B,H,W,C]; mask u8 [B,H,W]. */
int sdof_mix_propagated(const uint8_t* raw, const uint8_t* warped, const uint8_t* mask, int B, int H, int W, int C,
                        float ppw, uint8_t* out, sdof_stream_t stream);

/* M5 inner step: merge_images(method='naive') (ofgen_keyframe_inpaint.py:676-681): out = mask==255 ? second : base. */
int sdof_merge_select(const uint8_t* base, const uint8_t* second, const uint8_t* mask, int B, int H, int W, int C,
                      uint8_t* out, sdof_stream_t stream);

/* M5: greedy multi-reference composite (ofgen_keyframe_inpaint.py:995-1024, :741-770).
 * flow_mat [n,H,W,3] fp32 (flow x, flow y, confidence) is updated in place exactly as the
 * reference does (binarise conf > thres, subtract covered pixels, clip); ai_frames u8 [n,H,W,3].
 * Outputs: ret u8 [H,W,3], mask u8 [H,W], order int32[n] (device) = reference chosen per round.
 * workspace: sdof_greedy_workspace_bytes(n,H,W) bytes.                             */
int64_t sdof_greedy_workspace_bytes(int n, int H, int W);
int sdof_greedy_composite(float* flow_mat, const uint8_t* ai_frames, int n, int H, int W, float thres, uint8_t* ret,
                          uint8_t* mask, int32_t* order, void* workspace, sdof_stream_t stream);

/* A2: KeyframeConv score (ofgen_keyframe_inpaint.py:664-668): sums[s] = sum of the confidence
 * channel over `per_source` pixels of flow_mat [S, per_source, 3]; sums double[S] (device). */
int sdof_confidence_sums(const float* flow_mat, int S, int64_t per_source, double* sums, sdof_stream_t stream);

/* ------------------------------------------------- fused W1 + M1 + M3 + M4 (26 B/pixel)
 * One pass per non-key frame (ofgen_pixel_inpaint.py:335-349 with ppw = 1):
 *   conf  = softmax(weight_map)[0];  mask = dilate_ellipse(255*(conf < thres), ksize)
 *   out   = mask > 127 ? base : cubic_warp(src, x + flow)
 * src u8 [Bs,H,W,3] stylised key frame, base u8 [B,H,W,3], flow [B,H,W,2],
 * weight_map [B,2,H,W] -> out u8 [B,H,W,3], mask u8 [B,H,W].                      */
int sdof_warp_mask_composite(const uint8_t* src, const uint8_t* base, const float* flow, const float* weight_map,
                             int B, int src_batched, int H, int W, float thres, int ksize, uint8_t* out,
                             uint8_t* mask, sdof_stream_t stream);

/* ------------------------------------------------ RAFT update-loop glue
 * Element-wise steps between the cuDNN convolutions of RAFT's update block (RAFT/core/update.py:79-136) on dense
 * channels-last buffers [npix, C]; they replace the cat / relu / sigmoid / tanh / mul / add kernels of the eager
 * reference.  All pointers 16-byte aligned, strides/offsets in floats.
 *   relu_scatter : dst[:, off : off+C_valid] = relu(src[:, :C_valid]) into one or two (dst2 may be NULL) buffers
 *   gru_rh       : rhx[:, 0:hidden] = sigmoid(zr[:, hidden:2*hidden]) * h           (update.py:48-49, 55-56)
 *   gru_update   : h = (1-sigmoid(z))*h + sigmoid(z)*tanh(q), z = zr[:, 0:hidden]; also written to hx[:, 0:hidden]
 *   flow_update  : coords1 += delta (delta may be NULL); flow = coords1 - pixel grid -> flow [npix,2] and the
 *                  flow slots of hx / rhx (either may be NULL)                      (raft.py:126-131)
 *   convex_upsample : RAFT.upsample_flow (raft.py:72-83): mask [B,h,w,576] (times mask_scale), flow [B,h,w,2]
 *                  -> up [B,8h,8w,2]                                                                           */
int sdof_relu_scatter(const float* src, const float* src2, const float* bias, int64_t npix, int C, float* dst1, int dst1_stride, int dst1_off,
                      float* dst2, int dst2_stride, int dst2_off, int C_valid, sdof_stream_t stream);
int sdof_gru_rh(const float* zr, const float* bias_zr, const float* h, float* rhx, int64_t npix, int hidden, int rhx_stride,
                int bias_map, int zr_channels, sdof_stream_t stream);
int sdof_gru_update(const float* zr, const float* bias_zr, const float* q, const float* bias_q, float* h, float* hx,
                    int64_t npix, int hidden, int hx_stride, int bias_map, int zr_channels, sdof_stream_t stream);
/* zr_channels = 2*hidden: zr = [z | r].  zr_channels = 3*hidden: zr = [z | r | q_x], where q_x is the share of convq that
 * does not depend on r (its [motion | flow] input channels), computed by the same convolution as z and r; gru_update adds
 * it to q (which then only holds the convolution of r*h). */
/* bias_map = 0: bias_zr [2*hidden], bias_q [hidden] per-channel vectors; bias_map = 1: per-pixel maps [npix][2*hidden] and
 * [npix][hidden] holding bias + the convolution of the iteration-invariant context features (computed once per pair). */
int sdof_flow_update(const float* delta, float delta_bias_x, float delta_bias_y, float* coords1, float* flow, float* hx,
                     int hx_stride, int hx_off, float* rhx, int rhx_stride, int rhx_off, int B, int h, int w,
                     sdof_stream_t stream);
int sdof_convex_upsample(const float* mask, const float* mask_bias, float mask_scale, const float* flow, int B, int h, int w,
                         float* up, sdof_stream_t stream);
/* The convolution biases (bias*, delta_bias_*, mask_bias; NULL / 0 = none) are added here, so the cuDNN convolutions
 * run bias-free and no separate bias pass exists.
 * InstanceNorm2d (no affine, eps) + optional ReLU over `planes` = N*C contiguous planes of hw floats (NCHW), in place
 * or out of place: relu(norm(conv(x))) of the feature encoder (RAFT/core/extractor.py:49-50, 172-173).            */
int sdof_instnorm_relu_nchw(const float* x, float* y, int64_t planes, int64_t hw, float eps, int relu, sdof_stream_t stream);
/* Channels-last (NHWC) form of the same layer, so the encoder convolutions run without layout transposes:
 *   stats [N][C][2] fp64 (sum, sum of squares over the hw pixels of each image and channel), ZEROED by the caller;
 *   apply: y = relu?((x - mean) * rsqrt(var + eps)); with `residual` != NULL additionally y = relu(residual + y),
 *   the tail of ResidualBlock.forward (RAFT/core/extractor.py:49-58).  x, y, residual: dense [N, hw, C], C % 4 == 0.
 *   sdof_add_relu: y = relu(a + b) over n floats (the same tail for cnet, whose BatchNorm is folded into the convs). */
int sdof_instnorm_stats_nhwc(const float* x, int N, int64_t hw, int C, double* stats, sdof_stream_t stream);
int sdof_instnorm_apply_nhwc(const float* x, const double* stats, const float* residual, float* y, int N, int64_t hw, int C,
                             float eps, int relu, sdof_stream_t stream);
int sdof_add_relu(const float* a, const float* b, float* y, int64_t n, sdof_stream_t stream);
/* The same two kernels on fp16 activations (x, y, residual: dense [N, hw, C] halves, 8-byte aligned quads; statistics still
 * accumulate in fp32 partials / fp64 atomics): the feature encoder can run its cuDNN convolutions in fp16 -- the same 11-bit
 * operand precision TF32 gives, at half the activation bytes. */
int sdof_instnorm_stats_nhwc_h(const void* x, int N, int64_t hw, int C, double* stats, sdof_stream_t stream);
int sdof_instnorm_apply_nhwc_h(const void* x, const double* stats, const void* residual, void* y, int N, int64_t hw, int C,
                               float eps, int relu, sdof_stream_t stream);
/* Input side of RAFT_2.calc / RAFT.forward (ofgen.py:72-76, RAFT/core/raft.py:89-90, utils/utils.py:7-19) in one pass:
 * img u8 [B,H,W,3] -> out f32 [B,Hp,Wp,Cout] (NHWC) = 2*(x/255)-1 of the replicate-padded frame; pixel (y,x) of out reads
 * img at (clamp(y-top), clamp(x-left)).  Cout = 3, or 4 with a zero fourth channel (lets cuDNN run the 7x7 stem
 * convolution on tensor cores with a zero-padded filter).  swap_rb != 0 reads BGR frames as RGB (the `[:, :, ::-1]` of
 * RAFT_2.calc, ofgen.py:72-73). */
int sdof_normalize_pad_u8_nhwc(const uint8_t* img, int B, int H, int W, int top, int left, int Hp, int Wp, int Cout, int swap_rb,
                               float* out, sdof_stream_t stream);
/* The two convolutions of the update block that are too small / too thin for a tensor-core library kernel:
 *   sdof_conv7x7_c2_relu : BasicMotionEncoder.convf1 (RAFT/core/update.py:85,93): out[B,h,w,128] =
 *                          relu(conv7x7(flow[B,h,w,2], pad 3) + bias); wT = weight[128,2,7,7] permuted to [7,7,2,128].
 *   sdof_flowhead2_update: FlowHead.conv2 (update.py:10,14) + the coords update of RAFT.forward (raft.py:128-131):
 *                          delta = conv3x3(x[B,h,w,256], pad 1) + bias; coords1 += delta; flow = coords1 - grid, written to
 *                          `flow` and the flow slots of hx / rhx like sdof_flow_update; w2 = weight[2,256,3,3] permuted
 *                          to [3,3,2,256]; scratch = B*h*w*18 floats (per-pixel tap products).  fp32 FMA accumulation. */
int sdof_conv7x7_c2_relu(const float* flow, const float* wT, const float* bias, float* out, int B, int h, int w, sdof_stream_t stream);
int sdof_flowhead2_update(const float* x, const float* w2, float bias_x, float bias_y, float* coords1, float* flow, float* hx,
                          int hx_stride, int hx_off, float* rhx, int rhx_stride, int rhx_off, int B, int h, int w,
                          float* scratch, sdof_stream_t stream);

/* The SepConvGRU of the update block (RAFT/core/update.py:32-60) on tcgen05, one GRU pass = two launches (csrc/conv_tc.cu):
 * tap-shifted implicit GEMM by TMA (activations NHWC fp16, out-of-image taps zero-filled = the convolution's padding),
 * fp16 operands / fp32 accumulation in TMEM, gate arithmetic in the epilogue.  `horizontal` = 1: the 1x5 pass
 * (convz1/convr1/convq1), 0: the 5x1 pass (convz2/convr2/convq2).  All fp16 buffers 128-byte aligned.
 *   sdof_motion_tail16 : hx16[:, 128:254] = relu(mc + mf + bias)[:, :126], hx16[:, 254:256] = flow: the motion encoder's
 *                        output convolution (update.py:95-96; mc, mf = its two partial sums [npix,128]) written as the
 *                        fp16 GRU input [npix, hx16_stride]; channels [0,128) of hx16 hold the hidden state
 *   sdof_gru_zr_tc     : [z | r | q_x] = conv(hx16[B,h,w,256], w_zr16[384][5][256]); z = sigmoid(. + zrmap[:, :128]) -> z
 *                        [npix,128] fp32; rh16 = fp16(sigmoid(. + zrmap[:, 128:]) * h) [npix,128]; qx = the r-independent
 *                        share of convq [npix,128] fp32.  zrmap [npix,256] / qmap [npix,128] = bias + convolution of the
 *                        iteration-invariant context features (computed once per pair)
 *   sdof_gru_q_tc      : q = conv(rh16, w_q16[128][5][128]); h = (1 - z) h + z tanh(q + qx + qmap) in place (fp32) and
 *                        hx16[:, 0:128] = fp16(h)                                                                    */
int sdof_motion_tail16(const float* mc, const float* mf, const float* bias, const float* flow, int64_t npix, void* hx16, int hx16_stride,
                       sdof_stream_t stream);
int sdof_gru_zr_tc(const void* hx16, const void* w_zr16, const float* zrmap, const float* h, int B, int hh, int ww, int horizontal,
                   float* z, void* rh16, float* qx, sdof_stream_t stream);
int sdof_gru_q_tc(const void* rh16, const void* w_q16, const float* qmap, const float* qx, const float* z, int B, int hh, int ww,
                  int horizontal, float* h, void* hx16, int hx16_stride, sdof_stream_t stream);

/* The update block with fp16 activations (csrc/raft_glue16.cu): the glue kernels above for an update block whose cuDNN
 * convolutions run in fp16 (tensor-op, fp32 accumulation; 11-bit operands like TF32) -- the twelve convolutions of one
 * iteration take 93 us instead of 113 us at 768x512 batch 1.  fp16 buffers are dense channels-last, 8-byte aligned; the hidden
 * state's master copy, coordinates, flow and the per-pair bias maps stay fp32.  hidden = 128 (the basic model).
 *   sdof_corr_lookup_h       : sdof_corr_lookup_ex (channels-last) with fp16 output rows padded to out_channels (zeros)
 *   sdof_conv7x7_c2_relu_h   : sdof_conv7x7_c2_relu with fp16 output
 *   sdof_motion_tail16_h     : sdof_motion_tail16 reading fp16 partial sums
 *   sdof_gru_rh_h            : rh16 = sigmoid(zr16[:, 128:256] + zrmap[:, 128:256]) * h          (zr16 [npix, zr_channels])
 *   sdof_gru_update_h        : h = (1-z) h + z tanh(q16 + zr16[:, 256:384] + qmap), z = sigmoid(zr16[:, :128] + zrmap[:, :128]);
 *                              writes h (fp32, in place), hx16[:, :128] and the dense copy h16 (may be NULL)
 *   sdof_flowhead2_taps_h    : the tap products of sdof_flowhead2_update from fp16 activations, then
 *   sdof_flowhead2_gather_update : its 9-neighbour gather + coords / flow update                                        */
int sdof_corr_lookup_h(const void* pyramid, int elem_bytes, const float* coords, int B, int h1, int w1, int h2, int w2, int levels, int radius,
                       void* out16, int out_channels, sdof_stream_t stream);
int sdof_conv7x7_c2_relu_h(const float* flow, const float* wT, const float* bias, void* out16, int B, int h, int w, sdof_stream_t stream);
/* Deferred coords update of the update loop (RAFT/core/raft.py:128-131 `coords1 = coords1 + delta_flow`): the flow head of iteration i
 * only leaves its tap products (sdof_flowhead2_taps_h); iteration i+1 applies coords = coords_in + (bias + sum of the 9 neighbours'
 * taps) where the coordinates are consumed, one launch and one dependency less per iteration:
 *   sdof_corr_lookup_gather_h     : sdof_corr_lookup_h on the updated coordinates; also writes them to coords_out (must not alias
 *                                   coords_in) and flow_out = coords_out - pixel grid.  taps == NULL: coordinates pass through.
 *   sdof_conv7x7_c2_relu_coords_h : sdof_conv7x7_c2_relu_h (BasicMotionEncoder.convf1, update.py:85,93) on flow = updated
 *                                   coordinates - pixel grid, reading coords_in and the taps only (it runs beside the lookup).
 * Both use the summation order of sdof_flowhead2_gather_update: identical bits. */
int sdof_corr_lookup_gather_h(const void* pyramid, int elem_bytes, const float* coords_in, const float* taps, float bias_x, float bias_y,
                              float* coords_out, float* flow_out, int B, int h1, int w1, int h2, int w2, int levels, int radius, void* out16,
                              int out_channels, sdof_stream_t stream);
int sdof_conv7x7_c2_relu_coords_h(const float* coords, const float* taps, float tap_bias_x, float tap_bias_y, const float* wT, const float* bias,
                                  void* out16, int B, int h, int w, sdof_stream_t stream);
/* BasicMotionEncoder.convf1 (update.py:85,93) as a tensor-core GEMM: the im2col rows of the 7x7 x 2-channel convolution of
 * flow = coords + gather(taps) - pixel grid (taps may be NULL), fp16 with a hi/lo split of the flow:
 *   out16 [B,h,w,out_channels] halves, row = [49 taps x (fx_hi, fy_hi) | 49 taps x (fx_lo, fy_lo) | zeros], tap = ky*7 + kx, taps outside
 *   the image 0.  A 1x1 convolution of these rows with the filter [128][tap][ci] repeated for both halves is convf1. */
int sdof_flow_im2col7_h(const float* coords, const float* taps, float tap_bias_x, float tap_bias_y, void* out16, int out_channels, int B, int h,
                        int w, sdof_stream_t stream);
int sdof_motion_tail16_h(const void* mc16, const void* mf16, const float* bias, const float* flow, int64_t npix, void* hx16, int hx16_stride,
                         sdof_stream_t stream);
int sdof_gru_rh_h(const void* zr16, int zr_channels, const float* zrmap, const float* h, void* rh16, int64_t npix, sdof_stream_t stream);
int sdof_gru_update_h(const void* zr16, const float* zrmap, const void* q16, const float* qmap, float* h, void* hx16, int hx16_stride, void* h16,
                      int64_t npix, sdof_stream_t stream);
int sdof_flowhead2_taps_h(const void* x16, const float* w2, int64_t npix, float* scratch, sdof_stream_t stream);
int sdof_flowhead2_gather_update(const float* scratch, float bias_x, float bias_y, float* coords1, float* flow, float* hx, int hx_stride,
                                 int hx_off, int B, int h, int w, sdof_stream_t stream);

/* ---------------------------------------------------------------- before the path: key-frame detector
 * frame_generator's edge-change detector (ofgen_pixel_inpaint.py:127-176, 300-312), bit-exact to OpenCV:
 *   sdof_detect_edges   : edges = cv2.dilate(cv2.Canny(V(frame), low, high), ones(k,k)) with V = max(B,G,R) (8-bit HSV
 *                         value channel).  low < 0 or high < 0 selects the reference's rule
 *                         low, high = int(max(0, (1-1/3)*median(V))), int(min(255, (1+1/3)*median(V))) (:158-161).
 *                         frame_bgr u8 [H,W,3], edges u8 [H,W] (0/255).  Synchronises the stream (hysteresis runs to a
 *                         global fixed point; the host relaunches while any tile changed).
 *   sdof_abs_diff_sum_u8: sum |a - b| over n bytes into *sum (device, 8 bytes): mean_pixel_distance (:132-139) = sum / n. */
int64_t sdof_detect_edges_workspace_bytes(int H, int W);
int sdof_detect_edges(const uint8_t* frame_bgr, int H, int W, int dilate_k, int low, int high, uint8_t* edges, void* workspace,
                      int64_t workspace_bytes, sdof_stream_t stream);
int sdof_abs_diff_sum_u8(const uint8_t* a, const uint8_t* b, int64_t n, unsigned long long* sum, sdof_stream_t stream);

/* ---------------------------------------------------------------- after the path: mask blur, composite, latent mask
 * What GuidedLDM.img2img_inpaint does to the warped frame and the inpainting mask before Stable Diffusion runs
 * (guided_ldm_inpainting.py:290-309), bit-exact to Pillow (the library the reference calls there):
 *   sdof_mask_blur_composite : blurred = mask.filter(ImageFilter.GaussianBlur(mask_blur))            (:292-293)
 *                              out = Image.composite(reference, image, blurred)                        (:298)
 *                              mask, blurred u8 [B,H,W]; image, reference, out u8 [B,H,W,C] (NULL out = blur only).
 *   sdof_resize_bicubic_u8   : Image.resize((ow, oh)) (default BICUBIC, antialiased) of u8 [B,H,W] -> dst u8 [B,oh,ow]
 *                              and/or latmask f32 [B,4,oh,ow] = around(dst / 255) tiled over the 4 latent channels
 *                              (:304-308).  workspace: sdof_resize_bicubic_workspace_bytes() device bytes, 16-byte aligned.
 *                              Synchronises the stream once (coefficient tables are built on the host like Pillow's). */
int sdof_mask_blur_composite(const uint8_t* mask, const uint8_t* image, const uint8_t* reference, int B, int H, int W, int C,
                             float mask_blur, uint8_t* blurred, uint8_t* out, sdof_stream_t stream);
int64_t sdof_resize_bicubic_workspace_bytes(int B, int H, int W, int oh, int ow);
/* Host-side tables behind the two calls above (no GPU needed; pinned against Pillow by the CPU tests):
 *   sdof_box_blur_params: out3 = {radius, ww, fw} of the extended box for GaussianBlur(mask_blur) (BoxBlur.c);
 *   sdof_resample_table : per output pixel bounds[2*i] = first input pixel, bounds[2*i+1] = tap count, and
 *                         coeffs[i*ksize + k] = 22-bit coefficients (Resample.c::precompute_coeffs + normalize_coeffs_8bpc). */
int sdof_box_blur_params(float mask_blur, int32_t* out3);
int sdof_resample_table(int in_size, int out_size, int32_t* ksize, int32_t* bounds, int32_t* coeffs, int64_t coeffs_cap);
int sdof_resize_bicubic_u8(const uint8_t* src, int B, int H, int W, int oh, int ow, uint8_t* dst, float* latmask, void* workspace,
                           int64_t workspace_bytes, sdof_stream_t stream);

/* ---------------------------------------------------------------- diagnostics */
/* n / d (d >= 1, n < 2^31) evaluated on the host with the multiply-high constants the tiled kernels use for their tile
 * coordinates: lets the CPU tests pin those constants. */
uint32_t sdof_fastdiv_u31(uint32_t n, uint32_t d);
/* Tiles of sdof_warp_cubic_u8's tiled kernel on the current device since the last reset: out[0] = served from the staged
 * shared-memory source rectangle (the fast path the roofline is quoted on), out[1] = per-pixel global-memory fallback
 * (source rectangle of the 32x32 tile larger than the group's 62 KB region: non-smooth flow).  Synchronises the device. */
int sdof_warp_tile_stats(int64_t* out /* host, 2 values */, int reset);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
int64_t sdof_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SDOF_B200_H */
