// Internal declarations shared by the correlation translation units.
#pragma once

#include "sdof_common.cuh"

namespace sdof {

// corr_simt.cu
int launch_corr_volume_fp32(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, float* pyramid,
                            const sdof_pyramid_layout& lay, cudaStream_t st);
int launch_pool_levels(float* pyramid, const sdof_pyramid_layout& lay, int64_t rows, int first_level, cudaStream_t st);

// corr_tc.cu: tcgen05 volume + fused pyramid.  Returns SDOF_ERR_UNSUPPORTED (without touching the
// error text's severity) when the shape cannot go through TMA so the caller can fall back.
int launch_corr_volume_tc(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, int precision,
                          float* pyramid, const sdof_pyramid_layout& lay, void* workspace, int64_t workspace_bytes,
                          cudaStream_t st);
int64_t corr_tc_workspace_bytes(int B, int n1, int n2, int C, int precision);

// corr_tc_res.cu: resident-operand tcgen05 kernel, 16-bit operands (fmt 0 = fp16, 1 = bf16), pooled-feature pyramid
int launch_corr_pyramid_resident(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, int fmt,
                                 float* pyramid, const sdof_pyramid_layout& lay, void* workspace, int64_t workspace_bytes,
                                 cudaStream_t st);
int launch_corr_prepare_resident(const float* fmap1, const float* fmap2, int B, int n1, int h2, int w2, int C, int fmt,
                                 const sdof_pyramid_layout& lay, void* workspace, int64_t workspace_bytes, int part,
                                 cudaStream_t st);
int launch_corr_pyramid_prepared(int B, int n1, int h2, int w2, int C, int fmt, float* pyramid, const sdof_pyramid_layout& lay,
                                 void* workspace, int64_t workspace_bytes, cudaStream_t st);
int64_t corr_res_workspace_bytes(int B, int n1, int h2, int w2, int C, int levels);
// round 2: split operand buffers (source / target, each with a scale header), shared key-frame target, fp16 pyramid
int64_t corr_res_src_bytes(int B, int n1, int C);
int64_t corr_res_tgt_bytes(int B2, int h2, int w2, int C, int levels);
int launch_corr_prepare_parts(const float* fmap1, int B, int n1, void* src_ops, const float* fmap2, int B2, int h2, int w2,
                              void* tgt_ops, int C, int levels, int fmt, cudaStream_t st);
int launch_corr_pyramid_parts(const void* src_ops, const void* tgt_ops, int B, int n1, int B2, int h2, int w2, int C, int fmt,
                              int out_half, void* pyramid, const sdof_pyramid_layout& lay, cudaStream_t st);
bool corr_res_supported(int C, int levels);

}  // namespace sdof
