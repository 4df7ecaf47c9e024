#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over the small-size GPU tests of the kernels written this round
mkdir -p gpurun_out
SEL='test_cubic_u8_batched_and_shared_source or test_unaligned_source_pointer or test_numpy_api_matches_reference_golden or test_instnorm_nhwc_kernels or test_small_conv_kernels or (test_gaussian_blur_bit_exact and (61 or 7-5)) or (test_resize_bicubic_bit_exact and 100) or (test_detect_edges_bit_exact_vs_opencv and 61) or test_noise_image_and_swapped_thresholds'
for tool in memcheck racecheck; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_warp.py tests/test_gpu_raft.py tests/test_gpu_blur.py tests/test_gpu_keyframe.py -m gpu -q -x --timeout 1400 -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?"; grep -c "ERROR SUMMARY" gpurun_out/sanitizer_$tool.log; grep "ERROR SUMMARY\|passed\|failed\|Invalid\|Race reported\|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -12
done
