"""B200-native flow -> warp -> mask/composite hot path of zyddnys/sd_animation_optical_flow.

Layout (only what the path needs):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/sdof_b200.h)
  _capi.py         ctypes binding of libsdof_b200.so (no fallback: raises if missing)
  ops.py           device-resident operators over CUDA tensors
  corr.py          CorrBlock / AlternateCorrBlock with the reference's corr_fn protocol
  alt_cuda_corr.py drop-in for the reference's pybind module
  raft.py          RAFT network (PyTorch convs; same state-dict keys as the reference)
  engine.py        RaftEngine.estimate_flow / warp on the device, CUDA-graph capture
  pdcnet_of.py     drop-in for the reference's pdcnet_of.py (warp_frame, create_of_algo, ...)
  ofgen.py         drop-in for the flow/warp/mask helpers of the ofgen_*.py scripts
  shard.py         frame pairs sharded across the GPUs of one box
"""
__version__ = '0.1.0'
