"""CPU oracle for the flow -> warp -> mask/composite hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `sd_animation_optical_flow_b200/` may
import this package; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it, and there only
as the checker or the CPU baseline, never as the product path.

Each function is a NumPy restatement of one reference function and cites the
reference file:line it follows.  The oracle is pinned (tests/test_oracle_*.py,
`-m "not gpu"`) against

  * the reference's own Python (`RAFT/core/corr.py`, `RAFT/core/raft.py`)
    imported from /root/reference in the authoring container -> committed
    fixtures under tests/golden/ (generator: oracle/make_golden.py), and
  * OpenCV itself (cv2.remap / cv2.dilate / cv2.Laplacian / cv2.cvtColor),
    which is what the reference calls for the warp and mask steps.

PDCNet+ (`estimate_flow_and_confidence_map`) is third-party code that is not in
the reference tree and not pinned by it: parity for that network is UNPINNED;
only its post-processing (softmax confidence, masks, composite) is covered.
"""
