"""Shared constants of the test-suite."""
# the arrays slim_inference_and_save_result writes for a KITTI / nuScenes sample (experiment.py:391-456,459-471)
REFERENCE_TRIPLE_KEYS = {"static_threshold", "bev_range_m"} | {
    "bev_%s_%s" % (kind, d) for kind in ("raw_flow", "dynamicness") for d in ("t0_t1", "t1_t0", "t0_t2", "t2_t0", "t1_t2", "t2_t1")}
