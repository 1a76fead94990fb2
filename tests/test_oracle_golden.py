"""CPU: the oracle restatement vs (a) the reference's own golden vector / pinned comment and
(b) fixtures produced by running the unmodified reference (oracle/gen_golden.py)."""
import os

import numpy as np
import torch

from liso_b200.config import make_cfg
from liso_b200.weights import synth_weights_like
from oracle import slim_forward as SF
from oracle import slim_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_voxel_generator_golden_vector():
    """mmdetection3d/tests/test_models/test_voxel_encoder/test_voxel_generator.py:8-24 (float64 cloud)."""
    np.random.seed(0)
    points = np.random.rand(1000, 4)
    expected_coors = np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1], [7, 81, 0], [6, 80, 1], [7, 80, 0], [6, 80, 0]])
    expected_num = np.array([120, 121, 127, 134, 115, 127, 125, 131])
    for fn in (O.hard_voxelize, O.hard_voxelize_loop):
        voxels, coors, num, p2v = fn(points, [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
        assert voxels.shape == (8, 1000, 4)
        assert np.all(coors == expected_coors)
        assert np.all(num == expected_num)
        assert (p2v >= 0).all()


def test_initialize_flow_convention():
    """raft_mod.py:134-135: pixel_coords_t0[0, ..., 10, 20] -> tensor([20., 10.])"""
    assert O.coords_grid(1, 80, 80)[0, :, 10, 20].tolist() == [20.0, 10.0]
    g = np.load(os.path.join(GOLDEN, "corr_ref.npz"))
    assert g["coords_grid_10_20"].tolist() == [20.0, 10.0]


def test_voxelize_vs_reference_numba_kernel():
    g = np.load(os.path.join(GOLDEN, "voxelize_ref.npz"))
    for name in ("a", "b"):
        args = (g[name + "_points"], g[name + "_voxel_size"], g[name + "_range"], 20, int(g[name + "_max_voxels"]))
        v, c, n, p2v = O.hard_voxelize(*args)
        assert np.array_equal(c, g[name + "_coors"]) and np.array_equal(n, g[name + "_num"])
        assert np.array_equal(v, g[name + "_voxels"])
        # vectorised and line-by-line restatements agree, incl. the point -> pillar map
        v2, c2, n2, p2v2 = O.hard_voxelize_loop(*args)
        assert np.array_equal(c, c2) and np.array_equal(n, n2) and np.array_equal(v, v2) and np.array_equal(p2v, p2v2)
    assert int(g["b_coors"].shape[0]) == 500  # pillar cap reached


def test_pillar_encoder_vs_reference_module():
    g = np.load(os.path.join(GOLDEN, "pillar_encoder_ref.npz"))
    pp = "pts_voxel_encoder.pfn_layers.0."
    params = dict(linear_weight=torch.from_numpy(g["w_" + pp + "linear.weight"]), bn_weight=torch.from_numpy(g["w_" + pp + "norm.weight"]),
                  bn_bias=torch.from_numpy(g["w_" + pp + "norm.bias"]), running_mean=torch.from_numpy(g["w_" + pp + "norm.running_mean"]),
                  running_var=torch.from_numpy(g["w_" + pp + "norm.running_var"]))
    clouds = [g["points_0"], g["points_1"]]
    for mode in ("eval", "train"):
        out = O.pillar_encoder_forward(clouds, params, g["bev_range_m"], g["img_grid_size"], 10.0, mode == "train")
        assert np.array_equal(out["coors"].numpy(), g[mode + "_coors"])
        assert np.array_equal(out["num_points"].numpy(), g[mode + "_num_points"])
        np.testing.assert_allclose(out["pillar_features"].numpy(), g[mode + "_pillar_features"], rtol=1e-6, atol=1e-6)
        assert abs(float(out["canvas"].double().sum()) - float(g[mode + "_canvas_sum"])) < 1e-3 * max(1.0, abs(float(g[mode + "_canvas_sum"])))
        assert float(out["occupancy"].sum()) == float(g[mode + "_occupancy_sum"])
        if mode == "train":  # the reference ran voxelize + forward = one BN update in forward only
            np.testing.assert_allclose(out["running_mean"].numpy(), g["train_running_mean"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(out["running_var"].numpy(), g["train_running_var"], rtol=1e-5, atol=1e-6)


def test_corr_pyramid_and_lookup_vs_reference():
    g = np.load(os.path.join(GOLDEN, "corr_ref.npz"))
    for name in ("even", "odd"):
        f1, f2 = torch.from_numpy(g[name + "_f1"]), torch.from_numpy(g[name + "_f2"])
        pyr = O.corr_pyramid(f1, f2, 3)
        for l, lv in enumerate(pyr):
            np.testing.assert_allclose(lv.numpy(), g["%s_level%d" % (name, l)], rtol=1e-6, atol=1e-6)
        out = O.corr_lookup(pyr, torch.from_numpy(g[name + "_coords"]), 3)
        np.testing.assert_allclose(out.numpy(), g[name + "_lookup"], rtol=1e-5, atol=1e-5)


def test_slim_forward_port_vs_reference():
    g = np.load(os.path.join(GOLDEN, "slim_forward_tiny.npz"))
    cfg = make_cfg("T")
    cfg.data.img_grid_size = tuple(int(v) for v in g["img_grid_size"])
    cfg.data.bev_range_m = tuple(float(v) for v in g["bev_range_m"])
    from liso_b200.slim.slim import SLIM

    sd = synth_weights_like(SLIM(cfg).state_dict(), int(g["weight_seed"]))

    def sample(t):
        pcl = torch.from_numpy(g["pcl_" + t])
        return {"pcl_full_no_ground_ta": [torch.from_numpy(g["full_" + t])],
                "pcl_ta": {"pcl": pcl[None], "pcl_is_valid": torch.ones(1, pcl.shape[0], dtype=torch.bool),
                           "pillar_coors": torch.from_numpy(g["coors_" + t])[None]}}

    with torch.no_grad():
        of, ob, _ = SF.slim_forward(sd, cfg, sample("t0"), sample("t1"))
    for d, o in (("fw", of), ("bw", ob)):
        np.testing.assert_allclose(o[-1]["static_flow"][0].numpy(), g["bev_static_flow_" + d], rtol=0, atol=2e-5)
        np.testing.assert_allclose(o[-1]["pointwise_static_flow"][0].numpy(), g["pt_static_flow_" + d], rtol=0, atol=2e-5)
        np.testing.assert_allclose(o[-1]["dynamicness"][0].numpy(), g["bev_dynamicness_" + d], rtol=0, atol=2e-5)
        np.testing.assert_allclose(o[-1]["static_aggr_trafo"][0].numpy(), g["static_aggr_trafo_" + d], rtol=0, atol=1e-4)


def test_dataset_side_preprocessing_vs_reference_source():
    """a12 + SURVEY 8f.4 pins: ``O.pillar_coors_f64`` against the reference's ``voxelize_sample`` -> ``voxelize_pcl``
    (``torch_dataset_commons.py:975-987``, ``analyse_boxes.py:6-26``) and ``O.ground_label_cone_f32`` against
    ``infer_ground_label_using_cone`` (``:133-146``), both executed from the reference source by ``oracle/gen_golden.py``
    on points that sit exactly on cell edges, range limits, height limits and on the cone."""
    g = np.load(os.path.join(GOLDEN, "preprocess_ref.npz"))
    for name in ("k", "a", "odd"):
        pts = g[name + "_points"]
        coors, ok = O.pillar_coors_f64(pts, g[name + "_bev_range_m"], g[name + "_img_grid_size"])
        assert np.array_equal(coors, g[name + "_coors"]) and coors.dtype == np.int32
        assert np.array_equal(ok, g[name + "_in_range"])
        assert 0.3 < ok.mean() < 0.7  # both outcomes well represented
        ground = O.ground_label_cone_f32(pts, -1.5)
        # the reference environment is NumPy 1.18.5 (docker/Dockerfile.base:53): float32 arithmetic throughout
        assert np.array_equal(ground, g[name + "_ground_legacy"])
        # NumPy >= 2 promotes the product to float64: only points within rounding of the cone may flip
        flips = ground != g[name + "_ground_numpy2"]
        d = np.hypot(pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64))
        margin = np.abs(pts[:, 2].astype(np.float64) - (-1.5 + np.tan(0.8 / 180.0 * np.pi) * d))
        assert flips.sum() > 0 and margin[flips].max() < 1e-5
