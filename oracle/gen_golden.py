"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the authoring container (needs ``/root/reference``):  ``python -m oracle.gen_golden``
The fixtures travel to the GPU box, the reference does not.  Every array below is produced by
reference code (``oracle/ref_shims.py`` explains the dependency stubs); nothing here calls
``oracle/slim_oracle.py`` or ``liso_b200`` kernels.
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from liso_b200.config import make_cfg
from liso_b200.synth import make_sample_dicts
from liso_b200.weights import synth_weights_like
from oracle import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TINY = dict(bev_range_m=(14.0, 14.0), img_grid_size=(128, 128), n_points=4000, beams=32)
SMALL = dict(bev_range_m=(8.0, 8.0), img_grid_size=(64, 64))


def gen_voxelize(R):
    """Reference numba kernel (voxel_generator.py:76-208) on fp32 clouds incl. both caps."""
    import _ref_voxel_generator as vg  # loaded by ref_shims

    rng = np.random.default_rng(0)
    out = {}
    # case a: over the 20-point cap, under the pillar cap; case b: tiny pillar cap to exercise max_voxels
    for name, n, max_vox in (("a", 6000, 40000), ("b", 6000, 500)):
        pts = rng.uniform(-4.5, 4.5, size=(n, 4)).astype(np.float32)
        pts[:, 2] = rng.uniform(-11, 11, size=n)
        pts[:400, :2] = rng.uniform(1.0, 1.2, size=(400, 2))  # a few very full pillars
        vs = np.array([8.0 / 64, 8.0 / 64, 20.0], dtype=np.float32)
        rg = np.array([-4, -4, -10, 4, 4, 10], dtype=np.float32)
        v, c, k = vg.points_to_voxel(pts, vs, rg, 20, True, max_vox)
        out.update({f"{name}_points": pts, f"{name}_voxel_size": vs, f"{name}_range": rg, f"{name}_max_voxels": max_vox,
                    f"{name}_voxels": v, f"{name}_coors": c, f"{name}_num": k})
    np.savez_compressed(os.path.join(OUT, "voxelize_ref.npz"), **out)


def gen_pillar_encoder(R):
    """Reference PointsPillarFeatureNetWrapper.forward (pcl_to_feature_grid.py:104-107), eval and train BN."""
    cfg = make_cfg("T")
    cfg.data.bev_range_m, cfg.data.img_grid_size = SMALL["bev_range_m"], SMALL["img_grid_size"]
    rng = np.random.default_rng(1)
    clouds = []
    for n in (3000, 1800):
        pts = rng.uniform(-4.6, 4.6, size=(n, 4)).astype(np.float32)
        pts[:, 2] = rng.uniform(-3, 3, size=n)
        pts[:60, :2] = rng.uniform(0.5, 0.62, size=(60, 2))
        clouds.append(pts)
    torch.manual_seed(0)
    ref = R.PointsPillarFeatureNetWrapper(cfg)
    sd = synth_weights_like(ref.state_dict(), 3)
    ref.load_state_dict(sd)
    out = {"bev_range_m": np.array(cfg.data.bev_range_m), "img_grid_size": np.array(cfg.data.img_grid_size),
           "points_0": clouds[0], "points_1": clouds[1]}
    for k, v in sd.items():
        out["w_" + k] = v.numpy()
    for mode in ("eval", "train"):
        ref.load_state_dict(sd)
        ref.train(mode == "train")
        with torch.no_grad():
            voxels, num, coors = ref.voxelize([torch.from_numpy(c) for c in clouds])
            canvas, occ = ref([torch.from_numpy(c) for c in clouds])
        idx = coors.long()
        out[f"{mode}_coors"] = coors.numpy()
        out[f"{mode}_num_points"] = num.numpy()
        out[f"{mode}_pillar_features"] = canvas[idx[:, 0], :, idx[:, 2], idx[:, 3]].numpy()
        out[f"{mode}_canvas_sum"] = np.array(canvas.double().sum().item())
        out[f"{mode}_occupancy_sum"] = np.array(occ.sum().item())
        bn = ref.pts_voxel_encoder.pfn_layers[0].norm
        out[f"{mode}_running_mean"] = bn.running_mean.numpy().copy()
        out[f"{mode}_running_var"] = bn.running_var.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "pillar_encoder_ref.npz"), **out)


def gen_corr(R):
    """Reference CorrBlock (corr.py:6-56) incl. an odd-sized map (floor pooling) and the coords convention."""
    g = torch.Generator().manual_seed(5)
    out = {}
    for name, (B, h, w) in (("even", (1, 16, 16)), ("odd", (2, 9, 13))):
        f1 = torch.randn(B, 128, h, w, generator=g)
        f2 = torch.randn(B, 128, h, w, generator=g)
        coords = R.coords_grid(B, h, w, "cpu") + 2.0 * torch.randn(B, 2, h, w, generator=g)
        coords[:, :, 0, 0] = torch.tensor([-3.0, 1.5])
        coords[:, :, 0, 1] = torch.tensor([w + 4.0, 2.0])
        blk = R.CorrBlock(f1, f2, num_levels=3, radius=3)
        out.update({f"{name}_f1": f1.numpy(), f"{name}_f2": f2.numpy(), f"{name}_coords": coords.numpy(),
                    f"{name}_lookup": blk(coords).numpy()})
        for l, lv in enumerate(blk.corr_pyramid):
            out[f"{name}_level{l}"] = lv.numpy()
    out["coords_grid_10_20"] = R.initialize_flow(torch.zeros(1, 64, 640, 640), 8)[0, :, 10, 20].numpy()  # raft_mod.py:134-135
    np.savez_compressed(os.path.join(OUT, "corr_ref.npz"), **out)


def gen_slim_forward(R):
    """Reference SLIM.forward (slim.py:44-156), eval mode, one tiny pair, deterministic key-seeded weights."""
    cfg = make_cfg("T")
    cfg.data.bev_range_m, cfg.data.img_grid_size = TINY["bev_range_m"], TINY["img_grid_size"]
    model = R.SLIM(cfg, num_train_samples=15000)
    seed = 0
    model.load_state_dict(synth_weights_like(model.state_dict(), seed), strict=True)
    model.eval()
    s0, s1 = make_sample_dicts(TINY, [42])
    summ = {"writer": None, "imgs_eval": False, "metrics_eval": False, "aggregated_metrics": False}
    with torch.no_grad():
        pf, pb = model(copy.deepcopy(s0), copy.deepcopy(s1), summ)
    out = {"bev_range_m": np.array(cfg.data.bev_range_m), "img_grid_size": np.array(cfg.data.img_grid_size),
           "weight_seed": np.array(seed)}
    for t, s in (("t0", s0), ("t1", s1)):
        out["full_" + t] = s["pcl_full_no_ground_ta"][0].numpy()
        out["pcl_" + t] = s["pcl_ta"]["pcl"][0].numpy()
        out["coors_" + t] = s["pcl_ta"]["pillar_coors"][0].numpy()
    for d, p in (("fw", pf), ("bw", pb)):
        out["pt_static_flow_" + d] = p[-1].static_flow[0].numpy()
        out["bev_static_flow_" + d] = p[-1].modified_network_output.static_flow[0].numpy()
        out["bev_dynamicness_" + d] = p[-1].modified_network_output.dynamicness[0].numpy()
        out["static_aggr_trafo_" + d] = p[-1].static_aggr_trafo[0].numpy()
        out["net_out_iter0_sum_" + d] = np.array(p[0].modified_network_output.static_flow.double().abs().sum().item())
    np.savez_compressed(os.path.join(OUT, "slim_forward_tiny.npz"), **out)


def preprocess_points(rng, n, bev_range, grid):
    """Points around a (bev_range, grid) BEV: interior, exactly on cell edges and on the range limits, just outside,
    near the height limits and near the ground cone."""
    half = 0.5 * np.asarray(bev_range, dtype=np.float64)
    pts = rng.uniform(-1.08, 1.08, size=(n, 4)) * np.append(half, [3.0, 1.0])
    cell = np.asarray(bev_range, dtype=np.float64) / np.asarray(grid)
    k = n // 8
    pts[:k, 0] = (rng.integers(0, grid[0] + 1, size=k) * cell[0] - half[0])            # exactly on x cell edges (incl. both limits)
    pts[k:2 * k, 1] = (rng.integers(0, grid[1] + 1, size=k) * cell[1] - half[1])
    pts[2 * k:3 * k, 0] = -half[0] - rng.uniform(0, 1, size=k) * cell[0]                # (-cell, 0): int32 truncation lets them in
    pts[3 * k:4 * k, 2] = rng.choice([-2.0, 1.0, -2.0000002, 0.99999994, -1.9999999], size=k)  # strict height limits
    d = np.hypot(pts[4 * k:5 * k, 0], pts[4 * k:5 * k, 1])
    pts[4 * k:5 * k, 2] = -1.5 + np.tan(0.8 / 180.0 * np.pi) * d + rng.choice([0.0, 1e-7, -1e-7, 1e-4, -1e-4], size=k)  # on the cone
    return pts.astype(np.float32)


def gen_preprocess(R):
    """Reference dataset-side functions executed from their source (ref_shims.ref_preprocess_functions): the a12 point ->
    pillar map (voxelize_sample / voxelize_pcl + height filter) and the cone ground rule under both NumPy promotion rules."""
    import types

    cone_new, voxelize_sample, _ = ref_shims.ref_preprocess_functions(False)
    cone_legacy, _, _ = ref_shims.ref_preprocess_functions(True)
    rng = np.random.default_rng(7)
    out = {}
    for name, bev, grid in (("k", (70.0, 70.0), (640, 640)), ("a", (120.0, 120.0), (920, 920)), ("odd", (70.4, 51.2), (176, 128))):
        pts = preprocess_points(rng, 40000, bev, grid)
        # LidarDataset.__init__ (torch_dataset_commons.py:487-503) / get_bev_setup_params (utils/bev_utils.py:41-43)
        ds = types.SimpleNamespace(bev_range_m_np=np.array(bev, np.float32), img_grid_size_np=np.array(grid).astype(np.int32),
                                   height_range_m_np=np.array((-2.0, 1.0), np.float32))
        coors, in_range = voxelize_sample(ds, pts)
        out.update({name + "_points": pts, name + "_bev_range_m": np.array(bev), name + "_img_grid_size": np.array(grid),
                    name + "_coors": coors.astype(np.int32), name + "_in_range": in_range,
                    name + "_ground_numpy2": cone_new(pts, cone_z_threshold__m=-1.5),
                    name + "_ground_legacy": cone_legacy(pts, cone_z_threshold__m=-1.5)})
    np.savez_compressed(os.path.join(OUT, "preprocess_ref.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    R = ref_shims.ref_modules()
    if "--only-preprocess" in __import__("sys").argv:
        gen_preprocess(R)
        return
    gen_voxelize(R)
    gen_pillar_encoder(R)
    gen_corr(R)
    gen_slim_forward(R)
    gen_preprocess(R)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
