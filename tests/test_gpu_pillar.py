"""Stage-1 parity on the B200: CUDA pillar encoder (through the C ABI) vs the CPU oracle.

Bars (BASELINE.json north star): pillar indices and point->pillar assignment bit-exact;
PFN / canvas features within 1e-5 relative in fp32; occupancy exact.
"""
import numpy as np
import pytest
import torch

from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.networks.pcl_to_feature_grid import PointsPillarFeatureNetWrapper
from liso_b200.synth import make_frame_pair
from oracle import slim_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # stated tolerance: |a - b| <= RTOL * max(1, |b|_inf-scale)


FORMAT = {"fmt": "channels_last"}


@pytest.fixture(autouse=True, params=["channels_last", "contiguous"])
def canvas_format(request):
    """Every test runs for both canvas layouts: NHWC (k_pillar_nhwc, product default) and NCHW (k_tile_encode)."""
    FORMAT["fmt"] = request.param
    yield request.param


def _module(cfg, device, seed=0, training=False):
    torch.manual_seed(seed)
    m = PointsPillarFeatureNetWrapper(cfg, canvas_memory_format=FORMAT["fmt"])
    bn = m.pts_voxel_encoder.pfn_layers[0].norm
    with torch.no_grad():  # randomised BN so the zero-row value relu(beta - mu*gamma/sigma) is non-trivial
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
        bn.running_mean.uniform_(-0.3, 0.3)
        bn.running_var.uniform_(0.5, 2.0)
    m.train(training)
    return m.to(device)


def _params(m):
    pfn = m.pts_voxel_encoder.pfn_layers[0]
    return dict(linear_weight=pfn.linear.weight.detach().cpu(), bn_weight=pfn.norm.weight.detach().cpu(),
                bn_bias=pfn.norm.bias.detach().cpu(), running_mean=pfn.norm.running_mean.detach().cpu().clone(),
                running_var=pfn.norm.running_var.detach().cpu().clone())


def _check(cfg, clouds, device, training=False, c_in=4):
    m = _module(cfg, device, training=training)
    params = _params(m)
    ref = O.pillar_encoder_forward(clouds, params, cfg.data.bev_range_m, cfg.data.img_grid_size,
                                   cfg.data.z_pillar_cutoff_value, training)
    pc_range, voxel_size = O.pillar_geometry(cfg.data.bev_range_m, cfg.data.img_grid_size, cfg.data.z_pillar_cutoff_value)
    ref_vox, ref_num, ref_coors, ref_p2p = O.voxelize_batch(clouds, pc_range, voxel_size)
    with torch.no_grad():
        got = m.voxelize_debug([torch.from_numpy(c).to(device) for c in clouds])
    torch.cuda.synchronize()
    # --- integer work: bit-exact ---------------------------------------------------------
    assert got["coors"].shape[0] == ref_coors.shape[0], (got["coors"].shape, ref_coors.shape)
    assert np.array_equal(got["coors"].cpu().numpy(), ref_coors)
    assert np.array_equal(got["num_points"].cpu().numpy(), ref_num)
    assert np.array_equal(got["pt2pillar"].cpu().numpy(), np.concatenate(ref_p2p))
    assert np.array_equal(got["voxels"].cpu().numpy(), ref_vox)  # same points, same slot order
    assert torch.equal(got["occupancy"].cpu(), ref["occupancy"])
    # --- fp32 features: 1e-5 relative -----------------------------------------------------
    canvas = got["canvas"].cpu()
    want_fmt = torch.channels_last if FORMAT["fmt"] == "channels_last" else torch.contiguous_format
    assert got["canvas"].is_contiguous(memory_format=want_fmt) and canvas.shape == ref["canvas"].shape
    scale = max(1.0, float(ref["canvas"].abs().max()))
    err = float((canvas - ref["canvas"]).abs().max())
    assert err <= RTOL * scale, (err, scale)
    assert torch.equal(canvas != 0, ref["canvas"] != 0) or err <= RTOL * scale
    if training:
        bn = m.pts_voxel_encoder.pfn_layers[0].norm
        assert torch.allclose(bn.running_mean.cpu(), ref["running_mean"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(bn.running_var.cpu(), ref["running_var"], rtol=1e-5, atol=1e-6)
        assert int(bn.num_batches_tracked) == 1
    return got, ref


@pytest.mark.parametrize("workload", ["T", "N", "K", "A"])
def test_synthetic_lidar_eval(cuda, workload):
    cfg = make_cfg(workload)
    p0, p1, _ = make_frame_pair(WORKLOADS[workload], 5)
    _check(cfg, [p0, p1], cuda)


@pytest.mark.parametrize("workload", ["T", "K"])
def test_synthetic_lidar_train_mode_bn(cuda, workload):
    """Q4: the reference export runs BatchNorm1d in train mode (batch statistics incl. padded rows)."""
    cfg = make_cfg(workload)
    p0, p1, _ = make_frame_pair(WORKLOADS[workload], 6)
    _check(cfg, [p0, p1], cuda, training=True)


def test_pillar_cap_40000(cuda):
    """Uniform-random points occupy > 40000 pillars: later first-appearances are dropped."""
    cfg = make_cfg("K")
    rng = np.random.default_rng(0)
    pts = rng.uniform(-36.0, 36.0, size=(120_000, 4)).astype(np.float32)
    pts[:, 2] = rng.uniform(-3, 3, size=pts.shape[0])
    got, _ = _check(cfg, [pts], cuda)
    assert got["coors"].shape[0] == 40000


def test_heavy_cells_and_edges(cuda):
    """> 20, > 32 and > 1000 points in one pillar; points on/just outside the grid borders; z borders."""
    cfg = make_cfg("T")
    rng = np.random.default_rng(1)
    half = cfg.data.bev_range_m[0] / 2
    vs = cfg.data.bev_range_m[0] / cfg.data.img_grid_size[0]
    blobs = []
    for n, c in ((25, (1.0, 2.0)), (33, (-3.0, 4.0)), (70, (5.5, -5.5)), (1500, (-7.0, -7.0)), (20, (0.05, 0.05)), (21, (9.0, 9.0))):
        b = np.zeros((n, 4), np.float32)
        b[:, 0] = c[0] + rng.uniform(0, vs * 0.9, n)
        b[:, 1] = c[1] + rng.uniform(0, vs * 0.9, n)
        b[:, 2] = rng.uniform(-2, 2, n)
        b[:, 3] = rng.uniform(0, 1, n)
        blobs.append(b)
    edge = np.array([[-half, 0, 0, 1], [half, 0, 0, 1], [np.nextafter(np.float32(half), np.float32(0)), 0, 0, 1],
                     [0, -half, 0, 1], [0, half, 0, 1], [-half - 1e-3, 0, 0, 1], [100, 100, 0, 1],
                     [0, 0, -10, 1], [0, 0, 10, 1], [0, 0, 9.999, 1], [0, 0, -10.001, 1],
                     [vs, vs, 0, 1], [2 * vs, 3 * vs, 0, 1], [-vs, -vs, 0, 1]], dtype=np.float32)
    scatter = rng.uniform(-half - 2, half + 2, size=(5000, 4)).astype(np.float32)
    pts = np.concatenate(blobs + [edge, scatter], axis=0)
    pts = pts[rng.permutation(pts.shape[0])]
    _check(cfg, [pts, pts[::-1].copy()], cuda)


def test_empty_and_ragged_batch(cuda):
    cfg = make_cfg("T")
    rng = np.random.default_rng(2)
    a = rng.uniform(-15, 15, size=(777, 4)).astype(np.float32)
    empty = np.zeros((0, 4), np.float32)
    outside = np.full((10, 4), 500.0, np.float32)
    got, _ = _check(cfg, [a, empty, outside, a[:1]], cuda)
    assert float(got["canvas"][1].abs().max()) == 0.0 and float(got["occupancy"][2].abs().max()) == 0.0


def test_three_channel_cloud(cuda):
    cfg = make_cfg("T")
    cfg.data.use_lidar_intensity = False
    p0, _, _ = make_frame_pair(WORKLOADS["T"], 7)
    m = _module(cfg, cuda)
    clouds = [np.ascontiguousarray(p0[:, :3])]
    params = _params(m)
    # oracle PFN with 9 input channels
    ref = O.pillar_encoder_forward(clouds, params, cfg.data.bev_range_m, cfg.data.img_grid_size, 10.0, False)
    with torch.no_grad():
        canvas, occ = m([torch.from_numpy(clouds[0]).to(cuda)])
    assert torch.equal(occ.cpu(), ref["occupancy"])
    assert float((canvas.cpu() - ref["canvas"]).abs().max()) <= RTOL * max(1.0, float(ref["canvas"].abs().max()))


def test_forward_matches_voxelize_debug_and_is_deterministic(cuda):
    cfg = make_cfg("N")
    p0, p1, _ = make_frame_pair(WORKLOADS["N"], 8)
    m = _module(cfg, cuda)
    pts = [torch.from_numpy(p0).to(cuda), torch.from_numpy(p1).to(cuda)]
    with torch.no_grad():
        c1, o1 = m(pts)
        c2, o2 = m(pts)
    assert torch.equal(c1, c2) and torch.equal(o1, o2)  # run-to-run bit-identical (no float atomics)


def test_idempotence_property_full_size(cuda):
    """Size-independent property at the bench size (B=8 K frames): occupied canvas cells == occupancy,
    per-sample results do not depend on batch composition (eval mode)."""
    cfg = make_cfg("K")
    frames = []
    for s in range(4):
        p0, p1, _ = make_frame_pair(WORKLOADS["K"], 20 + s)
        frames += [p0, p1]
    m = _module(cfg, cuda)
    pts = [torch.from_numpy(f).to(cuda) for f in frames]
    with torch.no_grad():
        canvas, occ = m(pts)
        single, occ_s = m(pts[3:4])
    assert torch.equal(canvas[3], single[0]) and torch.equal(occ[3], occ_s[0])
    assert torch.equal((canvas.abs().sum(dim=1, keepdim=True) > 0) | (occ > 0), occ > 0)


def test_dataset_pillar_coors_f64(cuda):
    """a12: fp64 trunc arithmetic of voxelize_pcl, incl. the (-voxel, 0) -> 0 trunc-toward-zero case."""
    import ctypes as C

    from liso_b200 import _lib

    for workload in ("K", "A"):
        W = WORKLOADS[workload]
        rng = np.random.default_rng(3)
        half = W["bev_range_m"][0] / 2
        vs = W["bev_range_m"][0] / W["img_grid_size"][0]
        pts = rng.uniform(-half - 1, half + 1, size=(200_000, 4)).astype(np.float32)
        pts[:, 2] = rng.uniform(-2.5, 1.5, size=pts.shape[0])
        pts[:8, 0] = [-half - vs / 2, -half, -half + 1e-6, half, half - 1e-6, 0, vs, -vs]
        pts[:8, 2] = [0, 0, 0, 0, 0, -2.0, 1.0, 0.999]
        ref_c, ref_ok = O.pillar_coors_f64(pts, W["bev_range_m"], W["img_grid_size"])
        t = torch.from_numpy(pts).to(cuda)
        coors = torch.empty((pts.shape[0], 2), dtype=torch.int32, device=cuda)
        valid = torch.empty((pts.shape[0],), dtype=torch.uint8, device=cuda)
        rc = _lib.load().slimb200_pillar_coors_f64(t.data_ptr(), pts.shape[0], 4, W["bev_range_m"][0], W["bev_range_m"][1],
                                                   W["img_grid_size"][0], W["img_grid_size"][1], -2.0, 1.0,
                                                   coors.data_ptr(), valid.data_ptr(), _lib.current_stream_ptr())
        _lib.check(rc)
        assert np.array_equal(valid.cpu().numpy().astype(bool), ref_ok)
        assert np.array_equal(coors.cpu().numpy()[ref_ok], ref_c[ref_ok])
