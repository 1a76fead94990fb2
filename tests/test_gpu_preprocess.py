"""SURVEY 8(f).4 parity on the B200: GPU pre-processing of raw scans vs the CPU oracle restatement of the dataset code
(ground cone rule, fp64 pillar coordinates, BEV / height filtering, stable compaction, collate padding).
Integer / mask / order results are compared bit-exactly."""
import numpy as np
import pytest
import torch

from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.datasets import preprocess_scans
from liso_b200.networks.pcl_to_feature_grid import PointsPillarFeatureNetWrapper
from liso_b200.slim.slim import SLIM
from liso_b200.synth import make_frame_pair
from liso_b200.weights import synth_weights_like
from oracle import slim_oracle as O

pytestmark = pytest.mark.gpu


def _raw_scans(workload, seeds):
    """Synthetic raw scans WITH ground: the generator's frames have their ground removed, so add a ground sheet back."""
    out = []
    for s in seeds:
        p0, _, _ = make_frame_pair(WORKLOADS[workload], s)
        rng = np.random.default_rng(s)
        n_g = p0.shape[0] // 3
        half = WORKLOADS[workload]["bev_range_m"][0] / 2 + 4.0
        g = np.stack([rng.uniform(-half, half, n_g), rng.uniform(-half, half, n_g), rng.normal(-1.73, 0.05, n_g),
                      rng.uniform(0, 1, n_g)], axis=-1).astype(np.float32)
        scan = np.concatenate([p0, g], axis=0)
        out.append(scan[rng.permutation(scan.shape[0])])
    return out


@pytest.mark.parametrize("workload,seeds", [("T", [1, 2, 3]), ("K", [4, 5])])
def test_preprocess_matches_oracle(cuda, workload, seeds):
    cfg = make_cfg(workload)
    scans = _raw_scans(workload, seeds)
    labels = [None, (np.random.default_rng(9).uniform(size=scans[1].shape[0]) < 0.05)] + [None] * (len(scans) - 2)
    got = preprocess_scans([torch.from_numpy(s).to(cuda) for s in scans], cfg,
                           ground_labels=[None if l is None else torch.from_numpy(l).to(cuda) for l in labels])
    torch.cuda.synchronize()
    cap = max(s.shape[0] for s in scans)
    pcl, coors, valid = got["pcl_ta"]["pcl"].cpu().numpy(), got["pcl_ta"]["pillar_coors"].cpu().numpy(), got["pcl_ta"]["pcl_is_valid"].cpu().numpy()
    assert pcl.shape == (len(scans), cap, 4) and coors.shape == (len(scans), cap, 2) and valid.dtype == bool
    for b, s in enumerate(scans):
        _, ref_ta, ref_coors = O.preprocess_scan(s, cfg.data.bev_range_m, cfg.data.img_grid_size, labels[b])
        n = ref_ta.shape[0]
        assert int(got["counts"][b]) == n and 0 < n < s.shape[0]
        assert np.array_equal(pcl[b, :n], ref_ta)            # same points, same (scan) order
        assert np.array_equal(coors[b, :n], ref_coors)
        assert valid[b, :n].all() and not valid[b, n:].any()
        assert np.isnan(pcl[b, n:]).all() and (coors[b, n:] == -1).all()   # collate padding


def test_encoder_on_raw_scan_equals_encoder_on_no_ground_cloud(cuda):
    """The in-kernel ground filter gives exactly the canvas of the compacted pcl_full_no_ground cloud."""
    cfg = make_cfg("N")
    scans = _raw_scans("N", [7, 8])
    torch.manual_seed(0)
    m = PointsPillarFeatureNetWrapper(cfg).to(cuda).eval()
    with torch.no_grad():
        c_raw, o_raw = m([torch.from_numpy(s).to(cuda) for s in scans], raw_scan=True)
        no_ground = [O.preprocess_scan(s, cfg.data.bev_range_m, cfg.data.img_grid_size)[0] for s in scans]
        c_ref, o_ref = m([torch.from_numpy(s).to(cuda) for s in no_ground])
    assert torch.equal(o_raw, o_ref) and torch.equal(c_raw, c_ref)
    assert float(o_raw.sum()) > 100


def test_slim_forward_from_raw_scans(cuda):
    """End to end: raw scans -> preprocess_scans -> SLIM.forward == SLIM.forward on the CPU-prepared sample dicts."""
    cfg = make_cfg("T")
    model = SLIM(cfg).eval()
    model.load_state_dict(synth_weights_like(model.state_dict(), 0))
    model = model.to(cuda)
    raw0, raw1 = _raw_scans("T", [11]), _raw_scans("T", [12])

    def cpu_sample(scans):
        parts = [O.preprocess_scan(s, cfg.data.bev_range_m, cfg.data.img_grid_size) for s in scans]
        pcl = torch.nn.utils.rnn.pad_sequence([torch.from_numpy(p[1]) for p in parts], batch_first=True, padding_value=float("nan"))
        coors = torch.nn.utils.rnn.pad_sequence([torch.from_numpy(p[2]) for p in parts], batch_first=True, padding_value=-1)
        return {"pcl_full_no_ground_ta": [torch.from_numpy(p[0]) for p in parts],
                "pcl_ta": {"pcl": pcl, "pillar_coors": coors, "pcl_is_valid": torch.logical_not(torch.isnan(pcl).sum(-1))}}

    with torch.no_grad():
        a_fw, a_bw = model(preprocess_scans([torch.from_numpy(s).to(cuda) for s in raw0], cfg),
                           preprocess_scans([torch.from_numpy(s).to(cuda) for s in raw1], cfg), None)
        b_fw, b_bw = model(cpu_sample(raw0), cpu_sample(raw1), None)
    for a, b in ((a_fw, b_fw), (a_bw, b_bw)):
        assert torch.equal(a[-1].modified_network_output.static_flow, b[-1].modified_network_output.static_flow)
        n = b[-1].static_flow.shape[1]
        assert torch.equal(a[-1].static_flow[:, :n], b[-1].static_flow)
        assert float(a[-1].static_flow[:, n:].abs().sum()) == 0.0
        assert float((a[-1].static_aggr_trafo - b[-1].static_aggr_trafo).abs().max()) < 1e-9


def test_preprocess_edge_cases(cuda):
    """Empty scan, all-ground scan, single point, NaN point: counts, padding and masks stay consistent."""
    cfg = make_cfg("T")
    rng = np.random.default_rng(3)
    all_ground = np.stack([rng.uniform(-10, 10, 500), rng.uniform(-10, 10, 500), np.full(500, -1.9), rng.uniform(0, 1, 500)], -1).astype(np.float32)
    one = np.array([[1.0, 2.0, 0.1, 0.5]], dtype=np.float32)
    with_nan = np.array([[np.nan, 0.0, 0.0, 0.0], [3.0, -4.0, 0.2, 0.1], [1e9, 0.0, 0.0, 0.0]], dtype=np.float32)
    scans = [np.zeros((0, 4), np.float32), all_ground, one, with_nan]
    got = preprocess_scans([torch.from_numpy(s).to(cuda) for s in scans], cfg)
    torch.cuda.synchronize()
    counts = got["counts"].cpu().numpy()
    for b, s in enumerate(scans):
        _, ref_ta, ref_coors = O.preprocess_scan(s, cfg.data.bev_range_m, cfg.data.img_grid_size)
        assert counts[b] == ref_ta.shape[0]
        n = int(counts[b])
        assert np.array_equal(got["pcl_ta"]["pcl"][b, :n].cpu().numpy(), ref_ta)
        assert np.array_equal(got["pcl_ta"]["pillar_coors"][b, :n].cpu().numpy(), ref_coors)
        assert int(got["pcl_ta"]["pcl_is_valid"][b].sum()) == n
    assert list(counts) == [0, 0, 1, 1]


def test_kernels_against_reference_executed_fixture(cuda):
    """The CUDA kernels directly against what the REFERENCE code answered (tests/golden/preprocess_ref.npz, produced by
    oracle/gen_golden.py from the reference source): a12 pillar coordinates + range / height mask through the raw C ABI,
    and the cone ground rule + compaction through preprocess_scans -- points exactly on cell edges, limits and the cone."""
    import ctypes as C
    import os

    from liso_b200 import _lib
    from liso_b200.config import AttrDict

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_ref.npz"))
    lib = _lib.load()
    for name in ("k", "a", "odd"):
        pts = g[name + "_points"]
        bev, grid = g[name + "_bev_range_m"], g[name + "_img_grid_size"]
        t = torch.from_numpy(pts).to(cuda)
        coors = torch.empty((pts.shape[0], 2), dtype=torch.int32, device=cuda)
        valid = torch.empty((pts.shape[0],), dtype=torch.uint8, device=cuda)
        _lib.check(lib.slimb200_pillar_coors_f64(t.data_ptr(), pts.shape[0], 4, float(np.float32(bev[0])), float(np.float32(bev[1])),
                                                 int(grid[0]), int(grid[1]), -2.0, 1.0, coors.data_ptr(), valid.data_ptr(),
                                                 _lib.current_stream_ptr()))
        torch.cuda.synchronize()
        assert np.array_equal(valid.cpu().numpy().astype(bool), g[name + "_in_range"])
        # (out-of-int32-range garbage of rejected points aside, the integers agree wherever the reference keeps the point)
        keep = g[name + "_in_range"]
        assert np.array_equal(coors.cpu().numpy()[keep], g[name + "_coors"][keep])
        cfg = make_cfg("T")
        cfg.data.bev_range_m, cfg.data.img_grid_size = (float(bev[0]), float(bev[1])), (int(grid[0]), int(grid[1]))
        got = preprocess_scans([t], cfg)
        torch.cuda.synchronize()
        want = keep & ~g[name + "_ground_legacy"]
        n = int(want.sum())
        assert int(got["counts"][0]) == n
        assert np.array_equal(got["pcl_ta"]["pcl"][0, :n].cpu().numpy(), pts[want])
        assert np.array_equal(got["pcl_ta"]["pillar_coors"][0, :n].cpu().numpy(), g[name + "_coors"][want])
