"""CPU, authoring container only: the oracle restatement against the LIVE unmodified reference at larger
sizes than the committed fixtures.  Skipped where /root/reference does not exist (e.g. the GPU box)."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present")


def test_state_dict_keys_equal_reference():
    from liso_b200.config import make_cfg
    from liso_b200.slim.slim import SLIM

    R = ref_shims.ref_modules()
    cfg = make_cfg("T")
    ref = R.SLIM(cfg, num_train_samples=15000)
    mine = SLIM(cfg)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert mine.state_dict()[k].shape == v.shape and mine.state_dict()[k].dtype == v.dtype, k
    mine.load_state_dict(ref.state_dict(), strict=True)  # T10
    ref.load_state_dict(mine.state_dict(), strict=True)


def test_oracle_forward_equals_reference_forward():
    from liso_b200.config import WORKLOADS, make_cfg
    from liso_b200.synth import make_sample_dicts
    from liso_b200.weights import synth_weights_like
    from oracle import slim_forward as SF

    R = ref_shims.ref_modules()
    cfg = make_cfg("T")
    ref = R.SLIM(cfg, num_train_samples=15000)
    sd = synth_weights_like(ref.state_dict(), 1)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    s0, s1 = make_sample_dicts(WORKLOADS["T"], [5, 6])
    summ = {"writer": None, "imgs_eval": False, "metrics_eval": False, "aggregated_metrics": False}
    with torch.no_grad():
        pf, pb = ref(copy.deepcopy(s0), copy.deepcopy(s1), summ)
        of, ob, _ = SF.slim_forward(sd, cfg, s0, s1)
    for it in range(6):
        for p, o in ((pf, of), (pb, ob)):
            assert torch.equal(p[it].modified_network_output.static_flow, o[it]["static_flow"])
            assert torch.equal(p[it].modified_network_output.dynamicness, o[it]["dynamicness"])
            assert torch.equal(p[it].static_flow, o[it]["pointwise_static_flow"])
            assert torch.allclose(p[it].modified_network_output.static_aggr_flow, o[it]["static_aggr_flow"], atol=1e-6)


def test_oracle_pillar_encoder_equals_reference_kitti_size():
    from liso_b200.config import WORKLOADS, make_cfg
    from liso_b200.synth import make_frame_pair
    from oracle import slim_oracle as O

    R = ref_shims.ref_modules()
    cfg = make_cfg("K")
    torch.manual_seed(0)
    ref = R.PointsPillarFeatureNetWrapper(cfg)
    bn = ref.pts_voxel_encoder.pfn_layers[0].norm
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5), bn.bias.uniform_(-0.5, 0.5), bn.running_mean.uniform_(-0.3, 0.3), bn.running_var.uniform_(0.5, 2)
    p0, p1, _ = make_frame_pair(WORKLOADS["K"], 1)
    for training in (False, True):
        params = dict(linear_weight=ref.pts_voxel_encoder.pfn_layers[0].linear.weight.detach(), bn_weight=bn.weight.detach(),
                      bn_bias=bn.bias.detach(), running_mean=bn.running_mean.clone(), running_var=bn.running_var.clone())
        ref.train(training)
        with torch.no_grad():
            canvas, occ = ref([torch.from_numpy(p0), torch.from_numpy(p1)])
        out = O.pillar_encoder_forward([p0, p1], params, cfg.data.bev_range_m, cfg.data.img_grid_size, 10.0, training)
        assert torch.equal(canvas, out["canvas"]) and torch.equal(occ, out["occupancy"])
        assert torch.equal(bn.running_mean, out["running_mean"])


def test_dataset_side_preprocessing_live():
    """Larger live run of the pins of test_oracle_golden.py::test_dataset_side_preprocessing_vs_reference_source."""
    import types

    from oracle import slim_oracle as O
    from oracle.gen_golden import preprocess_points

    cone_legacy, voxelize_sample, voxelize_pcl = ref_shims.ref_preprocess_functions(True)
    rng = np.random.default_rng(11)
    for bev, grid in (((70.0, 70.0), (640, 640)), ((120.0, 120.0), (920, 920)), ((51.2, 70.4), (128, 176))):
        pts = preprocess_points(rng, 300000, bev, grid)
        ds = types.SimpleNamespace(bev_range_m_np=np.array(bev, np.float32), img_grid_size_np=np.array(grid).astype(np.int32),
                                   height_range_m_np=np.array((-2.0, 1.0), np.float32))
        ref_c, ref_ok = voxelize_sample(ds, pts)
        c, ok = O.pillar_coors_f64(pts, bev, grid)
        assert np.array_equal(c, ref_c) and np.array_equal(ok, ref_ok)
        # torch tensors take the other branch of voxelize_pcl (analyse_boxes.py:10-13): same integers
        tc, tok = voxelize_pcl(torch.from_numpy(pts), torch.from_numpy(np.append(ds.bev_range_m_np, np.array(1000.0))),
                               torch.from_numpy(np.append(ds.img_grid_size_np, np.array(1))))
        assert np.array_equal(tc.numpy()[:, :2], c)
        assert np.array_equal(O.ground_label_cone_f32(pts, -1.5), cone_legacy(pts, cone_z_threshold__m=-1.5))


def test_export_schema_keys_match_reference_source():
    """The arrays the reference export writes per sample, read off its source (experiment.py:391-456,459-471): every
    `preds["..."]` / `save_stuff[...]` key + bev_range_m == what AsyncNpzWriter writes for a triple, and the pair subset."""
    import os
    import re

    from liso_b200.slim import export
    from tests_keys import REFERENCE_TRIPLE_KEYS

    src = open(os.path.join(ref_shims.REF_ROOT, "liso/slim/experiment.py")).read()
    body = src[src.index("def slim_inference_and_save_result"):src.index("def run(", src.index("def slim_inference_and_save_result"))]
    keys = set(re.findall(r'preds\["([a-z0-9_]+)"\]', body)) | set(re.findall(r'save_stuff_cpu\["([a-z0-9_]+)"\]', body))
    keys |= set(re.findall(r'"(static_threshold)":', body))
    assert keys == REFERENCE_TRIPLE_KEYS
    assert set(export.export_keys(6)) | {"static_threshold", "bev_range_m"} == keys
    assert set(export.export_keys(2)) < keys
