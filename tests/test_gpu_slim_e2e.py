"""End-to-end parity on the B200: SLIM forward through the CUDA hot path vs the CPU oracle port of the
reference forward.  Bar: final per-point flow within 1 cm average end-point error (AEE as in
``liso/slim/utils/metrics.py:113-121``); BEV masks identical."""
import os

import numpy as np
import pytest
import torch

from liso_b200 import _lib
from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.slim.slim import SLIM
from liso_b200.synth import make_sample_dicts
from liso_b200.weights import synth_weights_like
from oracle import slim_forward as SF

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _model(cfg, device, seed=0, **kw):
    m = SLIM(cfg, **kw).eval()
    sd = synth_weights_like(m.state_dict(), seed)
    m.load_state_dict(sd, strict=True)
    return m.to(device), sd


def _aee(pred, ref, valid):
    return float((pred - ref).norm(dim=-1)[valid].mean())


@pytest.mark.parametrize("fused_lookup", [False, "always"])
@pytest.mark.parametrize("workload,seeds", [("T", [3, 4]), ("N", [5]), ("K", [6]), ("A", [7])])
def test_slim_forward_flow_within_1cm(cuda, workload, seeds, fused_lookup):
    """`fused_lookup`: the lookup + conv_stat_corr1 + ReLU as ONE tcgen05 kernel with tf32 operands instead of
    the lookup kernel + the stock fp32 convolution; same 1 cm bar."""
    cfg = make_cfg(workload)
    model, sd = _model(cfg, cuda)
    model.raft_network.fuse_lookup_conv = fused_lookup
    n_fused = _lib.load().slimb200_launch_count(_lib.K_LOOKUP_CONV)
    s0, s1 = make_sample_dicts(WORKLOADS[workload], seeds)
    with torch.no_grad():
        pf, pb = model(s0, s1, None)
        torch.cuda.synchronize()
        of, ob, _ = SF.slim_forward(sd, cfg, s0, s1, decode_all_iterations=False)
    for p, o, s in ((pf, of, s0), (pb, ob, s1)):
        valid = s["pcl_ta"]["pcl_is_valid"]
        aee = _aee(p[-1].static_flow.cpu(), o[-1]["pointwise_static_flow"], valid)
        bev = float((p[-1].modified_network_output.static_flow.cpu() - o[-1]["static_flow"]).abs().max())
        dyn = float((p[-1].modified_network_output.dynamicness.cpu() - o[-1]["dynamicness"]).abs().max())
        mag = float(o[-1]["pointwise_static_flow"].norm(dim=-1)[valid].mean())
        print("%s: AEE %.3e m, BEV max |d| %.3e m, dynamicness max |d| %.3e, mean |flow| %.3f m" % (workload, aee, bev, dyn, mag))
        assert aee <= 0.01
        assert torch.equal(p[-1].modified_network_output.static_flow.cpu() != 0, o[-1]["static_flow"] != 0)
        assert dyn < 5e-2
    # the fused kernel really ran (12 launches: 6 iterations x 2 directions, captured once) / really did not
    assert (_lib.load().slimb200_launch_count(_lib.K_LOOKUP_CONV) - n_fused > 0) == bool(fused_lookup)


def test_batch_of_8_kitti_pairs_every_sample_within_1cm(cuda):
    """The bench batch (configs[1]: 8 KITTI-sized pairs in one forward): EVERY sample of the batch against the oracle run
    on that sample alone (eval mode: samples are independent), fused lookup on."""
    cfg = make_cfg("K")
    model, sd = _model(cfg, cuda, decode_iterations="last", static_aggregation=False)
    model.raft_network.fuse_lookup_conv = "always"
    seeds = [1000 + i for i in range(8)]
    s0, s1 = make_sample_dicts(WORKLOADS["K"], seeds)
    with torch.no_grad():
        pf, pb = model(s0, s1, None)
    torch.cuda.synchronize()
    got = [(pf[-1].static_flow.cpu(), pf[-1].modified_network_output.static_flow.cpu()),
           (pb[-1].static_flow.cpu(), pb[-1].modified_network_output.static_flow.cpu())]
    for b, seed in enumerate(seeds):
        o0, o1 = make_sample_dicts(WORKLOADS["K"], [seed])
        with torch.no_grad():
            of, ob, _ = SF.slim_forward(sd, cfg, o0, o1, decode_all_iterations=False)
        for (pt, bev), o, full, one in ((got[0], of, s0, o0), (got[1], ob, s1, o1)):
            n = int(one["pcl_ta"]["pcl_is_valid"].shape[1])
            valid = one["pcl_ta"]["pcl_is_valid"][0]
            assert torch.equal(full["pcl_ta"]["pcl_is_valid"][b, :n], valid) and not bool(full["pcl_ta"]["pcl_is_valid"][b, n:].any())
            aee = _aee(pt[b, :n], o[-1]["pointwise_static_flow"][0], valid)
            assert aee <= 0.01, (b, aee)
            assert torch.equal(bev[b] != 0, o[-1]["static_flow"][0] != 0), b


@pytest.mark.parametrize("workload,seeds,graph", [("T", [3, 4], True), ("T", [8], False), ("N", [5], True)])
def test_train_mode_batchnorm_export_within_1cm(cuda, workload, seeds, graph):
    """Q4: the reference's flow export never calls model.eval() (experiment.py:164-198,225-361), so the pillar encoder's
    BatchNorm1d normalises with BATCH statistics (padded rows included, all samples of the batch pooled) and updates its
    running statistics between the two frames.  Same forward here with model.train() under no_grad -- also through the
    CUDA graph, whose part of the network does not depend on the mode -- against the oracle's bn_training path."""
    cfg = make_cfg(workload)
    model, sd = _model(cfg, cuda)
    model.train()
    model.raft_network.use_cuda_graph = graph
    s0, s1 = make_sample_dicts(WORKLOADS[workload], seeds)
    with torch.no_grad():
        pf, pb = model(s0, s1, None)
        torch.cuda.synchronize()
        of, ob, aux = SF.slim_forward(sd, cfg, s0, s1, bn_training=True, decode_all_iterations=False)
    assert (getattr(model.raft_network, "n_graph_captures", 0) > 0) == graph
    for p, o, s in ((pf, of, s0), (pb, ob, s1)):
        valid = s["pcl_ta"]["pcl_is_valid"]
        aee = _aee(p[-1].static_flow.cpu(), o[-1]["pointwise_static_flow"], valid)
        assert aee <= 0.01, aee
        assert torch.equal(p[-1].modified_network_output.static_flow.cpu() != 0, o[-1]["static_flow"] != 0)
    bn = model.raft_network.pp_layer.pts_voxel_encoder.pfn_layers[0].norm
    # two frames = two running-stat updates (momentum 0.01, unbiased variance), like the oracle's
    np.testing.assert_allclose(bn.running_mean.cpu().numpy(), aux["enc"][1]["running_mean"].numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(bn.running_var.cpu().numpy(), aux["enc"][1]["running_var"].numpy(), rtol=1e-4, atol=1e-6)
    # and the eval-mode answer differs (the test would not notice a silently ignored mode otherwise)
    ev, _ = _model(cfg, cuda)
    with torch.no_grad():
        ef, _ = ev(s0, s1, None)
    assert float((ef[-1].static_flow - pf[-1].static_flow).abs().max()) > 1e-4


def test_forward_results_survive_the_next_forward(cuda):
    """Graph mode: by default SLIM.forward hands out copies (the reference's export keeps the t0->t1 predictions across two
    more model() calls, experiment.py:386-456); with outputs_alias_static_buffers they are views that the next forward
    overwrites."""
    cfg = make_cfg("T")
    model, _ = _model(cfg, cuda)
    a0, a1 = make_sample_dicts(WORKLOADS["T"], [31])
    b0, b1 = make_sample_dicts(WORKLOADS["T"], [32])
    with torch.no_grad():
        pf, _ = model(a0, a1, None)
        keep_flow = pf[-1].modified_network_output.static_flow.clone()
        keep_pts = pf[-1].static_flow.clone()
        model(b0, b1, None)
        assert getattr(model.raft_network, "n_graph_captures", 0) > 0
        assert torch.equal(pf[-1].modified_network_output.static_flow, keep_flow) and torch.equal(pf[-1].static_flow, keep_pts)
        assert torch.equal(pf[0].modified_network_output.dynamicness, pf[0].modified_network_output.class_probs[..., 1])  # views of ONE copy
        model.outputs_alias_static_buffers = True
        qf, _ = model(a0, a1, None)
        assert torch.equal(qf[-1].modified_network_output.static_flow, keep_flow)
        model(b0, b1, None)
        assert not torch.equal(qf[-1].modified_network_output.static_flow, keep_flow)  # the documented hazard of the alias mode


def test_decode_last_equals_decode_all(cuda):
    """The export shortcut (decode only the last iteration, no Kabsch) returns the same exported tensors."""
    cfg = make_cfg("T")
    full, sd = _model(cfg, cuda)
    fast, _ = _model(cfg, cuda, decode_iterations="last", static_aggregation=False)
    s0, s1 = make_sample_dicts(WORKLOADS["T"], [9])
    with torch.no_grad():
        a_fw, a_bw = full(s0, s1, None)
        b_fw, b_bw = fast(s0, s1, None)
    assert len(a_fw) == 6 and len(b_fw) == 1
    for a, b in ((a_fw, b_fw), (a_bw, b_bw)):
        assert torch.equal(a[-1].modified_network_output.static_flow, b[-1].modified_network_output.static_flow)
        assert torch.equal(a[-1].modified_network_output.dynamicness, b[-1].modified_network_output.dynamicness)
        assert torch.equal(a[-1].static_flow, b[-1].static_flow)


def test_against_committed_golden(cuda):
    """Golden produced by the *unmodified reference* in the authoring container (oracle/gen_golden.py)."""
    path = os.path.join(GOLDEN, "slim_forward_tiny.npz")
    g = np.load(path)
    cfg = make_cfg("T")
    cfg.data.img_grid_size = tuple(int(v) for v in g["img_grid_size"])
    cfg.data.bev_range_m = tuple(float(v) for v in g["bev_range_m"])
    model, _ = _model(cfg, cuda, seed=int(g["weight_seed"]))

    def sample(t):
        pcl = torch.from_numpy(g["pcl_%s" % t])
        return {"pcl_full_no_ground_ta": [torch.from_numpy(g["full_%s" % t])],
                "pcl_ta": {"pcl": pcl[None], "pcl_is_valid": torch.ones(1, pcl.shape[0], dtype=torch.bool),
                           "pillar_coors": torch.from_numpy(g["coors_%s" % t])[None]},
                "gt": {"odom_ta_tb": torch.eye(4, dtype=torch.float64)[None]}}

    with torch.no_grad():
        pf, pb = model(sample("t0"), sample("t1"), None)
    for p, d in ((pf, "fw"), (pb, "bw")):
        ref_pt = torch.from_numpy(g["pt_static_flow_%s" % d])
        aee = float((p[-1].static_flow[0].cpu() - ref_pt).norm(dim=-1).mean())
        assert aee <= 0.01, aee
        ref_bev = torch.from_numpy(g["bev_static_flow_%s" % d])
        assert torch.equal(p[-1].modified_network_output.static_flow[0].cpu() != 0, ref_bev != 0)


def test_cuda_graphed_gru_loop_equals_eager_and_tracks_weight_updates(cuda):
    """the CUDA-graphed refinement loop replays the same kernels as the eager loop (bit-identical outputs),
    is re-captured when a weight changes in place, and different inputs flow through the same captured graph."""
    cfg = make_cfg("T")
    model, sd = _model(cfg, cuda)
    s0, s1 = make_sample_dicts(WORKLOADS["T"], [31])
    t0, t1 = make_sample_dicts(WORKLOADS["T"], [32])
    with torch.no_grad():
        model.raft_network.use_cuda_graph = False
        e_a = model(s0, s1, None)[0][-1].modified_network_output.static_flow.clone()
        e_b = model(t0, t1, None)[0][-1].modified_network_output.static_flow.clone()
        model.raft_network.use_cuda_graph = True
        g_a = model(s0, s1, None)[0][-1].modified_network_output.static_flow.clone()   # capture + replay
        g_b = model(t0, t1, None)[0][-1].modified_network_output.static_flow.clone()   # replay with new inputs
        assert torch.equal(e_a, g_a) and torch.equal(e_b, g_b) and not torch.equal(g_a, g_b)
        model.raft_network.update_block.static_flow_head.conv2.bias.add_(0.25)         # in-place weight update
        g_c = model(s0, s1, None)[0][-1].modified_network_output.static_flow.clone()
        model.raft_network.use_cuda_graph = False
        e_c = model(s0, s1, None)[0][-1].modified_network_output.static_flow.clone()
        assert torch.equal(g_c, e_c) and not torch.equal(g_c, g_a)


def test_graphed_decoder_static_inputs_follow_the_batch(cuda):
    """The output decoders run as side branches of the CUDA graph on static input buffers (points, pillar coordinates,
    validity) that are refreshed before every replay: pairs with fewer / more points than the captured capacity, and a
    switch of the decode mode, give exactly what the eager decoder gives."""
    cfg = make_cfg("T")
    model, _ = _model(cfg, cuda)
    model.POINT_CAPACITY_STEP = 256  # small steps so that the second pair outgrows the captured buffers
    pairs = [make_sample_dicts(WORKLOADS["T"], [51]), make_sample_dicts(WORKLOADS["T"], [52, 53])[0:2],
             make_sample_dicts(WORKLOADS["T"], [54])]

    def run(s0, s1):
        pf, pb = model(s0, s1, None)
        return [t.clone() for p in (pf[-1], pb[-1], pf[0])
                for t in (p.static_flow, p.static_aggr_flow, p.modified_network_output.dynamicness)], len(pf)

    with torch.no_grad():
        for s0, s1 in pairs:
            model.raft_network.use_cuda_graph = False
            eager, n_e = run(s0, s1)
            model.raft_network.use_cuda_graph = True
            graphed, n_g = run(s0, s1)
            assert n_e == n_g == 6
            for a, b in zip(eager, graphed):
                assert a.shape == b.shape and torch.equal(a, b)
        assert model.raft_network.n_graph_captures >= 2  # batch size 1 -> 2 -> 1 (and grown point buffers)
        model.decode_iterations = model.raft_network.output_iterations = "last"
        s0, s1 = pairs[0]
        pf, pb = model(s0, s1, None)
        assert len(pf) == len(pb) == 1
        assert torch.equal(pf[-1].static_flow, eager_last(model, s0, s1))


def eager_last(model, s0, s1):
    model.raft_network.use_cuda_graph = False
    try:
        return model(s0, s1, None)[0][-1].static_flow.clone()
    finally:
        model.raft_network.use_cuda_graph = True


def test_graph_replay_after_eager_calls_returns_the_graphs_own_results(cuda):
    """The sink (decoder) only runs while the graph is captured; eager calls made in between must not leak their
    results into later replays of the still valid graph."""
    cfg = make_cfg("T")
    model, _ = _model(cfg, cuda)
    s0, s1 = make_sample_dicts(WORKLOADS["T"], [61])
    t0, t1 = make_sample_dicts(WORKLOADS["T"], [62])
    n_s, n_t = s0["pcl_ta"]["pcl_is_valid"].shape[1], t0["pcl_ta"]["pcl_is_valid"].shape[1]
    with torch.no_grad():
        g_s = model(s0, s1, None)[0][-1].static_flow.clone()          # capture + replay
        model.raft_network.use_cuda_graph = False
        e_t = model(t0, t1, None)[0][-1].static_flow.clone()          # eager call on another pair
        model.raft_network.use_cuda_graph = True
        g_s2 = model(s0, s1, None)[0][-1].static_flow.clone()         # replay of the graph captured above
        g_t = model(t0, t1, None)[0][-1].static_flow.clone()
    assert g_s.shape[1] == n_s and g_t.shape[1] == n_t and e_t.shape[1] == n_t
    assert torch.equal(g_s2, g_s)
    assert torch.equal(g_t, e_t)


def test_triple_export_pass_equals_three_forwards(cuda, tmp_path):
    """SURVEY 8f.3 / experiment.py:386-456: SLIM.forward_triple (every frame encoded once, six directions in one graph)
    returns what three independent SLIM.forward calls return -- t0 -> t1, t0 -> t2, t1 -> t2 -- bit for bit in the
    exported maps; and run_flow_export writes them as the reference's 12-array file."""
    from liso_b200.slim import export
    from liso_b200.synth import SyntheticExportDataset

    cfg = make_cfg("T")
    model, _ = _model(cfg, cuda, decode_iterations="last", static_aggregation=False)
    # (encoders frame by frame: with the frames of a pass batched through the encoders a triple runs them on 3 B samples and a
    # pair on 2 B, and cuDNN's summation order depends on the batch size -- checked at round-off level below)
    model.raft_network.batched_encoders = False
    ds = SyntheticExportDataset(WORKLOADS["T"], 5, frames=3, pool=3)
    items = [ds[i] for i in range(2)]
    batch = export.collate_pairs([it[1:] for it in items])
    with torch.no_grad():
        tri = model.forward_triple(*batch)
        tri = {k: (v[-1].modified_network_output.static_flow.clone(), v[-1].modified_network_output.dynamicness.clone(),
                   v[-1].static_flow.clone()) for k, v in tri.items()}
        enc_before = _lib.load().slimb200_launch_count(_lib.K_PILLAR_NHWC)
        model.forward_triple(*batch)
        assert _lib.load().slimb200_launch_count(_lib.K_PILLAR_NHWC) - enc_before == 1  # three frames, ONE encoder pass (12 in the reference)
        for a, b in ((0, 1), (0, 2), (1, 2)):
            pf, pb = model(batch[a], batch[b], None)
            for key, p in (("t%d_t%d" % (a, b), pf), ("t%d_t%d" % (b, a), pb)):
                assert torch.equal(tri[key][0], p[-1].modified_network_output.static_flow), key
                assert torch.equal(tri[key][1], p[-1].modified_network_output.dynamicness), key
                assert torch.equal(tri[key][2], p[-1].static_flow), key
        # default setting (all frames of the pass through the encoders at once): the same maps to fp32 round-off
        model.raft_network.batched_encoders = True
        tri_b = model.forward_triple(*batch)
        for key, ref in tri.items():
            got = tri_b[key][-1].modified_network_output.static_flow
            assert torch.equal(got != 0, ref[0] != 0), key
            assert float((got - ref[0]).abs().max()) <= 1e-3 * max(1.0, float(ref[0].abs().max())), key
        model.raft_network.batched_encoders = False
    out = export.run_flow_export(model, ds, str(tmp_path), cfg.data.bev_range_m, batch_size=2, device=cuda, writer_workers=2)
    assert out["pairs"] == 5 and out["files"] == 5
    z = np.load(os.path.join(str(tmp_path), "000001.npz"))
    assert len(z.files) == 14
    assert np.array_equal(z["bev_raw_flow_t2_t0"], tri["t2_t0"][0][1].cpu().numpy())
    assert np.array_equal(z["bev_dynamicness_t1_t2"], tri["t1_t2"][1][1].cpu().numpy())


def test_gpu_compressed_export_writes_the_same_files(cuda, tmp_path):
    """run_flow_export(compress_on_gpu=True) -- maps deflated on the device, zip members framed on the host --
    writes files np.load reads exactly like the np.savez_compressed ones (pair and triple schema), several batches deep so
    that every download slot is reused."""
    from liso_b200.slim import export
    from liso_b200.synth import SyntheticExportDataset

    cfg = make_cfg("T")
    model, _ = _model(cfg, cuda, decode_iterations="last", static_aggregation=False)
    for frames, n_keys in ((2, 6), (3, 14)):
        ds = SyntheticExportDataset(WORKLOADS["T"], 9, frames=frames, pool=3)
        d_ref, d_gpu = os.path.join(str(tmp_path), "ref%d" % frames), os.path.join(str(tmp_path), "gpu%d" % frames)
        a = export.run_flow_export(model, ds, d_ref, cfg.data.bev_range_m, batch_size=2, device=cuda, writer_workers=2,
                                   compress_on_gpu=False)  # np.savez_compressed on host threads
        b = export.run_flow_export(model, ds, d_gpu, cfg.data.bev_range_m, batch_size=2, device=cuda, writer_workers=2,
                                   compress_on_gpu=True)
        assert a["files"] == b["files"] == 9
        again = export.run_flow_export(model, ds, d_gpu, cfg.data.bev_range_m, batch_size=2, device=cuda, writer_workers=2,
                                       compress_on_gpu=True, skip_existing=True)
        assert again["files"] == 0 and again["skipped"] == 9 and again["pairs"] == 0  # experiment.py:380-382: existing targets are skipped
        for i in range(9):
            zr, zg = np.load(os.path.join(d_ref, "%06d.npz" % i)), np.load(os.path.join(d_gpu, "%06d.npz" % i))
            assert sorted(zr.files) == sorted(zg.files) and len(zg.files) == n_keys
            for k in zr.files:
                assert zr[k].dtype == zg[k].dtype and zr[k].shape == zg[k].shape, k
                assert np.array_equal(zr[k].view(np.uint8) if zr[k].ndim else zr[k], zg[k].view(np.uint8) if zg[k].ndim else zg[k]), (i, k)


def test_gpu_export_from_raw_scans_with_loader_threads(cuda, tmp_path):
    """SURVEY 8f.3 + 8f.4: the export fed with RAW scans only (the device applies the ground rule, computes the pillar map and
    compacts the decoder inputs), dataset access on loader threads, maps deflated on the device: every file equals what
    SLIM.forward returns for the same scans prepared by preprocess_scans."""
    from liso_b200.datasets import preprocess_scans
    from liso_b200.slim import export
    from liso_b200.synth import SyntheticExportDataset

    cfg = make_cfg("T")
    model, _ = _model(cfg, cuda, decode_iterations="last", static_aggregation=False)
    ds = SyntheticExportDataset(WORKLOADS["T"], 7, frames=2, pool=2, raw=True)
    out = export.run_flow_export(model, ds, str(tmp_path), cfg.data.bev_range_m, batch_size=3, device=cuda, writer_workers=2,
                                 compress_on_gpu=True, loader_workers=2)
    assert out["pairs"] == 7 and out["files"] == 7 and out["d2h_bytes"] > 0
    assert out["d2h_bytes"] < 0.5 * 7 * 6 * WORKLOADS["T"]["img_grid_size"][0] * WORKLOADS["T"]["img_grid_size"][1] * 4
    for chunk in ((0, 1, 2), (3, 4, 5), (6, 6, 6)):  # the export's batches (the ragged last one padded with its last sample)
        items = [ds[i] for i in chunk]
        with torch.no_grad():
            pf, pb = model(preprocess_scans([it[1]["pcl_full_w_ground_ta"].to(cuda) for it in items], cfg),
                           preprocess_scans([it[2]["pcl_full_w_ground_ta"].to(cuda) for it in items], cfg), None)
        for b, i in enumerate(chunk):
            if b and i == chunk[b - 1]:
                continue  # padding copies: the file holds slot 0 (stock convolutions differ by ~1e-6 between batch positions)
            z = np.load(os.path.join(str(tmp_path), "%06d.npz" % i))
            assert np.array_equal(z["bev_raw_flow_t0_t1"], pf[-1].modified_network_output.static_flow[b].cpu().numpy()), i
            assert np.array_equal(z["bev_dynamicness_t1_t0"], pb[-1].modified_network_output.dynamicness[b].cpu().numpy()), i
            assert float(np.abs(z["bev_raw_flow_t0_t1"]).sum()) > 0
    assert sorted(os.listdir(str(tmp_path))) == ["%06d.npz" % i for i in range(7)]
