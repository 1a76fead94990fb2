"""CPU: host-side logic around the kernels -- module wiring, state-dict contract, frame sharding and
the world_size-2 gather (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.slim import export
from liso_b200.synth import make_sample_dicts, pillar_coors_f64_numpy
from liso_b200.weights import synth_weights_like
from oracle import slim_forward as SF
from oracle import slim_oracle as O

REFERENCE_KEYS_HEAD = [
    "moving_dynamicness_threshold.start_value",
    "moving_dynamicness_threshold.update_weight",
    "moving_dynamicness_threshold.bias_counter",
    "moving_dynamicness_threshold.moving_average_importance",
    "raft_network.pp_layer.pts_voxel_encoder.pfn_layers.0.norm.weight",
    "raft_network.pp_layer.pts_voxel_encoder.pfn_layers.0.norm.bias",
    "raft_network.pp_layer.pts_voxel_encoder.pfn_layers.0.norm.running_mean",
    "raft_network.pp_layer.pts_voxel_encoder.pfn_layers.0.norm.running_var",
    "raft_network.pp_layer.pts_voxel_encoder.pfn_layers.0.norm.num_batches_tracked",
    "raft_network.pp_layer.pts_voxel_encoder.pfn_layers.0.linear.weight",
    "raft_network.fnet.norm1.weight",
]


def test_state_dict_contract():
    """Keys the reference SLIM checkpoint has (verified against the live reference
    in test_oracle_vs_reference.py): 150 entries, 2,423,606 parameters, quirky norm3/downsample.1 aliases."""
    from liso_b200.slim.slim import SLIM

    m = SLIM(make_cfg("K"))
    keys = list(m.state_dict().keys())
    assert keys[: len(REFERENCE_KEYS_HEAD)] == REFERENCE_KEYS_HEAD
    assert len(keys) == 150
    assert sum(p.numel() for p in m.parameters()) == 2423606
    assert "raft_network.fnet.layer2.1.downsample.1.weight" in keys and "raft_network.fnet.layer2.1.norm3.weight" in keys
    assert "raft_network.cnet.layer2.1.downsample.0.weight" in keys and "raft_network.cnet.layer1.1.downsample.0.weight" not in keys
    assert tuple(m.state_dict()["raft_network.update_block.gru.convz.weight"].shape) == (96, 304, 3, 3)


def test_slim_wiring_with_oracle_stages(monkeypatch):
    """Everything around the two CUDA stages (stock encoders, GRU loop, upsampling, decoder) equals the
    oracle port when the two stages are replaced by their oracle counterparts."""
    import liso_b200.slim.raft as raft_mod
    from liso_b200.slim.slim import SLIM

    cfg = make_cfg("T")
    m = SLIM(cfg).eval()
    sd = synth_weights_like(m.state_dict(), 0)
    m.load_state_dict(sd, strict=True)

    class OracleCorr:
        def __init__(self, f1, f2, num_levels=4, radius=4):
            self.pyr, self.r = O.corr_pyramid(f1, f2, num_levels), radius

        def __call__(self, coords):
            return O.corr_lookup(self.pyr, coords, self.r)

    monkeypatch.setattr(raft_mod, "CorrBlock", OracleCorr)
    pp = m.raft_network.pp_layer
    pfn = pp.pts_voxel_encoder.pfn_layers[0]

    def pp_forward(pts, img=None):
        params = dict(linear_weight=pfn.linear.weight, bn_weight=pfn.norm.weight, bn_bias=pfn.norm.bias,
                      running_mean=pfn.norm.running_mean, running_var=pfn.norm.running_var)
        e = O.pillar_encoder_forward([p.numpy() for p in pts], params, cfg.data.bev_range_m, cfg.data.img_grid_size, 10.0, False)
        return e["canvas"], e["occupancy"]

    monkeypatch.setattr(pp, "forward", pp_forward)
    # the output decoder is CUDA-only in the product too: the stock-PyTorch restatement of the tests stands in
    from torch_decoder import forward_torch

    for dec in (m.head_decoder_fw, m.head_decoder_bw):
        monkeypatch.setattr(dec, "forward", lambda o, thr, _d=dec, **kw: forward_torch(
            _d, o, thr, pc=kw["pc"], pointwise_voxel_coordinates=kw["pointwise_voxel_coordinates"],
            pointwise_valid_mask=kw["pointwise_valid_mask"], filled_pillar_mask=kw["filled_pillar_mask"],
            static_aggregation=kw.get("static_aggregation", True)))
    s0, s1 = make_sample_dicts(WORKLOADS["T"], [3])
    with torch.no_grad():
        pf, pb = m(s0, s1, None)
        of, ob, _ = SF.slim_forward(sd, cfg, s0, s1)
    assert len(pf) == len(pb) == 6
    for it in (0, 5):
        for p, o in ((pf, of), (pb, ob)):
            assert torch.equal(p[it].modified_network_output.static_flow, o[it]["static_flow"])
            assert torch.equal(p[it].modified_network_output.dynamicness, o[it]["dynamicness"])
            assert torch.equal(p[it].static_flow, o[it]["pointwise_static_flow"])
            assert torch.allclose(p[it].static_aggr_trafo, o[it]["static_aggr_trafo"], atol=1e-9)


def test_synth_pillar_coors_match_oracle_a12():
    W = WORKLOADS["A"]
    rng = np.random.default_rng(0)
    pts = rng.uniform(-61, 61, size=(50000, 4)).astype(np.float32)
    pts[:, 2] = rng.uniform(-2.5, 1.5, size=pts.shape[0])
    c0, ok0 = pillar_coors_f64_numpy(pts, W["bev_range_m"], W["img_grid_size"])
    c1, ok1 = O.pillar_coors_f64(pts, W["bev_range_m"], W["img_grid_size"])
    assert np.array_equal(ok0, ok1) and np.array_equal(c0, c1)


def test_shard_indices_cover_exactly_once():
    """experiment.py:330-332: sample_idx % world_size == worker_id."""
    for n in (0, 1, 7, 10000):
        for ws in (1, 2, 4, 8):
            seen = np.zeros(n, dtype=np.int32)
            for r in range(ws):
                idx = export.shard_indices(n, ws, r)
                assert all(i % ws == r for i in idx)
                seen[idx] += 1
            assert (seen == 1).all()
    with pytest.raises(ValueError):
        export.shard_indices(10, 2, 2)
    assert list(export.iterate_batches(range(5), 2)) == [[0, 1], [2, 3], [4]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = export.shard_indices(11, world, rank)
    local = {"pairs": float(len(idx)), "index_sum": float(sum(idx)), "elapsed_s_max": 1.0 + rank}
    tot = export.reduce_counters(local)
    torch.save(tot, os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_world_size_2_gather_gloo(tmp_path):
    """The N>1 path: shard by the modulo rule, one collective at the end; totals equal the single-process run."""
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    single = export.reduce_counters({"pairs": 11.0, "index_sum": float(sum(range(11))), "elapsed_s_max": 2.0})
    for r in range(2):
        tot = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        assert tot == single


def test_async_npz_writer_schema_matches_reference_consumer(tmp_path):
    """files written off the critical path have the reference's schema (experiment.py:391-402,459-471) and
    can be consumed like torch_dataset_commons.py:619-675 does (np.load, bev_range_m, bev_raw_flow_* [H,W,2])."""
    from liso_b200.slim.export import AsyncNpzWriter

    rng = np.random.default_rng(0)
    H, W, B = 32, 24, 3
    host = [torch.from_numpy(rng.normal(size=(B, H, W, 2)).astype(np.float32)),
            torch.from_numpy(rng.normal(size=(B, H, W, 2)).astype(np.float32)),
            torch.from_numpy(rng.uniform(size=(B, H, W)).astype(np.float32)),
            torch.from_numpy(rng.uniform(size=(B, H, W)).astype(np.float32))]
    keep = [t.clone() for t in host]
    w = AsyncNpzWriter(str(tmp_path), bev_range_m=(70.0, 70.0), workers=2, max_pending=2)
    ids = ["seq_a/000001", "seq_a/000002", "000003"]
    assert w.submit_batch(ids, host, torch.tensor(0.5)) == 3
    for t in host:  # the caller reuses its (pinned) buffers right away
        t.zero_()
    assert w.close() == 3
    for b, sid in enumerate(ids):
        f = np.load(tmp_path / (sid + ".npz"), allow_pickle=True)
        assert set(f.files) == {"static_threshold", "bev_raw_flow_t0_t1", "bev_raw_flow_t1_t0", "bev_dynamicness_t0_t1",
                                "bev_dynamicness_t1_t0", "bev_range_m"}
        assert f["bev_raw_flow_t0_t1"].shape == (H, W, 2) and f["bev_raw_flow_t0_t1"].dtype == np.float32
        assert f["bev_dynamicness_t1_t0"].shape == (H, W)
        assert np.array_equal(f["bev_raw_flow_t0_t1"], keep[0][b].numpy()) and np.array_equal(f["bev_dynamicness_t1_t0"], keep[3][b].numpy())
        assert float(f["static_threshold"]) == 0.5 and f["static_threshold"].shape == ()
        grid_size = np.append(f["bev_raw_flow_t0_t1"].shape[:2], np.array(1))  # what the consumer derives
        assert tuple(grid_size) == (H, W, 1) and np.allclose(f["bev_range_m"], (70.0, 70.0))
    w2 = AsyncNpzWriter(str(tmp_path), bev_range_m=(70.0, 70.0), workers=1, skip_existing=True)
    assert w2.submit_batch(ids, keep, 0.5) == 0 and w2.close() == 0


def test_stacked_parallel_convolutions_equal_the_separate_ones():
    """`raft._stacked_params`: one block-diagonal (or output-stacked) convolution == the two parallel convolutions it
    replaces (update.py:49-66: flow / logits branches; update.py:6-20: the two heads), padding channels ignored."""
    import torch
    from torch import nn

    from liso_b200.slim import raft as R

    torch.manual_seed(0)
    a, b = nn.Conv2d(2, 64, 7, padding=3), nn.Conv2d(4, 64, 7, padding=3)
    w, bias = R._stacked_params(a, b, shared_input=False, pad_in_to=8)
    assert tuple(w.shape) == (128, 8, 7, 7)
    x = torch.randn(2, 8, 9, 11)
    ref = torch.cat([a(x[:, :2]), b(x[:, 2:6])], dim=1)
    got = torch.nn.functional.conv2d(x, w, bias, padding=3)  # channels 6, 7 carry garbage: zero weights
    assert float((ref - got).abs().max()) < 1e-5
    c, d = nn.Conv2d(96, 128, 3, padding=1), nn.Conv2d(96, 128, 3, padding=1)
    w2, b2 = R._stacked_params(c, d, shared_input=True)
    h = torch.randn(1, 96, 6, 5)
    assert float((torch.cat([c(h), d(h)], 1) - torch.nn.functional.conv2d(h, w2, b2, padding=1)).abs().max()) < 1e-5
    assert R._same_geometry(a, b) and not R._same_geometry(a, c)
    # the heads' 3x3 output convolution as 1x1 taps + window sum (what slimb200_iter_update_taps evaluates)
    e = nn.Conv2d(16, 6, 3, padding=1, bias=False)
    w_taps = e.weight.detach().permute(2, 3, 0, 1).reshape(54, 16, 1, 1)
    y = torch.randn(1, 16, 7, 8)
    taps = torch.nn.functional.conv2d(y, w_taps).reshape(1, 3, 3, 6, 7, 8)
    out = torch.zeros(1, 6, 7, 8)
    padded = torch.nn.functional.pad(taps, (1, 1, 1, 1))
    for ky in range(3):
        for kx in range(3):
            out += padded[:, ky, kx, :, ky:ky + 7, kx:kx + 8]
    assert float((out - e(y)).abs().max()) < 1e-5


def test_pyramid_layout_formula_matches_header():
    """The 4-pixel x 8-column interleaved half-tile layout of include/slimb200.h: pack -> address formula -> unpack."""
    import random

    import torch

    from liso_b200.slim import corr as Cc

    random.seed(0)
    for B, h, w in ((2, 12, 20), (1, 23, 29)):
        L = Cc.make_layout(B, 128, h, w, 4)
        nf = h * w
        levels = [torch.randn(B * nf, 1, L.level_h[l], L.level_w[l]) for l in range(4)]
        packed = Cc.pack_pyramid_f32(levels, L)
        assert packed.numel() == B * L.n_panels * L.rows_padded * 128 and L.rows_padded % 128 == 0 and L.rows_padded >= nf
        flat = packed.view(-1)
        for l in range(4):
            assert torch.equal(Cc.unpack_level(packed, L, l), levels[l])
        for _ in range(500):
            b, i, l = random.randrange(B), random.randrange(nf), random.randrange(4)
            y, x = random.randrange(L.level_h[l]), random.randrange(L.level_w[l])
            j = L.level_offset[l] + y * L.level_w[l] + x
            c = j % 128
            t = ((b * L.n_panels + j // 128) * (L.rows_padded // 128) + i // 128) * 2 + c // 64
            idx = t * 8192 + ((i % 128) // 4) * 256 + ((c % 64) // 8) * 32 + (i % 4) * 8 + c % 8
            assert flat[idx] == levels[l][b * nf + i, 0, y, x]


def test_decoded_outputs_of_a_replayed_graph_are_the_captured_ones():
    """`SLIM._sink_results` bookkeeping (no GPU needed): eager calls between two replays of a graph overwrite the sink's
    scratch dict, the replay must still hand out the outputs recorded when the graph was captured."""
    import types

    import pytest

    from liso_b200.config import make_cfg
    from liso_b200.slim.slim import SLIM

    m = SLIM(make_cfg("T"))
    net = types.SimpleNamespace(_graphs={}, n_graph_captures=0)
    m._sink_preds = {(0, 5): "captured"}
    net._graphs["net"], net.n_graph_captures = {"key": "K1"}, 1
    assert m._sink_results(net, True, 0) == {(0, 5): "captured"}
    m._sink_preds = {(0, 5): "eager"}
    assert m._sink_results(net, False, 1) == {(0, 5): "eager"}
    assert m._sink_results(net, True, 1) == {(0, 5): "captured"}       # replay, no new capture
    m._sink_preds = {(0, 5): "captured again"}
    net._graphs["net"], net.n_graph_captures = {"key": "K2"}, 2
    assert m._sink_results(net, True, 1) == {(0, 5): "captured again"}
    net._graphs["net"] = {"key": "K3"}                                   # a graph SLIM has no record of
    with pytest.raises(RuntimeError):
        m._sink_results(net, True, 2)


def test_run_flow_export_shards_batches_and_writes_reference_schema(tmp_path):
    """`run_flow_export` (experiment.py:225-361): modulo sharding, collate padding, one .npz per pair in the reference's
    schema, skip_existing -- with a stand-in pipeline that produces the exported tensors on the CPU."""
    import numpy as np
    import torch

    from liso_b200.slim import export

    H = W = 4

    class FakeThreshold:
        def value(self):
            return torch.tensor(0.25)

    class FakeModel:
        moving_dynamicness_threshold = FakeThreshold()

    class FakePipeline:  # same contract as ExportPipeline.run: consume(index, [flow_fw, flow_bw, dyn_fw, dyn_bw])
        def __init__(self, model, device):
            self.seen = []

        def run(self, batches, consume):
            n = 0
            for j, (d0, d1) in enumerate(batches):
                n += 1
                B = len(d0["pcl_full_no_ground_ta"])
                assert d0["pcl_ta"]["pcl"].shape[0] == B and d0["pcl_ta"]["pcl_is_valid"].dtype == torch.bool
                # padding rows: NaN points, -1 coordinates, invalid
                inval = ~d0["pcl_ta"]["pcl_is_valid"]
                assert bool(torch.isnan(d0["pcl_ta"]["pcl"][inval]).all()) and bool((d0["pcl_ta"]["pillar_coors"][inval] == -1).all())
                n0 = torch.tensor([float(t.shape[0]) for t in d0["pcl_full_no_ground_ta"]])
                flow = n0[:, None, None, None].expand(B, H, W, 2).contiguous()
                consume(j, [flow, -flow, flow[..., 0].contiguous(), flow[..., 1].contiguous()])
            return n

    def sample(n):
        return {"pcl_full_no_ground_ta": torch.zeros(n + 3, 4),
                "pcl_ta": {"pcl": torch.ones(n, 4), "pillar_coors": torch.zeros(n, 2, dtype=torch.int32)}}

    dataset = [("seq/%03d" % i, sample(5 + i), sample(7 + i)) for i in range(7)]
    out = export.run_flow_export(FakeModel(), dataset, str(tmp_path), (70.0, 70.0), world_size=2, worker_id=1, batch_size=2,
                                 pipeline_factory=FakePipeline, writer_workers=2)
    assert out["pairs"] == 3 and out["files"] == 3 and out["skipped"] == 0  # indices 1, 3, 5
    files = sorted(p.name for p in (tmp_path / "seq").iterdir())
    assert files == ["001.npz", "003.npz", "005.npz"]
    z = np.load(tmp_path / "seq" / "003.npz")
    assert set(z.files) == {"static_threshold", "bev_raw_flow_t0_t1", "bev_raw_flow_t1_t0", "bev_dynamicness_t0_t1",
                            "bev_dynamicness_t1_t0", "bev_range_m"}
    assert z["bev_raw_flow_t0_t1"].shape == (H, W, 2) and float(z["bev_raw_flow_t0_t1"][0, 0, 0]) == 5 + 3 + 3
    assert float(z["static_threshold"]) == 0.25 and tuple(z["bev_range_m"]) == (70.0, 70.0)
    again = export.run_flow_export(FakeModel(), dataset, str(tmp_path), (70.0, 70.0), world_size=2, worker_id=1, batch_size=2,
                                   pipeline_factory=FakePipeline, skip_existing=True)
    assert again["pairs"] == 0 and again["skipped"] == 3


class _CpuThreshold:
    def value(self):
        return torch.tensor(0.5)


class _CpuModel:
    moving_dynamicness_threshold = _CpuThreshold()


class _CpuPipeline:
    """Stand-in for ExportPipeline (same `run(batches, consume)` contract) that fabricates the exported tensors."""

    def __init__(self, model, device):
        pass

    def run(self, batches, consume):
        n = 0
        for j, (d0, _d1) in enumerate(batches):
            B = len(d0["pcl_full_no_ground_ta"])
            flow = torch.zeros(B, 2, 2, 2)
            consume(j, [flow, flow, flow[..., 0].contiguous(), flow[..., 0].contiguous()])
            n += 1
        return n


def _export_dataset(n):
    def sample(k):
        return {"pcl_full_no_ground_ta": torch.zeros(k + 2, 4),
                "pcl_ta": {"pcl": torch.ones(k, 4), "pillar_coors": torch.zeros(k, 2, dtype=torch.int32)}}

    return [("p%02d" % i, sample(3 + i), sample(4 + i)) for i in range(n)]


def _gloo_export_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tot = export.run_flow_export(_CpuModel(), _export_dataset(9), os.path.join(out_dir, "npz"), (70.0, 70.0), world_size=world,
                                 worker_id=rank, batch_size=2, pipeline_factory=_CpuPipeline, writer_workers=1)
    torch.save(tot, os.path.join(out_dir, "t%d.pt" % rank))
    dist.destroy_process_group()


def test_world_size_2_flow_export_gloo(tmp_path):
    """Two workers export disjoint shares of the pairs (modulo rule) into one directory; the reduced counters equal the
    whole job on every rank and every pair has exactly one file."""
    port = _free_port()
    mp.spawn(_gloo_export_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        tot = torch.load(os.path.join(str(tmp_path), "t%d.pt" % r))
        assert tot["pairs"] == 9.0 and tot["files"] == 9.0 and tot["skipped"] == 0.0 and tot["elapsed_s_max"] > 0.0
    assert sorted(os.listdir(os.path.join(str(tmp_path), "npz"))) == ["p%02d.npz" % i for i in range(9)]


def test_derived_parameter_cache_is_per_module_and_identity_checked():
    """Derived weights (stacked update|reset gates, block-diagonal pairs, tap weights) are cached on the module that owns
    the sources and checked by tensor IDENTITY + version: a new model whose parameters land on the addresses of a freed
    one must never be served the old model's concatenations."""
    import gc

    from liso_b200.slim import raft as R

    stale = 0
    for seed in range(6):
        torch.manual_seed(seed)
        a, b = torch.nn.Conv2d(8, 4, 3, padding=1), torch.nn.Conv2d(8, 4, 3, padding=1)
        owner = torch.nn.Module()
        owner.a, owner.b = a, b
        got = R._cat_params(owner, "w", (a.weight, b.weight))
        assert torch.equal(got, torch.cat([a.weight, b.weight]).detach())
        assert R._cat_params(owner, "w", (a.weight, b.weight)) is got                      # hit
        w, bias = R._stacked_params(a, b, shared_input=False, pad_in_to=24)
        assert w.shape == (8, 24, 3, 3) and torch.equal(w[:4, :8], a.weight.detach()) and torch.equal(w[4:, 8:16], b.weight.detach())
        assert float(w[:4, 8:].abs().max()) == 0.0 and torch.equal(bias, torch.cat([a.bias, b.bias]).detach())
        with torch.no_grad():
            a.weight.add_(1.0)                                                              # in-place update: new version
        again = R._cat_params(owner, "w", (a.weight, b.weight))
        assert again is not got and torch.equal(again[:4], a.weight.detach())
        stale += int(not torch.equal(R._stacked_params(a, b, shared_input=False, pad_in_to=24)[0][:4, :8], a.weight.detach()))
        del a, b, owner, got, again, w, bias
        gc.collect()
    assert stale == 0
    assert not hasattr(R, "_PARAM_CAST_CACHE")  # no process-global cache keyed by raw addresses


def _rn32(fr):
    """Fraction -> nearest float32 (ties to even), as an exactly representable python float."""
    import math
    from fractions import Fraction

    if fr == 0:
        return 0.0
    sgn = -1 if fr < 0 else 1
    a = abs(fr)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    while Fraction(2) ** e > a:
        e -= 1
    while Fraction(2) ** (e + 1) <= a:
        e += 1
    e = max(e, -126)
    scaled = a / (Fraction(2) ** (e - 23))
    m = scaled.numerator // scaled.denominator
    rem = scaled - m
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (m & 1)):
        m += 1
    return sgn * float(Fraction(m) * Fraction(2) ** (e - 23))


def test_lookup_division_by_precomputed_reciprocal_is_correctly_rounded():
    """csrc/lookup_core.cuh::div_by_const -- q = a * rinv; twice q += fma(-b, q, a) * rinv with rinv = RN(1 / b) -- equals
    the IEEE quotient RN(a / b) that bilinear_sampler's `2 * x / (W - 1)` (raft_code/utils.py:19-20) asks for.  Checked in
    exact rational arithmetic for every divisor a level size up to 130 can produce and typical / adversarial dividends."""
    import random
    from fractions import Fraction

    def fma32(x, y, z):
        return _rn32(Fraction(x) * Fraction(y) + Fraction(z))

    random.seed(1)
    n = 0
    for b in list(range(1, 130)) + [159, 199, 255, 1023, 4095]:
        bf = float(b)
        y = _rn32(Fraction(1) / Fraction(bf))
        for _ in range(12):
            kind = random.random()
            if kind < 0.5:
                a = np.float32(2.0 * random.uniform(-10, b + 10))
            elif kind < 0.8:
                a = np.float32(2.0 * (random.randint(-8, b + 8) + random.choice([0, 1e-6, -1e-6, 0.5])))
            else:
                a = np.float32(random.uniform(-2e7, 2e7) * random.choice([1, 1e-3, 1e-30, 1e-41]))
            a = float(a)
            q = _rn32(Fraction(a) * Fraction(y))
            q = fma32(fma32(-bf, q, a), y, q)
            q = fma32(fma32(-bf, q, a), y, q)
            assert q == _rn32(Fraction(a) / Fraction(bf)), (a, b)
            n += 1
    assert n > 1500


# (pinned against the reference source by tests/test_oracle_vs_reference.py::test_export_schema_keys_match_reference_source)
from tests_keys import REFERENCE_TRIPLE_KEYS  # noqa: E402


def test_triple_export_writes_the_twelve_maps_of_the_reference(tmp_path):
    """KITTI / nuScenes export: (t0, t1, t2) samples -> one file with t0<->t1, t0<->t2, t1<->t2 (experiment.py:404-456),
    consumable like torch_dataset_commons.py:619-675 (np.load -> bev_raw_flow_* (H, W, 2), bev_dynamicness_* (H, W))."""
    H, W = 6, 5
    seen = []

    class FakePipeline:
        def __init__(self, model, device):
            pass

        def run(self, batches, consume):
            n = 0
            for j, batch in enumerate(batches):
                assert len(batch) == 3  # collated t0, t1, t2
                B = len(batch[0]["pcl_full_no_ground_ta"])
                seen.append(B)
                n_pts = [torch.tensor([float(t.shape[0]) for t in s["pcl_full_no_ground_ta"]]) for s in batch]
                flows, dyns = [], []
                for k, d in enumerate(export.DIRECTIONS_TRIPLE):  # value = 100 * direction + points of the source frame
                    src = n_pts[int(d[1])]
                    f = (100.0 * k + src)[:, None, None, None].expand(B, H, W, 2).contiguous()
                    flows.append(f)
                    dyns.append(f[..., 0].contiguous() + 0.5)
                consume(j, flows + dyns)
                n += 1
            return n

    def sample(n):
        return {"pcl_full_no_ground_ta": torch.zeros(n, 4),
                "pcl_ta": {"pcl": torch.ones(n - 1, 4), "pillar_coors": torch.zeros(n - 1, 2, dtype=torch.int32)}}

    dataset = [("%03d" % i, sample(10 + i), sample(20 + i), sample(30 + i)) for i in range(5)]
    out = export.run_flow_export(_CpuModel(), dataset, str(tmp_path), (70.0, 70.0), batch_size=2, pipeline_factory=FakePipeline,
                                 writer_workers=2)
    assert out["pairs"] == 5 and out["files"] == 5 and seen == [2, 2, 2]  # the ragged last batch is padded, its copy not written
    assert sorted(os.listdir(tmp_path)) == ["%03d.npz" % i for i in range(5)]
    seen.clear()
    out = export.run_flow_export(_CpuModel(), dataset, str(tmp_path / "ragged"), (70.0, 70.0), batch_size=2, pipeline_factory=FakePipeline,
                                 writer_workers=2, pad_last_batch=False, loader_workers=2)
    assert out["files"] == 5 and seen == [2, 2, 1]
    z = np.load(tmp_path / "003.npz")
    assert set(z.files) == REFERENCE_TRIPLE_KEYS
    for k, d in enumerate(export.DIRECTIONS_TRIPLE):
        src = {"0": 13, "1": 23, "2": 33}[d[1]]
        assert z["bev_raw_flow_" + d].shape == (H, W, 2) and z["bev_raw_flow_" + d].dtype == np.float32
        assert float(z["bev_raw_flow_" + d][0, 0, 0]) == 100.0 * k + src
        assert z["bev_dynamicness_" + d].shape == (H, W) and float(z["bev_dynamicness_" + d][1, 1]) == 100.0 * k + src + 0.5
    assert export.export_keys(2) == ["bev_raw_flow_t0_t1", "bev_raw_flow_t1_t0", "bev_dynamicness_t0_t1", "bev_dynamicness_t1_t0"]


def test_synthetic_export_dataset_is_deterministic_and_distinct():
    from liso_b200.synth import SyntheticExportDataset

    ds = SyntheticExportDataset(WORKLOADS["T"], 9, frames=3, pool=2)
    a, b, c = ds[1], ds[3], SyntheticExportDataset(WORKLOADS["T"], 9, frames=3, pool=2)[3]
    assert len(a) == 4 and a[0] == "000001" and b[0] == "000003"
    assert torch.equal(b[1]["pcl_full_no_ground_ta"], c[1]["pcl_full_no_ground_ta"])                      # deterministic in i
    assert b[1]["pcl_full_no_ground_ta"].shape == a[1]["pcl_full_no_ground_ta"].shape                    # same cast scene ...
    assert not torch.equal(b[1]["pcl_full_no_ground_ta"], a[1]["pcl_full_no_ground_ta"])                # ... rotated + shifted
    assert not torch.equal(b[1]["pcl_ta"]["pillar_coors"][:50], a[1]["pcl_ta"]["pillar_coors"][:50])
    d0, d1, d2 = export.collate_pairs([ds[0][1:], ds[1][1:]])
    assert d2["pcl_ta"]["pcl"].shape[0] == 2 and d2["pcl_ta"]["pcl_is_valid"].dtype == torch.bool


def test_lazy_sample_sources_pack_like_materialised_ones():
    """Loader protocol: a sample source with shape / dtype / write_into produces its data straight into the pinned upload
    buffer (ShiftedCloud: shift + copy in one pass); the packed batch equals the one packed from materialised tensors."""
    from liso_b200.synth import SyntheticExportDataset

    lazy = SyntheticExportDataset(WORKLOADS["T"], 12, frames=3, pool=2, raw=True, motion="shift", lazy=True)
    eager = SyntheticExportDataset(WORKLOADS["T"], 12, frames=3, pool=2, raw=True, motion="shift", lazy=False)
    fl = export.collate_pairs([lazy[i][1:] for i in (3, 5)])
    fe = export.collate_pairs([eager[i][1:] for i in (3, 5)])
    arena = export.PinnedArena(2)
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(2) as ex:
        packed = arena.pack(fl, ex)
    views = packed.views(packed.slot["buf"].clone())
    assert len(views) == 3 and views[0]["raw_scan"] is True
    for t in range(3):
        for a, b in zip(views[t]["pcl_full_w_ground_ta"], fe[t]["pcl_full_w_ground_ta"]):
            assert torch.equal(a, b)
    pinned = export._pin(fl[1])  # the path without an arena materialises
    assert torch.equal(pinned["pcl_full_w_ground_ta"][0], fe[1]["pcl_full_w_ground_ta"][0])
    arena.give_back(packed.slot)
    assert arena.free.qsize() == 2
