"""SURVEY 8(f).1 parity on the B200: fused output decoder (slimb200_head_decode) vs the CPU oracle restatement of
HeadDecoder.forward + static aggregation + weighted Kabsch (oracle/slim_forward.py::head_decoder, pinned against the
executed reference by tests/golden/slim_forward_tiny.npz).

Stated tolerances: logits / flows are copies (exact); softmax probabilities <= 2e-6 abs (fp32 expf); class decisions
exact except where a probability sits within 2e-6 of its decision boundary; gathered point values identical to the BEV
cell they come from; Kabsch transform <= 2e-5 abs per entry (fp64 moments here, fp32 sums in the reference);
static-aggregated flow <= 2e-4 m.
"""
import numpy as np
import pytest
import torch

from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.slim.slim import HeadDecoder
from torch_decoder import forward_torch
from liso_b200.synth import make_sample_dicts
from oracle import slim_forward as SF

pytestmark = pytest.mark.gpu


def _inputs(workload, seeds, gen_seed, flow_scale=0.3):
    cfg = make_cfg(workload)
    s0, _ = make_sample_dicts(WORKLOADS[workload], seeds)
    H, W = cfg.data.img_grid_size
    B = len(seeds)
    g = torch.Generator().manual_seed(gen_seed)
    net_out = torch.randn(B, H, W, 8, generator=g)
    net_out[..., 4:8] *= flow_scale
    filled = torch.rand(B, H, W, generator=g) < 0.3
    pt = s0["pcl_ta"]
    return cfg, net_out, filled, pt["pcl"], pt["pillar_coors"], pt["pcl_is_valid"]


def _decoder(cfg):
    half = 0.5 * np.array(cfg.data.bev_range_m)
    return HeadDecoder(cfg.SLIM, name="t", bev_extent=np.concatenate([-half, half], axis=0))


@pytest.mark.parametrize("workload,seeds", [("T", [1, 2]), ("N", [3]), ("K", [4, 5])])
@pytest.mark.parametrize("aggregation", [True, False])
def test_fused_decoder_matches_oracle(cuda, workload, seeds, aggregation):
    cfg, net_out, filled, pc, coors, valid = _inputs(workload, seeds, 7)
    thr = 0.37
    dec = _decoder(cfg)
    ref = SF.head_decoder(net_out, thr, pc, coors, valid, filled, dec.bev_extent, 1, with_static_aggregation=aggregation)
    with torch.no_grad():
        got = dec(net_out.to(cuda), torch.tensor(thr, device=cuda), pc=pc.to(cuda), pointwise_voxel_coordinates=coors.to(cuda),
                  pointwise_valid_mask=valid.to(cuda), filled_pillar_mask=filled.to(cuda), static_aggregation=aggregation)
    torch.cuda.synchronize()
    md = got.modified_network_output
    assert torch.equal(md.static_flow.cpu(), ref["static_flow"]) and torch.equal(md.dynamic_flow.cpu(), ref["dynamic_flow"])
    assert torch.equal(md.class_logits.cpu(), ref["class_logits"])
    assert float((md.dynamicness.cpu() - ref["dynamicness"]).abs().max()) <= 2e-6
    assert float((md.staticness.cpu() - ref["staticness"]).abs().max()) <= 2e-6
    near = (ref["dynamicness"] - thr).abs() < 2e-6
    assert torch.equal(md.is_dynamic.cpu() | near, ref["is_dynamic"] | near)
    near_s = near | ((ref["staticness"] - (1 - ref["staticness"] - ref["dynamicness"])).abs() < 4e-6)
    assert torch.equal(md.is_static.cpu() | near_s, ref["is_static"] | near_s)
    assert torch.equal(got.static_flow.cpu(), ref["pointwise_static_flow"])
    if aggregation:
        assert torch.equal(got.not_enough_points.cpu(), ref["not_enough_points"])
        dT = float((got.static_aggr_trafo.cpu() - ref["static_aggr_trafo"]).abs().max())
        assert dT <= 2e-5, dT
        dF = float((md.static_aggr_flow.cpu() - ref["static_aggr_flow"]).abs().max())
        assert dF <= 2e-4, dF
        # orthogonality of the rotation (U V^T, no determinant fix)
        R = got.static_aggr_trafo[:, :3, :3].cpu()
        assert float((R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max()) < 1e-12


def test_fused_decoder_equals_stock_torch_decoder(cuda):
    """Same interface, two implementations: the stock-PyTorch restatement on the GPU vs the fused kernels."""
    cfg, net_out, filled, pc, coors, valid = _inputs("T", [11, 12], 3)
    dec = _decoder(cfg)
    kw = dict(pc=pc.to(cuda), pointwise_voxel_coordinates=coors.to(cuda), pointwise_valid_mask=valid.to(cuda),
              filled_pillar_mask=filled.to(cuda), static_aggregation=True)
    thr = torch.tensor(0.5, device=cuda)
    with torch.no_grad():
        a = dec._forward_fused(net_out.to(cuda), thr, kw["pc"], kw["pointwise_voxel_coordinates"], kw["pointwise_valid_mask"],
                               kw["filled_pillar_mask"], True)
        b = forward_torch(dec, net_out.to(cuda), thr, **kw)
    for k in ("static_flow", "dynamic_flow", "dynamicness", "staticness", "aggregated_flow", "static_aggr_flow"):
        assert float((a[k] - b[k]).abs().max()) <= 2e-4, k
    ma, mb = a.modified_network_output, b.modified_network_output
    for k in ("disappearing_logit", "static_logit", "dynamic_logit", "ground_logit", "class_probs", "groundness",
              "masked_static_aggr_flow"):
        assert ma[k].shape == mb[k].shape, k
        assert float((ma[k] - mb[k]).abs().max()) <= 2e-4, k
    assert ma.is_ground.shape == mb.is_ground.shape and ma.is_ground.dtype == torch.bool
    for k in ("aggregated_flow", "static_flow"):
        assert float((a.dense_maps[k] - b.dense_maps[k]).abs().max()) <= 1e-6, k


def test_not_enough_points_branch(cuda):
    """< 3 positively weighted points: weights += 1e-7 (weighted_pc_alignment.py:31-35)."""
    cfg, net_out, filled, pc, coors, valid = _inputs("T", [21], 5)
    filled[:] = False  # staticness * filled == 0 everywhere -> no positive weight
    dec = _decoder(cfg)
    ref = SF.head_decoder(net_out, 0.5, pc, coors, valid, filled, dec.bev_extent, 1, with_static_aggregation=True)
    with torch.no_grad():
        got = dec(net_out.to(cuda), torch.tensor(0.5, device=cuda), pc=pc.to(cuda), pointwise_voxel_coordinates=coors.to(cuda),
                  pointwise_valid_mask=valid.to(cuda), filled_pillar_mask=filled.to(cuda), static_aggregation=True)
    assert bool(got.not_enough_points.all()) and bool(ref["not_enough_points"].all())
    assert float((got.static_aggr_trafo.cpu() - ref["static_aggr_trafo"]).abs().max()) <= 2e-4


@pytest.mark.parametrize("B,h,w,n", [(1, 8, 8, 8), (2, 32, 32, 8), (1, 80, 80, 8), (1, 23, 17, 4)])
def test_raft_output_kernel_matches_stock_ops(cuda, B, h, w, n):
    """SURVEY 8f.2 glue: upflow_n + uplogits_n + flip/scale + concat2network_output (raft_mod.py:216-266) in one kernel
    vs the stock PyTorch ops; <= 1e-5 relative (interpolation arithmetic may contract differently), min exact."""
    from liso_b200.slim import raft as R
    from liso_b200.slim.corr import uplogits_n, upflow_n

    g = torch.Generator().manual_seed(B * 100 + h)
    flow = (3.0 * torch.randn(B, 2, h, w, generator=g)).to(cuda)
    logits = (2.0 * torch.randn(B, 4, h, w, generator=g)).to(cuda)
    rr, rc = 0.875, 0.875
    res = torch.tensor([rr, rc], device=cuda)[None, :, None, None]
    flow_m = torch.flip(upflow_n(flow, n=n), dims=[1]) * res
    ref = R.concat2network_output(uplogits_n(logits, n=n), flow_m, flow_m)
    got = R.raft_output_fused(flow, logits, n, rr, rc)
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (B, h * n, w * n, 8)
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 1e-5 * scale
    # the carried minimum == min of the static / dynamic logits the kernel itself wrote
    key = int(got._slimb200_min_key.item()) & 0xffffffff
    bits = (key & 0x7fffffff) if key & 0x80000000 else (~key & 0xffffffff)
    val = np.frombuffer(np.uint32(bits).tobytes(), dtype=np.float32)[0]
    assert val == float(got[..., 1:3].min())


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 40, 56), (1, 64, 160, 160), (3, 96, 17, 23), (8, 32, 320, 320)])
@pytest.mark.parametrize("relu", [True, False])
def test_instance_norm_nhwc_kernel(cuda, B, C, H, W, relu):
    """Glue kernel vs F.instance_norm (+ ReLU) on a channels-last tensor: <= 2e-5 absolute on unit-scale data, also with
    a large mean (Welford + Chan merge, no E[x^2] - mean^2 cancellation)."""
    from liso_b200.slim import raft as R

    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(B, C, H, W, generator=g) * 1.7 + 25.0 * torch.randn(B, C, 1, 1, generator=g)).to(cuda)
    x = x.contiguous(memory_format=torch.channels_last)
    norm = torch.nn.InstanceNorm2d(C, eps=1e-3, affine=True).to(cuda)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
        ref = norm(x.contiguous())
        ref = torch.relu(ref) if relu else ref
        got = R.instance_norm_nhwc(norm, x, relu)
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert float((got - ref).abs().max()) <= 2e-5
