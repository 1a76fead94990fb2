"""TEST INFRASTRUCTURE ONLY -- CPU port of the reference ``SLIM.forward`` (functional, fp32).

Checker and CPU baseline for the end-to-end path; never imported by ``liso_b200/``.
Weights come in as a plain state dict with the reference's keys
(``raft_network.pp_layer...``, ``raft_network.fnet...`` etc., ``experiment.py:221-223``), so the
port shares no module code with the product.  Pinned against the unmodified reference run in
the authoring container (``oracle/gen_golden.py`` -> ``tests/golden/slim_forward_tiny.npz``;
``tests/test_oracle_vs_reference.py`` when ``/root/reference`` is present).

Follows:
* ``liso/slim/model/slim.py:44-156``           SLIM.forward
* ``liso/slim/model/raft_mod.py:82-266``       RAFT.forward / predict_single_flow_map_and_classes
* ``liso/slim/model/extractor.py:5-71,211-297`` SmallEncoder / ResidualBlock
* ``liso/slim/model/update.py:6-164``          SmallUpdateBlock / ConvGRU / SmallMotionEncoder
* ``liso/slim/model/head_decoder.py:66-496,517-717`` HeadDecoder (default ``output_modification``)
* ``liso/slim/slim_loss/static_aggregation.py:8-110``, ``weighted_pc_alignment.py:10-80``,
  ``liso/torch_symm_ortho/__init__.py:50-69``
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import slim_oracle as O

SD = Dict[str, torch.Tensor]


# ---------------------------------------------------------------- extractor.py
def _norm(x, sd: SD, key: str, norm_fn: str):
    if norm_fn == "none":
        return x
    assert norm_fn == "instance_affine"
    return F.instance_norm(x, weight=sd[key + ".weight"], bias=sd[key + ".bias"], eps=1e-3)


def _residual_block(x, sd: SD, p: str, norm_fn: str, stride: int):
    """extractor.py:58-68.  The downsample branch exists whenever the *stage* changes width or
    stride, also for the second block of such a stage (``dummy_in_filters`` quirk, :19-21,264-279)."""
    y = F.relu(_norm(F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=stride, padding=1), sd, p + ".norm1", norm_fn))
    y = F.relu(_norm(F.conv2d(y, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1), sd, p + ".norm2", norm_fn))
    if (p + ".downsample.0.weight") in sd:
        x = F.conv2d(x, sd[p + ".downsample.0.weight"], sd[p + ".downsample.0.bias"], stride=stride)
        x = _norm(x, sd, p + ".norm3", norm_fn)
    return F.relu(x + y)


def small_encoder(x, sd: SD, p: str, norm_fn: str):
    """extractor.py:281-297"""
    x = F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=2, padding=3)
    x = F.relu(_norm(x, sd, p + ".norm1", norm_fn))
    for layer, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        x = _residual_block(x, sd, f"{p}.{layer}.0", norm_fn, stride)
        x = _residual_block(x, sd, f"{p}.{layer}.1", norm_fn, 1)
    return F.conv2d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])


# ---------------------------------------------------------------- update.py
def _conv(x, sd: SD, key: str, padding: int):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], padding=padding)


def update_block(net, inp, corr, flow, logits, sd: SD, p: str):
    """update.py:71-93 (motion encoder), :30-38 (ConvGRU), :128-164 (heads)."""
    me = p + ".motion_encoder"
    c = F.relu(_conv(corr, sd, me + ".conv_stat_corr1", 0))
    f = F.relu(_conv(flow, sd, me + ".conv_flow1", 3))
    f = F.relu(_conv(f, sd, me + ".conv_flow2", 1))
    lg = F.relu(_conv(logits, sd, me + ".conv_class1", 3))
    lg = F.relu(_conv(lg, sd, me + ".conv_class2", 1))
    out = F.relu(_conv(torch.cat([c, f, lg], dim=1), sd, me + ".conv", 1))
    motion = torch.cat([out, lg, f], dim=1)
    x = torch.cat([inp, motion], dim=1)
    hx = torch.cat([net, x], dim=1)
    z = torch.sigmoid(_conv(hx, sd, p + ".gru.convz", 1))
    r = torch.sigmoid(_conv(hx, sd, p + ".gru.convr", 1))
    q = torch.tanh(_conv(torch.cat([r * net, x], dim=1), sd, p + ".gru.convq", 1))
    net = (1 - z) * net + z * q
    dflow = _conv(F.relu(_conv(net, sd, p + ".static_flow_head.conv1", 1)), sd, p + ".static_flow_head.conv2", 1)
    dlogits = _conv(F.relu(_conv(net, sd, p + ".classification_head.conv1", 1)), sd, p + ".classification_head.conv2", 1)
    return net, dflow, dlogits


# ---------------------------------------------------------------- raft_mod.py
def raft_direction(img_t0, fmap_t0, fmap_t1, sd: SD, cfg, metres_per_px: float, corr_hook=None) -> List[torch.Tensor]:
    """raft_mod.py:124-259 -> list over iterations of (B,H,W,8) [4 logits | static xy | dynamic xy]."""
    m = cfg.SLIM.model
    ds = m.feature_downsampling_factor
    b, _, H, W = img_t0.shape
    h, w = H // ds, W // ds
    coords0 = O.coords_grid(b, h, w)
    coords1 = O.coords_grid(b, h, w)
    logits = torch.zeros((b, 4, h, w), dtype=torch.float32)
    pyramid = O.corr_pyramid(fmap_t0, fmap_t1, m.corr_cfg.num_levels)
    if corr_hook is not None:
        pyramid = corr_hook(pyramid)
    cnet = small_encoder(img_t0, sd, "raft_network.cnet", "none")
    net, inp = torch.split(cnet, [96, 64], dim=1)
    net, inp = torch.tanh(net), torch.relu(inp)
    adapter = torch.tensor([metres_per_px, metres_per_px], dtype=torch.float32)[None, :, None, None]
    outs = []
    for _ in range(m.num_iters):
        corr = O.corr_lookup(pyramid, coords1, m.corr_cfg.search_radius)
        flow = coords1 - coords0
        net, dflow, dlogits = update_block(net, inp, corr, flow, logits, sd, "raft_network.update_block")
        coords1 = coords1 + dflow
        logits = logits + dlogits
        up = ds * F.interpolate(coords1 - coords0, size=(ds * h, ds * w), mode="bilinear", align_corners=True)
        flow_m = torch.flip(up, dims=[1]) * adapter  # raft_mod.py:262-266
        up_logits = F.interpolate(logits, size=(ds * h, ds * w), mode="bilinear", align_corners=True)
        outs.append(torch.cat([up_logits, flow_m, flow_m], dim=1).permute(0, 2, 3, 1))
    return outs


# ---------------------------------------------------------------- head_decoder.py
def voxel_center_coords_m(bev_extent_m: np.ndarray, shape) -> np.ndarray:
    """head_decoder.py:498-514"""
    v = np.stack(np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij"), axis=-1) + 0.5
    v /= shape
    v *= bev_extent_m[2:] - bev_extent_m[:2]
    v += bev_extent_m[:2]
    return v


def grid_to_points(grid, coors, valid, default):
    """static_aggregation.py:8-31 (the in-place zeroing of invalid coords is done on a copy)."""
    coors = coors.clone()
    coors[~valid] = 0
    bidx = torch.arange(valid.shape[0])[:, None].expand(-1, valid.shape[1])
    out = grid[bidx, coors[..., 0].long(), coors[..., 1].long()]
    out[~valid] = default
    return out


def weighted_pc_alignment(cloud_t0, cloud_t1, weights):
    """weighted_pc_alignment.py:10-80 with ``use_epsilon_on_weights=False``; R = U @ Vh without
    determinant fix (torch_symm_ortho/__init__.py:68-69)."""
    not_enough = (weights > 0).sum() < 3
    if not_enough:
        weights = weights + 1e-7
    cum = weights.sum(dim=-1)
    mx = (cloud_t0 * weights[..., None]).sum(dim=0) / cum
    my = (cloud_t1 * weights[..., None]).sum(dim=0) / cum
    Xc, Yc = cloud_t0 - mx[None, :], cloud_t1 - my[None, :]
    S = (Yc * weights[..., None]).T @ Xc / cum
    U, _, Vh = torch.linalg.svd(S.to(torch.double))
    Rm = U @ Vh
    t = my.to(torch.double) - Rm @ mx.to(torch.double)
    Rm = torch.cat([Rm, torch.zeros((1, 3), dtype=Rm.dtype)], dim=0)
    t = torch.cat([t, torch.ones((1,), dtype=t.dtype)], dim=-1)
    return torch.cat([Rm, t[:, None]], dim=-1), not_enough


def head_decoder(net_out, dyn_threshold, pc, pillar_coors, valid, filled, bev_extent, final_scale: int = 1,
                 with_static_aggregation: bool = True):
    """head_decoder.py:410-496 with the default ``output_modification`` (disappearing off, static/dynamic
    logits from the net, ground off; liso_config.yml:303-310).  Returns the tensors the export and
    the parity tests read."""
    coors_fs = torch.div(pillar_coors, final_scale, rounding_mode="trunc")
    filled = filled[..., None]
    static_logit, dynamic_logit = net_out[..., 1:2], net_out[..., 2:3]
    static_flow, dynamic_flow = net_out[..., 4:6], net_out[..., 6:8]
    ones = torch.ones_like(static_logit)
    disappearing_logit = -100 * ones
    ground_logit = torch.min(torch.cat([static_logit, dynamic_logit], dim=0)) - 100.0 * ones
    # mask non-filled pillars (head_decoder.py:568-609)
    disappearing_logit = torch.where(filled, disappearing_logit, -100.0 * ones)
    static_logit = torch.where(filled, static_logit, 0.0 * ones)
    dynamic_logit = torch.where(filled, dynamic_logit, -100.0 * ones)
    ground_logit = torch.where(filled, ground_logit, -100.0 * ones)
    static_flow = torch.where(filled, static_flow, torch.zeros_like(static_flow))
    dynamic_flow = torch.where(filled, dynamic_flow, torch.zeros_like(dynamic_flow))
    class_logits = torch.cat([static_logit, dynamic_logit, ground_logit], dim=-1)
    class_probs = F.softmax(class_logits, dim=-1)
    staticness, dynamicness, groundness = class_probs[..., 0], class_probs[..., 1], class_probs[..., 2]
    is_dynamic = dynamicness >= dyn_threshold
    is_static = (staticness >= groundness) & (~is_dynamic)
    out = dict(static_flow=static_flow, dynamic_flow=dynamic_flow, dynamicness=dynamicness, staticness=staticness,
               class_logits=class_logits, is_static=is_static, is_dynamic=is_dynamic)
    flow3 = torch.cat([static_flow, torch.zeros_like(static_flow[..., :1])], dim=-1)
    out["pointwise_static_flow"] = grid_to_points(flow3, coors_fs, valid, 0.0)
    if with_static_aggregation:
        # static_aggregation.py:34-110
        weight_map = staticness * filled[..., 0].float()
        pt_flow = out["pointwise_static_flow"]
        pt_w = grid_to_points(weight_map[..., None], coors_fs, valid, 0.0)[..., 0]
        centers = torch.from_numpy(voxel_center_coords_m(np.array(bev_extent), net_out.shape[1:3]))
        pc0_grid = torch.cat([centers, torch.zeros_like(centers[..., :1]), torch.ones_like(centers[..., :1])], dim=-1)
        aggr, Ts, neps = [], [], []
        for b in range(net_out.shape[0]):
            T, nep = weighted_pc_alignment(pc[b][valid[b]][..., :3], (pc[b][..., :3] + pt_flow[b])[valid[b]], pt_w[b][valid[b]])
            aggr.append(torch.einsum("ij,hwj->hwi", T - torch.eye(4, dtype=torch.float64), pc0_grid)[..., 0:2].float())
            Ts.append(T)
            neps.append(nep)
        out["static_aggr_flow"] = torch.stack(aggr, 0)
        out["static_aggr_trafo"] = torch.stack(Ts, 0)
        out["not_enough_points"] = torch.stack(neps, 0)
    return out


# ---------------------------------------------------------------- slim.py
def slim_forward(sd: SD, cfg, sample_t0, sample_t1, bn_training: bool = False, decode_all_iterations: bool = True,
                 corr_hook=None):
    """slim.py:44-156 on CPU.  Returns (preds_fw, preds_bw, aux) where preds_* are lists (one dict per
    decoded iteration; only the last when ``decode_all_iterations`` is False) and aux carries the
    pillar-encoder intermediates used by stage-level parity tests."""
    d = cfg.data
    pp = "raft_network.pp_layer.pts_voxel_encoder.pfn_layers.0."
    params = dict(linear_weight=sd[pp + "linear.weight"], bn_weight=sd[pp + "norm.weight"], bn_bias=sd[pp + "norm.bias"],
                  running_mean=sd[pp + "norm.running_mean"], running_var=sd[pp + "norm.running_var"])
    enc, imgs = [], []
    for s in (sample_t0, sample_t1):
        pts = [np.asarray(p, dtype=np.float32) for p in s["pcl_full_no_ground_ta"]]
        e = O.pillar_encoder_forward(pts, params, d.bev_range_m, d.img_grid_size, d.z_pillar_cutoff_value, bn_training)
        if bn_training:  # running stats are mutated between the two frames (Q4)
            params = dict(params, running_mean=e["running_mean"], running_var=e["running_var"])
        enc.append(e)
        imgs.append(e["canvas"])
    fm0 = small_encoder(imgs[0], sd, "raft_network.fnet", cfg.SLIM.model.raft_fnet_norm)
    fm1 = small_encoder(imgs[1], sd, "raft_network.fnet", cfg.SLIM.model.raft_fnet_norm)
    mpp = float(d.bev_range_m[0]) / d.img_grid_size[0] * cfg.SLIM.model.u_net.final_scale
    outs_fw = raft_direction(imgs[0], fm0, fm1, sd, cfg, mpp, corr_hook)
    outs_bw = raft_direction(imgs[1], fm1, fm0, sd, cfg, mpp, corr_hook)
    half = 0.5 * np.array(d.bev_range_m)
    bev_extent = np.concatenate([-half, half], axis=0)
    thr = sd.get("moving_dynamicness_threshold.start_value", torch.tensor(0.5))
    preds = ([], [])
    its = range(len(outs_fw)) if decode_all_iterations else [len(outs_fw) - 1]
    for it in its:
        for k, (o, s, e) in enumerate(((outs_fw[it], sample_t0, enc[0]), (outs_bw[it], sample_t1, enc[1]))):
            filled = torch.squeeze(e["occupancy"] > 0.5, dim=1)
            preds[k].append(head_decoder(o, thr, s["pcl_ta"]["pcl"], s["pcl_ta"]["pillar_coors"], s["pcl_ta"]["pcl_is_valid"],
                                         filled, bev_extent, cfg.SLIM.model.u_net.final_scale))
    aux = dict(enc=enc, fmaps=(fm0, fm1), net_out_fw=outs_fw, net_out_bw=outs_bw)
    return preds[0], preds[1], aux
