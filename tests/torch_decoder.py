"""TEST INFRASTRUCTURE ONLY -- a stock-PyTorch restatement of the output decoder with the interface of
``liso_b200.slim.slim.HeadDecoder`` (``head_decoder.py:410-496,517-717``, ``static_aggregation.py:8-110``,
``weighted_pc_alignment.py:10-80``): a second checker for the fused kernels beside ``oracle/slim_forward.py::head_decoder``.
The product module has no such path (CPU tensors raise)."""
import numpy as np
import torch
import torch.nn.functional as F

from liso_b200.config import AttrDict


def _grid_to_points(grid, coors, valid, default):
    """``batched_grid_data_to_pointwise_data`` (``static_aggregation.py:8-31``), without mutating ``coors``."""
    coors = torch.where(valid[..., None], coors, torch.zeros_like(coors)).long()
    bidx = torch.arange(valid.shape[0], device=grid.device)[:, None].expand(-1, valid.shape[1])
    out = grid[bidx, coors[..., 0], coors[..., 1]]
    return torch.where(valid[..., None], out, torch.full_like(out, default))


def _weighted_kabsch(cloud_t0, cloud_t1, weights):
    """``weighted_pc_alignment.py:10-80`` (no epsilon) + ``torch_symm_ortho`` (U @ Vh, no det fix)."""
    not_enough = (weights > 0).sum() < 3
    weights = torch.where(not_enough, weights + 1e-7, weights)
    cum = weights.sum(dim=-1)
    mx = (cloud_t0 * weights[..., None]).sum(dim=0) / cum
    my = (cloud_t1 * weights[..., None]).sum(dim=0) / cum
    S = ((cloud_t1 - my[None]) * weights[..., None]).T @ (cloud_t0 - mx[None]) / cum
    U, _, Vh = torch.linalg.svd(S.to(torch.double))
    R = U @ Vh
    t = my.to(torch.double) - R @ mx.to(torch.double)
    T = torch.eye(4, dtype=torch.double, device=R.device)
    T[:3, :3] = R
    T[:3, 3] = t
    return T, not_enough


def forward_torch(dec, network_output, dynamicness_threshold, *, pc, pointwise_voxel_coordinates, pointwise_valid_mask,
              filled_pillar_mask, static_aggregation: bool = True):
    """Stock-PyTorch restatement (CPU tensors, host-logic tests)."""
    fs = dec.cfg.model.u_net.final_scale
    coors = torch.div(pointwise_voxel_coordinates, fs, rounding_mode="trunc")
    filled = filled_pillar_mask[..., None]
    o = network_output
    static_logit, dynamic_logit = o[..., 1:2], o[..., 2:3]
    ones = torch.ones_like(static_logit)
    # ground "off": global min of the live logits - 100 (head_decoder.py:919-935), before masking
    ground_logit = torch.min(torch.cat([static_logit, dynamic_logit], dim=0)) - 100.0 * ones
    neg = -100.0 * ones
    md = AttrDict()
    md.disappearing_logit = neg
    md.static_logit = torch.where(filled, static_logit, 0.0 * ones)
    md.dynamic_logit = torch.where(filled, dynamic_logit, neg)
    md.ground_logit = torch.where(filled, ground_logit, neg)
    md.static_flow = torch.where(filled, o[..., 4:6], torch.zeros_like(o[..., 4:6]))
    md.dynamic_flow = torch.where(filled, o[..., 6:8], torch.zeros_like(o[..., 6:8]))
    md.class_logits = torch.cat([md.static_logit, md.dynamic_logit, md.ground_logit], dim=-1)
    md.class_probs = F.softmax(md.class_logits, dim=-1)
    md.staticness, md.dynamicness, md.groundness = (md.class_probs[..., k] for k in range(3))
    md.is_dynamic = md.dynamicness >= dynamicness_threshold
    md.is_static = (md.staticness >= md.groundness) & (~md.is_dynamic)
    md.is_ground = ~(md.is_static | md.is_dynamic)

    zeros1 = torch.zeros_like(md.static_flow[..., :1])
    static3 = torch.cat([md.static_flow, zeros1], dim=-1)
    dynamic3 = torch.cat([md.dynamic_flow, zeros1], dim=-1)
    valid = pointwise_valid_mask
    ret = AttrDict()
    ret.static_flow = _grid_to_points(static3, coors, valid, 0.0)
    ret.dynamic_flow = _grid_to_points(dynamic3, coors, valid, 0.0)
    ret.dynamicness = _grid_to_points(md.dynamicness[..., None], coors, valid, 0.0)[..., 0]
    ret.staticness = _grid_to_points(md.staticness[..., None], coors, valid, 0.0)[..., 0]
    aggregated = torch.where(md.is_static[..., None], static3, dynamic3 * (1.0 - md.groundness[..., None]))
    ret.aggregated_flow = _grid_to_points(aggregated, coors, valid, 0.0)
    ret.dense_maps = AttrDict(aggregated_flow=aggregated, static_flow=static3)
    ret.dynamicness_threshold = dynamicness_threshold
    if static_aggregation:
        weight_map = md.staticness * filled[..., 0].float()
        pt_w = _grid_to_points(weight_map[..., None], coors, valid, 0.0)[..., 0]
        shape = o.shape[1:3]
        ext = np.asarray(dec.bev_extent, dtype=np.float64)
        ctr = np.stack(np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij"), axis=-1) + 0.5
        ctr = ctr / np.asarray(shape) * (ext[2:] - ext[:2]) + ext[:2]
        grid_h = torch.from_numpy(
            np.concatenate([ctr, np.zeros_like(ctr[..., :1]), np.ones_like(ctr[..., :1])], axis=-1)).to(o.device)
        flows, Ts, neps = [], [], []
        eye = torch.eye(4, dtype=torch.float64, device=o.device)
        for b in range(o.shape[0]):
            m = valid[b]
            T, nep = _weighted_kabsch(pc[b][m][..., :3], (pc[b][..., :3] + ret.static_flow[b])[m], pt_w[b][m])
            flows.append(torch.einsum("ij,hwj->hwi", T - eye, grid_h)[..., 0:2].float())
            Ts.append(T)
            neps.append(nep)
        md.static_aggr_flow = torch.stack(flows, 0)
        md.masked_static_aggr_flow = torch.where(filled, md.static_aggr_flow, torch.zeros_like(md.static_aggr_flow))
        ret.static_aggr_flow = _grid_to_points(torch.cat([md.static_aggr_flow, zeros1], dim=-1), coors, valid, 0.0)
        ret.static_aggr_trafo = torch.stack(Ts, 0)
        ret.not_enough_points = torch.stack(neps, 0)
    else:
        ret.not_enough_points = torch.zeros((o.shape[0],), dtype=torch.bool, device=o.device)
    ret.modified_network_output = md
    return ret


