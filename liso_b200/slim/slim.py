"""SLIM top module around the B200 hot path (reference ``liso/slim/model/slim.py``).

``SLIM(cfg, num_train_samples).forward(sample_data_t0, sample_data_t1, summaries)`` keeps the
reference signature, sub-module names (``raft_network``, ``head_decoder_fw/bw``,
``moving_dynamicness_threshold``) and state-dict keys.  The decoder below is the forward /
export part of ``HeadDecoder`` (``head_decoder.py:66-496``) for the released configuration
(``output_modification`` defaults, ``liso_config.yml:303-310``), in stock PyTorch.

Two switches exist that the reference does not have; both default to the reference behaviour:
``decode_iterations`` ("all" | "last") and ``static_aggregation`` (bool).  The flow export only
reads ``static_flow`` / ``dynamicness`` of the last iteration (``experiment.py:391-399``), so the
export driver sets ("last", False) and skips 11 of 12 decoder calls and every fp64 3x3 SVD.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
from torch import nn

from .. import _lib as _lib_module
from ..config import AttrDict
from .raft import RAFT


def get_network_input_pcls(cfg, sample_data, time_key: str, to_device=None) -> List[torch.Tensor]:
    """``liso/kabsch/main_utils.py:247-261`` (+ raw scans from ``liso_b200.datasets.preprocess_scans``: the pillar
    encoder removes the ground in-kernel, so the raw cloud stands in for ``pcl_full_no_ground``)."""
    key = ("pcl_full_w_ground_%s" if cfg.data.use_ground_for_network else "pcl_full_no_ground_%s") % time_key
    if key not in sample_data and sample_data.get("raw_scan", False):
        key = "pcl_full_w_ground_%s" % time_key
    if to_device:
        return [el.to(to_device, non_blocking=True) for el in sample_data[key]]
    return sample_data[key]


def _cfg_get(cfg, path, default=False):
    """cfg.a.b.c for attribute- or item-style configs; `default` when a key is missing."""
    cur = cfg
    for k in path:
        try:
            cur = cur[k] if isinstance(cur, dict) else getattr(cur, k)
        except (KeyError, AttributeError):
            return default
    return cur


class MovingAverageThreshold(nn.Module):
    """State + ``value()`` of ``slim_loss/movavg_cls_threshold.py`` (the update rule is training-only)."""

    def __init__(self, num_train_samples: int, num_moving: int, num_still=None, resolution: int = 100000,
                 start_value: float = 0.5, value_range=(0.0, 1.0)):
        super().__init__()
        assert num_still is None, "supervised SLIM training is out of scope"
        self.value_range = (value_range[0], value_range[1] - value_range[0])
        self.resolution = resolution
        total = num_moving
        assert num_train_samples > 0
        update_weight = 1.0 / min(2.0 * total, 5_000.0 * (total / num_train_samples))
        self.register_buffer("start_value", torch.tensor(start_value, dtype=torch.float))
        self.register_buffer("update_weight", torch.tensor(update_weight, dtype=torch.double))
        self.register_buffer("bias_counter", torch.zeros((), dtype=torch.double))
        self.register_buffer("moving_average_importance", torch.zeros((resolution,), dtype=torch.float))

    def value(self):
        """Device scalar.  The reference branches on ``bias_counter > 0`` (a host sync in the middle of every forward);
        the buffers only change in training, so the result is cached per buffer version."""
        key = (self.bias_counter._version, self.moving_average_importance._version, self.start_value._version,
               self.bias_counter.data_ptr(), self.start_value.device)
        if getattr(self, "_value_cache", (None, None))[0] != key:
            self._value_cache = (key, self._value())
        return self._value_cache[1]

    def _value(self):
        if self.bias_counter > 0.0:
            mai = self.moving_average_importance
            cum = torch.cat([torch.zeros((1,), dtype=mai.dtype, device=mai.device), torch.cumsum(mai, 0)], dim=0)
            idx = torch.mean(torch.where(torch.min(cum) == cum)[0].to(torch.float))
            return self.value_range[0] + idx * self.value_range[1] / self.resolution
        return self.start_value


class HeadDecoder(nn.Module):
    def __init__(self, cfg, name, bev_extent):
        super().__init__()
        self.cfg = cfg
        self.name = name
        self.bev_extent = bev_extent
        om = cfg.model.output_modification
        if not (om.disappearing_logit is False and om.static_logit == "net" and om.dynamic_logit == "net"
                and om.ground_logit is False and om.static_flow == "net" and om.dynamic_flow == "net"):
            raise NotImplementedError("only the released SLIM output_modification is implemented")
        # switches of the reference decoder that change its arithmetic (head_decoder.py:463-480, static_aggregation.py:62-68):
        # the fused kernel implements the released setting only -- anything else must fail loudly, not decode differently
        flags = {"model.use_static_aggr_flow_for_aggr_flow": _cfg_get(cfg, ("model", "use_static_aggr_flow_for_aggr_flow")),
                 "model.dynamic_flow_is_non_rigid_flow": _cfg_get(cfg, ("model", "dynamic_flow_is_non_rigid_flow")),
                 "losses.unsupervised.use_epsilon_for_weighted_pc_alignment":
                     _cfg_get(cfg, ("losses", "unsupervised", "use_epsilon_for_weighted_pc_alignment"))}
        for name, value in flags.items():
            if value:
                raise NotImplementedError("SLIM.%s = %r: only False (the released configuration) is implemented" % (name, value))

    def concat2network_output(self, *, logits, static_flow, dynamic_flow, weight_logits_for_static_aggregation=None):
        assert weight_logits_for_static_aggregation is None
        return torch.cat([logits, static_flow, dynamic_flow], dim=1).permute(0, 2, 3, 1)

    def forward(self, network_output, dynamicness_threshold, *, pc, pointwise_voxel_coordinates, pointwise_valid_mask,
                filled_pillar_mask, odom=None, inv_odom=None, summaries=None, static_aggregation: bool = True, **_):
        from .. import _lib

        _lib.require_cuda(network_output, pc, pointwise_voxel_coordinates, pointwise_valid_mask, filled_pillar_mask)  # no CPU path
        return self._forward_fused(network_output, dynamicness_threshold, pc, pointwise_voxel_coordinates,
                                   pointwise_valid_mask, filled_pillar_mask, static_aggregation)

    @_lib_module.on_device_of_args
    def _forward_fused(self, o, thr, pc, coors, valid, filled, static_aggregation):
        """One call into ``slimb200_head_decode``: five launches, no host sync; the tensors of the
        reference's result are channel slices of two packed buffers."""
        import ctypes as C

        from .. import _lib

        lib = _lib.load()
        dev = o.device
        B, H, W, ch = o.shape
        assert ch == 8, o.shape
        if o.dtype != torch.float32 or not o.is_contiguous():
            o = o.float().contiguous()
        N = int(valid.shape[1])
        p = _lib.DecodeParams()
        p.batch, p.H, p.W, p.n_points = B, H, W, N
        pc = pc.float().contiguous()
        p.pc_stride = int(pc.shape[-1])
        p.final_scale = int(self.cfg.model.u_net.final_scale)
        p.static_aggregation = 1 if static_aggregation else 0
        ext = [float(v) for v in self.bev_extent]
        p.ext_min_x, p.ext_min_y, p.ext_max_x, p.ext_max_y = ext
        coors = coors.to(torch.int32).contiguous()
        valid_u8 = valid.contiguous().view(torch.uint8)
        filled_u8 = filled.contiguous().view(torch.uint8)
        thr_t = torch.as_tensor(thr, dtype=torch.float32, device=dev).reshape(1)
        bev = torch.empty((B, H, W, _lib.DECODE_BEV_CHANNELS), dtype=torch.float32, device=dev)
        aggr = torch.empty((B, H, W, 4), dtype=torch.float32, device=dev) if static_aggregation else None
        cls = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
        pts = torch.empty((B, N, _lib.DECODE_POINT_CHANNELS), dtype=torch.float32, device=dev)
        trafo = torch.empty((B, 4, 4), dtype=torch.float64, device=dev)
        nep = torch.zeros((B,), dtype=torch.uint8, device=dev)
        ws_bytes = lib.slimb200_head_decode_workspace_bytes(C.byref(p))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        min_key = getattr(o, "_slimb200_min_key", None)  # set by RAFT's fused output kernel: spares a pass over `o`
        _lib.check(lib.slimb200_head_decode(
            o.data_ptr(), min_key.data_ptr() if min_key is not None else None, filled_u8.data_ptr(), pc.data_ptr(), coors.data_ptr(), valid_u8.data_ptr(), thr_t.data_ptr(),
            C.byref(p), bev.data_ptr(), aggr.data_ptr() if aggr is not None else None, cls.data_ptr(), pts.data_ptr(),
            trafo.data_ptr(), nep.data_ptr(),
            ws.data_ptr(), ws.numel(), _lib.current_stream_ptr()))
        return self._assemble(bev, aggr, cls, pts, trafo, nep, thr, static_aggregation)

    @staticmethod
    def _assemble(bev, aggr, cls, pts, trafo, nep, thr, static_aggregation):
        """The reference's result structure as channel slices of the kernel's packed buffers (kept in `_packed`)."""
        dev, B = bev.device, bev.shape[0]
        clsb = cls.view(torch.bool)
        md = AttrDict()
        md.disappearing_logit = bev[..., 0:1]
        md.static_logit, md.dynamic_logit, md.ground_logit = bev[..., 1:2], bev[..., 2:3], bev[..., 3:4]
        md.class_logits = bev[..., 1:4]
        md.class_probs = bev[..., 4:7]
        md.staticness, md.dynamicness, md.groundness = bev[..., 4], bev[..., 5], bev[..., 6]
        md.static_flow, md.dynamic_flow = bev[..., 7:9], bev[..., 10:12]
        md.is_dynamic, md.is_static, md.is_ground = clsb[..., 0], clsb[..., 1], clsb[..., 2]
        ret = AttrDict()
        ret.static_flow, ret.dynamic_flow = pts[..., 0:3], pts[..., 3:6]
        ret.dynamicness, ret.staticness = pts[..., 6], pts[..., 7]
        ret.aggregated_flow = pts[..., 8:11]
        ret.dense_maps = AttrDict(aggregated_flow=bev[..., 13:16], static_flow=bev[..., 7:10])
        ret.dynamicness_threshold = thr
        if static_aggregation:
            md.static_aggr_flow = aggr[..., 0:2]
            md.masked_static_aggr_flow = aggr[..., 2:4]
            ret.static_aggr_flow = pts[..., 11:14]
            ret.static_aggr_trafo = trafo
            ret.not_enough_points = nep.view(torch.bool)
        else:
            ret.not_enough_points = torch.zeros((B,), dtype=torch.bool, device=dev)
        ret.modified_network_output = md
        ret._packed = (bev, aggr, cls, pts, trafo, nep, static_aggregation)
        return ret


class SLIM(nn.Module):
    def __init__(self, cfg, num_train_samples: int = 15000, decode_iterations: str = "all",
                 static_aggregation: bool = True):
        super().__init__()
        self.cfg = cfg
        self.slim_cfg = cfg.SLIM
        half = 0.5 * np.array(cfg.data.bev_range_m)
        bev_extent = np.concatenate([-half, half], axis=0)
        self.head_decoder_fw = HeadDecoder(self.slim_cfg, bev_extent=bev_extent, name="head_decoder_forward")
        self.head_decoder_bw = HeadDecoder(self.slim_cfg, bev_extent=bev_extent, name="head_decoder_backward")
        assert self.slim_cfg.phases.train.mode == "unsupervised"
        self.moving_dynamicness_threshold = MovingAverageThreshold(num_train_samples, num_moving=621013971)
        self.raft_network = RAFT(cfg, self.head_decoder_fw, self.head_decoder_bw)
        assert decode_iterations in ("all", "last")
        self.decode_iterations = decode_iterations
        self.raft_network.output_iterations = decode_iterations
        self.static_aggregation = static_aggregation
        # decode each network output as soon as it exists (a side branch of the CUDA graph) instead of after the network
        self.decode_as_sink = True
        # In CUDA-graph mode the decoded outputs live in the graph's static buffers, which the NEXT forward overwrites.
        # False (default): forward returns copies the caller may keep -- the reference's export holds the t0->t1 predictions
        # across two further model() calls (experiment.py:386-456).  True: forward returns views of the static buffers, valid
        # until the next forward (no copy of ~0.4 GB per decoded iteration); `ExportPipeline` works this way and takes packed
        # copies of the four tensors it exports.
        self.outputs_alias_static_buffers = False
        self._dec_static, self._dec_ctx, self._dec_n, self._sink_preds, self._sink_filled = {}, {}, {}, {}, {}
        self._graph_preds = None  # (graph key, decoded static outputs of the captured pass)

    # ---- output decoding as a sink of the network (SURVEY 8f.1 + 8f.2) ------------------------------------------
    # The reference decodes every iteration's network output after the network has run (slim.py:77-148).  Here the
    # decoder is handed to RAFT as `output_sink`: each (B, H, W, 8) output is decoded right behind the kernel that wrote
    # it -- inline in eager mode, and as a SIDE BRANCH of the CUDA graph when the network is graph-captured, where the
    # HBM-bound decode kernels overlap the latency-bound GRU iterations that follow.  In graph mode the decoder's inputs
    # (points, pillar coordinates, validity, threshold) live in grow-only static buffers that are refreshed before every
    # replay, and the returned predictions are views of the graph's static outputs: valid until the next forward.
    POINT_CAPACITY_STEP = 8192

    def _sink_begin(self):
        self._sink_preds = {}
        self._sink_filled = {}

    def _sink(self, direction, it, net_out, occupancy):
        pc, coors, valid, thr = self._dec_ctx[direction]
        dec = self.head_decoder_fw if direction % 2 == 0 else self.head_decoder_bw
        # one mask per direction and pass, not one per decoded iteration (per direction: the sinks of two directions are
        # different branches of the CUDA graph, a mask shared between them would need an edge of its own)
        fkey = (direction, occupancy.data_ptr(), tuple(occupancy.shape))
        filled = self._sink_filled.get(fkey)
        if filled is None:
            filled = self._sink_filled[fkey] = torch.squeeze(occupancy > 0.5, dim=1)
        self._sink_preds[(direction, it)] = dec(net_out, thr, pc=pc, pointwise_voxel_coordinates=coors,
                                                pointwise_valid_mask=valid, filled_pillar_mask=filled,
                                                static_aggregation=self.static_aggregation)

    def _stage_decode_inputs(self, samples, dev, thr, static: bool):
        """Decoder inputs per direction; `static`: copies into grow-only buffers the captured graph points at (padding
        rows are invalid points).  Returns the graph key part that changes when a buffer is re-allocated."""
        self._dec_ctx, self._dec_n = {}, {}
        keys = []
        for k, sample in enumerate(samples):
            pt = sample["pcl_ta"]
            pcl, coors, valid = pt["pcl"], pt["pillar_coors"], pt["pcl_is_valid"]
            B, N = int(valid.shape[0]), int(valid.shape[1])
            self._dec_n[k] = N
            if not static:
                self._dec_ctx[k] = (pcl.to(dev, non_blocking=True), coors.to(dev, non_blocking=True),
                                    valid.to(dev, non_blocking=True), thr)
                continue
            C = int(pcl.shape[-1])
            st = self._dec_static.get(k)
            if st is None or st["pc"].shape[0] != B or st["pc"].shape[2] != C or st["pc"].shape[1] < N or st["pc"].device != dev:
                step = self.POINT_CAPACITY_STEP
                cap = max((N + step - 1) // step * step, step, 0 if st is None else int(st["pc"].shape[1]))
                st = {"pc": torch.full((B, cap, C), float("nan"), dtype=torch.float32, device=dev),
                      "coors": torch.full((B, cap, 2), -1, dtype=torch.int32, device=dev),
                      "valid": torch.zeros((B, cap), dtype=torch.bool, device=dev),
                      "thr": torch.zeros((1,), dtype=torch.float32, device=dev)}
                self._dec_static[k] = st
            st["pc"][:, :N].copy_(pcl, non_blocking=True)
            st["coors"][:, :N].copy_(coors, non_blocking=True)
            st["valid"][:, :N].copy_(valid, non_blocking=True)
            if N < st["valid"].shape[1]:
                st["valid"][:, N:].zero_()
            st["thr"].copy_(torch.as_tensor(thr, dtype=torch.float32).reshape(1), non_blocking=True)
            self._dec_ctx[k] = (st["pc"], st["coors"], st["valid"], st["thr"])
            keys += [st["pc"].data_ptr(), st["coors"].data_ptr(), st["valid"].data_ptr(), st["thr"].data_ptr(),
                     int(st["pc"].shape[1])]
        return tuple(keys)

    def _sink_results(self, net, static: bool, captures_before: int):
        """The decoded outputs that belong to the call `net` just served.  Eager call: what the sink produced during it.
        Graph call: the sink only runs while a graph is captured, so the results of THAT pass (the graph's static
        outputs) are kept with the graph's key and handed out on every replay -- `_sink_preds` itself may meanwhile
        hold the results of eager calls made in between."""
        if not static:
            return self._sink_preds
        st = net._graphs.get(getattr(net, "_last_graph_slot", "net"))
        if getattr(net, "n_graph_captures", 0) != captures_before:  # this call captured (again)
            self._graph_preds = (st["key"], dict(self._sink_preds))
            if not hasattr(self, "_graph_preds_by_slot"):
                self._graph_preds_by_slot = {}
            self._graph_preds_by_slot[getattr(net, "_last_graph_slot", "net")] = self._graph_preds
        rec = self._graph_preds_by_slot.get(getattr(net, "_last_graph_slot", "net")) if hasattr(self, "_graph_preds_by_slot") else self._graph_preds
        if st is None or rec is None or rec[0] != st["key"]:
            raise RuntimeError("SLIM: the replayed CUDA graph has no decoded outputs on record (it was not captured "
                               "through SLIM.forward)")
        return rec[1]

    _POINTWISE = ("static_flow", "dynamic_flow", "dynamicness", "staticness", "aggregated_flow", "static_aggr_flow")

    def _present(self, ret, n_points: int, thr, copy: bool = False):
        """The prediction with its point-wise tensors cut to the real number of points (graph mode pads to capacity);
        `copy`: on private copies of the packed buffers instead of views of the graph's static outputs."""
        if copy:
            bev, aggr, cls, pts, trafo, nep, sa = ret._packed
            ret = HeadDecoder._assemble(bev.clone(), aggr.clone() if aggr is not None else None, cls.clone(), pts.clone(),
                                        trafo.clone(), nep.clone(), ret.dynamicness_threshold, sa)
        out = AttrDict(ret)
        for name in self._POINTWISE:
            if name in ret and ret[name].shape[1] != n_points:
                out[name] = ret[name][:, :n_points]
        out.dynamicness_threshold = thr
        return out

    def forward(self, sample_data_t0, sample_data_t1, summaries=None):
        preds = self.forward_frames([sample_data_t0, sample_data_t1], [(0, 1)])
        self.predictions_fw, self.predictions_bw = preds[0], preds[1]
        return preds[0], preds[1]

    TRIPLE_PAIRS = ((0, 1), (0, 2), (1, 2))  # t0 -> t1, t0 -> t2, t1 -> t2 (experiment.py:386-456)

    def forward_triple(self, sample_data_t0, sample_data_t1, sample_data_t2):
        """The three model() calls the reference's KITTI / nuScenes export makes per sample (``experiment.py:386-456``) as ONE
        pass in which every frame is encoded once.  Returns {"t0_t1": preds, "t1_t0": ..., "t0_t2", "t2_t0", "t1_t2",
        "t2_t1"} -- each what ``forward`` returns for that direction (a list over decoded iterations)."""
        preds = self.forward_frames([sample_data_t0, sample_data_t1, sample_data_t2], self.TRIPLE_PAIRS)
        out = {}
        for p, (a, b) in enumerate(self.TRIPLE_PAIRS):
            out["t%d_t%d" % (a, b)] = preds[2 * p]
            out["t%d_t%d" % (b, a)] = preds[2 * p + 1]
        return out

    def forward_frames(self, samples, pairs):
        """``samples[f]``: the sample dictionary of frame f, ``pairs``: [(a, b), ...].  Returns the decoded predictions per
        direction: [a -> b of pair 0, b -> a of pair 0, a -> b of pair 1, ...]; a direction is decoded with the points,
        pillar coordinates and occupancy of its SOURCE frame, like ``slim.py:77-148``."""
        dev = next(self.parameters()).device
        raw = bool(samples[0].get("raw_scan", False)) and not self.cfg.data.use_ground_for_network
        assert all((bool(s.get("raw_scan", False)) and not self.cfg.data.use_ground_for_network) == raw for s in samples)
        net = self.raft_network
        pcls = [get_network_input_pcls(self.cfg, s, "ta", to_device=dev) for s in samples]
        thr = self.moving_dynamicness_threshold.value()
        sources = [samples[f] for a, b in pairs for f in (a, b)]  # per direction
        n_dirs = len(sources)
        if self.decode_as_sink and dev.type == "cuda":
            static = net.will_use_graph(*pcls)
            key = self._stage_decode_inputs(sources, dev, thr, static)
            net.output_sink, net.output_sink_begin = self._sink, self._sink_begin
            net.graph_extra_key = (key, self.static_aggregation) if static else None
            captures_before = getattr(net, "n_graph_captures", 0)
            net.forward_frames(pcls, pairs, raw_scans=raw)
            decoded = self._sink_results(net, static, captures_before)
            copy = static and not self.outputs_alias_static_buffers
            return [[self._present(decoded[(k, it)], self._dec_n[k], thr, copy)
                     for it in sorted(i for (kk, i) in decoded if kk == k)] for k in range(n_dirs)]
        net.output_sink = net.output_sink_begin = net.graph_extra_key = None
        outs, occs = net.forward_frames(pcls, pairs, raw_scans=raw)
        frame_of = [f for a, b in pairs for f in (a, b)]
        preds = [[] for _ in range(n_dirs)]
        moved: Dict[int, tuple] = {}
        for k in range(n_dirs):
            f = frame_of[k]
            if f not in moved:
                pt = samples[f]["pcl_ta"]
                moved[f] = (pt["pcl"].to(dev, non_blocking=True), pt["pillar_coors"].to(dev, non_blocking=True),
                            pt["pcl_is_valid"].to(dev, non_blocking=True), torch.squeeze(occs[f] > 0.5, dim=1))
            pc, coors, valid, filled = moved[f]
            dec = self.head_decoder_fw if k % 2 == 0 else self.head_decoder_bw
            for o in outs[k]:  # one entry per iteration, or only the last one in "last" mode
                preds[k].append(dec(o, thr, pc=pc, pointwise_voxel_coordinates=coors, pointwise_valid_mask=valid,
                                    filled_pillar_mask=filled, static_aggregation=self.static_aggregation))
        return preds
