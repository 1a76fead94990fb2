"""B200 drop-in for ``liso/networks/pcl_to_feature_grid/pcl_to_feature_grid.py``.

``PointsPillarFeatureNetWrapper(cfg)`` keeps the reference's constructor, ``forward`` /
``extract_pts_feat`` / ``voxelize`` signatures, child-module names and state-dict keys
(``pts_voxel_encoder.pfn_layers.0.{linear.weight, norm.*}``), so a checkpoint written by the
reference loads with ``strict=True`` and ``RAFT`` / the detector nets can use it unchanged.
The whole body of ``extract_pts_feat`` (mmcv Voxelization -> PillarFeatureNet -> 2x
PointPillarsScatter, reference ``:86-102``) is one call into ``slimb200_pillar_encode``.

Forward only: the export path runs under ``torch.no_grad()`` (``experiment.py:323,363``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from .. import _lib


class HardVoxelization(nn.Module):
    """Parameter holder with the attributes of ``mmcv.ops.Voxelization`` (no weights)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels, deterministic=False):
        super().__init__()
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.max_num_points = int(max_num_points)
        self.max_voxels = tuple(max_voxels) if isinstance(max_voxels, (tuple, list)) else (max_voxels, max_voxels)
        self.deterministic = deterministic
        vs = np.asarray(self.voxel_size, dtype=np.float32)
        rg = np.asarray(self.point_cloud_range, dtype=np.float32)
        # fp32, like mmcv: grid_size = round((max - min) / voxel_size)
        self.grid_size = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)

    def extra_repr(self) -> str:
        return "voxel_size=%s, point_cloud_range=%s, max_num_points=%d, max_voxels=%s (B200 fused kernel)" % (
            self.voxel_size, self.point_cloud_range, self.max_num_points, self.max_voxels)


class PFNLayer(nn.Module):
    """Weights of ``mmdet3d`` ``PFNLayer`` (``voxel_encoders/utils.py:107-144``), last layer, max mode."""

    def __init__(self, in_channels: int, out_channels: int, eps: float = 1e-3, momentum: float = 0.01):
        super().__init__()
        self.units = out_channels
        self.norm = nn.BatchNorm1d(out_channels, eps=eps, momentum=momentum)
        self.linear = nn.Linear(in_channels, out_channels, bias=False)
        self.mode = "max"
        self.last_vfe = True


class PillarFeatureNet(nn.Module):
    """Weights + constants of ``PillarFeatureNet`` (``pillar_encoder.py:41-91``); compute is fused."""

    def __init__(self, in_channels, feat_channels, voxel_size, point_cloud_range, norm_eps=1e-3, norm_momentum=0.01):
        super().__init__()
        assert len(feat_channels) == 1, "SLIM uses a single PFN layer"
        self.legacy = True
        self.in_channels = in_channels + 6  # + cluster centre + voxel centre offsets
        self.pfn_layers = nn.ModuleList([PFNLayer(self.in_channels, feat_channels[0], norm_eps, norm_momentum)])
        self.vx, self.vy, self.vz = float(voxel_size[0]), float(voxel_size[1]), float(voxel_size[2])
        self.x_offset = self.vx / 2 + float(point_cloud_range[0])
        self.y_offset = self.vy / 2 + float(point_cloud_range[1])
        self.z_offset = self.vz / 2 + float(point_cloud_range[2])
        self.point_cloud_range = point_cloud_range


class PointPillarsScatter(nn.Module):
    """Shape holder of ``PointPillarsScatter`` (``pillar_scatter.py:20-26``); the scatter is fused."""

    def __init__(self, in_channels: int, output_shape: Sequence[int]):
        super().__init__()
        self.output_shape = output_shape
        self.ny = int(output_shape[0])
        self.nx = int(output_shape[1])
        self.in_channels = in_channels


class PointsPillarFeatureNetWrapper(nn.Module):
    """``canvas_memory_format``: ``"channels_last"`` (default) returns the ``(B, C, H, W)`` canvas in
    channels-last memory format -- same shape, same values, what the cuDNN convolutions that consume it
    want and what stores fastest -- ``"contiguous"`` returns the reference's NCHW-contiguous tensor.
    Can also be set with ``cfg.network.b200_canvas_memory_format``."""

    def __init__(self, cfg, canvas_memory_format=None) -> None:
        super().__init__()
        self.cfg = cfg
        if canvas_memory_format is None:
            canvas_memory_format = getattr(cfg.network, "b200_canvas_memory_format", None) \
                if not isinstance(cfg.network, dict) else cfg.network.get("b200_canvas_memory_format")
        canvas_memory_format = canvas_memory_format or "channels_last"
        if canvas_memory_format not in ("channels_last", "contiguous"):
            raise ValueError("canvas_memory_format must be 'channels_last' or 'contiguous'")
        self.canvas_memory_format = canvas_memory_format
        # ground rule applied inside the encoder when it is handed RAW scans (forward(..., raw_scan=True)):
        # cone_z_threshold__m = data.ground_height_map.ground_threshold, cone angle 0.8 deg
        # (torch_dataset_commons.py:133-146, 1171-1174)
        ghm = cfg.data.get("ground_height_map", None) if hasattr(cfg.data, "get") else getattr(cfg.data, "ground_height_map", None)
        self.ground_cone_z = float(ghm["ground_threshold"]) if ghm is not None and "ground_threshold" in ghm else -1.5
        self.ground_cone_deg = 0.8
        z_cut = cfg.data.setdefault("z_pillar_cutoff_value", 5.0)
        assert z_cut > 0.0, z_cut
        half = np.append(np.array(cfg.data.bev_range_m) / 2.0, z_cut)
        pc_range = np.concatenate([-half, half], axis=0)
        voxel_size = np.append(np.array(cfg.data.bev_range_m) / np.array(cfg.data.img_grid_size), 2 * z_cut)
        self.pts_voxel_layer = HardVoxelization(
            max_num_points=20, voxel_size=voxel_size, max_voxels=(40000, 40000), point_cloud_range=pc_range,
            deterministic=False,
        )
        if "use_lidar_intensity" in cfg.data:
            c_in = [3, 4][cfg.data.use_lidar_intensity]
        else:
            print("Warning - legacy mode: Using no lidar intensity!")
            c_in = 3
        crf = cfg.network.centerpoint.setdefault("channel_reduction_factor", 1)
        self.pts_voxel_encoder = PillarFeatureNet(
            in_channels=c_in, feat_channels=[64 // crf], voxel_size=voxel_size, point_cloud_range=pc_range,
            norm_eps=0.001, norm_momentum=0.01,
        )
        self.pts_middle_encoder = PointPillarsScatter(64 // crf, cfg.data.img_grid_size)
        self.debug_occupancy_pts_middle_encoder = PointPillarsScatter(1, cfg.data.img_grid_size)
        self._c_in = c_in
        self._workspace = None
        grid = self.pts_voxel_layer.grid_size
        assert int(grid[0]) == self.pts_middle_encoder.ny and int(grid[1]) == self.pts_middle_encoder.nx, (
            grid, cfg.data.img_grid_size)

    # ------------------------------------------------------------------ C-ABI plumbing
    def _params(self, training: bool, raw_scan: bool = False) -> _lib.PillarParams:
        vl, ve = self.pts_voxel_layer, self.pts_voxel_encoder
        p = _lib.PillarParams()
        rg = np.asarray(vl.point_cloud_range, dtype=np.float32)
        vs = np.asarray(vl.voxel_size, dtype=np.float32)
        for j in range(3):
            p.range_min[j] = float(rg[j])
            p.voxel_size[j] = float(vs[j])
            p.grid[j] = int(vl.grid_size[j])
        p.max_points = vl.max_num_points
        p.max_voxels = vl.max_voxels[0] if training else vl.max_voxels[1]
        p.vx, p.vy, p.vz = ve.vx, ve.vy, ve.vz
        p.x_offset, p.y_offset, p.z_offset = ve.x_offset, ve.y_offset, ve.z_offset
        p.c_in = self._c_in
        p.c_out = ve.pfn_layers[0].units
        p.bn_training = 1 if training else 0
        p.bn_eps = ve.pfn_layers[0].norm.eps
        p.bn_momentum = ve.pfn_layers[0].norm.momentum
        p.ground_filter = 1 if raw_scan else 0
        p.ground_cone_z = float(np.float32(self.ground_cone_z))
        p.ground_cone_tan = float(np.float32(np.tan(self.ground_cone_deg / 180.0 * np.pi)))
        nhwc = self.canvas_memory_format == "channels_last" and p.c_out in (4, 8, 16, 32, 64)
        p.canvas_layout = _lib.CANVAS_NHWC if nhwc else _lib.CANVAS_NCHW
        return p

    def _prep_points(self, pts: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        out = []
        for t in pts:
            _lib.require_cuda(t)
            if t.dim() != 2 or t.shape[1] != self._c_in:
                raise ValueError("expected (N, %d) points, got %s" % (self._c_in, tuple(t.shape)))
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous() or t.data_ptr() % 16:
                t = t.float().contiguous().clone()
            out.append(t)
        return out

    @_lib.on_device_of_args
    def _encode(self, pts: Sequence[torch.Tensor], want_voxels: bool, raw_scan: bool = False, out=None):
        assert isinstance(pts, (list, tuple)), type(pts)
        lib = _lib.load()
        pfn = self.pts_voxel_encoder.pfn_layers[0]
        pts = self._prep_points(pts)
        B = len(pts)
        if B < 1 or B > _lib.MAX_BATCH:
            raise ValueError("batch size must be in [1, %d]" % _lib.MAX_BATCH)
        dev = pts[0].device
        training = self.training
        p = self._params(training, raw_scan)
        ny, nx = self.pts_middle_encoder.ny, self.pts_middle_encoder.nx
        total = int(sum(t.shape[0] for t in pts))
        fmt = torch.channels_last if p.canvas_layout == _lib.CANVAS_NHWC else torch.contiguous_format
        if out is not None:  # caller-owned outputs (e.g. the static inputs of a CUDA graph)
            canvas, occupancy = out
            if (tuple(canvas.shape) != (B, p.c_out, ny, nx) or tuple(occupancy.shape) != (B, 1, ny, nx)
                    or canvas.dtype != torch.float32 or not canvas.is_contiguous(memory_format=fmt)
                    or not occupancy.is_contiguous() or canvas.device != dev):
                raise ValueError("out= buffers do not match the canvas this call produces")
        else:
            canvas, occupancy = self.empty_outputs(B, dev)
        ws_bytes = lib.slimb200_pillar_workspace_bytes(B, total, C.byref(p))
        if ws_bytes == 0:
            raise RuntimeError("slimb200_pillar_workspace_bytes rejected the configuration")
        if self._workspace is None or self._workspace.numel() < ws_bytes or self._workspace.device != dev:
            self._workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ptrs = (C.c_void_p * B)(*[t.data_ptr() for t in pts])
        counts = (C.c_int32 * B)(*[t.shape[0] for t in pts])
        extra = {}
        null = C.c_void_p(0)
        pc = co = npn = vo = p2p = null
        if want_voxels:
            cap = min(total, B * p.max_voxels)
            extra["pillar_counts"] = torch.zeros(B + 1, dtype=torch.int32, device=dev)
            extra["coors"] = torch.zeros((cap, 4), dtype=torch.int32, device=dev)
            extra["num_points"] = torch.zeros((cap,), dtype=torch.int32, device=dev)
            extra["voxels"] = torch.zeros((cap, p.max_points, p.c_in), dtype=torch.float32, device=dev)
            extra["pt2pillar"] = torch.empty((total,), dtype=torch.int32, device=dev)
            pc, co, npn, vo, p2p = (C.c_void_p(extra[k].data_ptr()) for k in
                                    ("pillar_counts", "coors", "num_points", "voxels", "pt2pillar"))
        norm = pfn.norm
        w = pfn.linear.weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        rc = lib.slimb200_pillar_encode(
            C.cast(ptrs, C.POINTER(C.c_void_p)), counts, B, C.byref(p),
            w.data_ptr(), norm.weight.data_ptr(), norm.bias.data_ptr(),
            norm.running_mean.data_ptr(), norm.running_var.data_ptr(),
            canvas.data_ptr(), occupancy.data_ptr(), pc, co, npn, vo, p2p,
            self._workspace.data_ptr(), self._workspace.numel(), _lib.current_stream_ptr(),
        )
        _lib.check(rc)
        if training and norm.num_batches_tracked is not None:
            norm.num_batches_tracked += 1
        return canvas, occupancy, extra

    def empty_outputs(self, batch: int, device):
        """Uninitialised (canvas, occupancy) with the shapes / memory format ``forward`` returns."""
        c_out = self.pts_voxel_encoder.pfn_layers[0].units
        nhwc = self.canvas_memory_format == "channels_last" and c_out in (4, 8, 16, 32, 64)
        ny, nx = self.pts_middle_encoder.ny, self.pts_middle_encoder.nx
        canvas = torch.empty((batch, c_out, ny, nx), dtype=torch.float32, device=device,
                             memory_format=torch.channels_last if nhwc else torch.contiguous_format)
        return canvas, torch.empty((batch, 1, ny, nx), dtype=torch.float32, device=device)

    # ------------------------------------------------------------------ reference interface
    @torch.no_grad()
    def voxelize(self, points):
        """Reference ``:56-84``: returns (voxels (P,20,C), num_points (P,), coors (P,4) = (b,z,xi,yi)).

        Needs the pillar count on the host to size the result, i.e. one device sync; the fused
        ``forward`` never does that.  Note: a full encode runs (outputs discarded)."""
        _, _, ex = self._encode(points, want_voxels=True)
        n = int(ex["pillar_counts"][-1].item())
        return ex["voxels"][:n], ex["num_points"][:n], ex["coors"][:n]

    @torch.no_grad()
    def voxelize_debug(self, points):
        """Everything the parity tests compare: dict(voxels, num_points, coors, pt2pillar, pillar_counts, canvas, occupancy)."""
        canvas, occ, ex = self._encode(points, want_voxels=True)
        n = int(ex["pillar_counts"][-1].item())
        return dict(voxels=ex["voxels"][:n], num_points=ex["num_points"][:n], coors=ex["coors"][:n],
                    pt2pillar=ex["pt2pillar"], pillar_counts=ex["pillar_counts"], canvas=canvas, occupancy=occ)

    def extract_pts_feat(self, pts, raw_scan: bool = False, out=None) -> Tuple[torch.Tensor, torch.Tensor]:
        if torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters()):
            raise RuntimeError(
                "liso_b200 pillar encoder is forward-only (flow export): call it under torch.no_grad() "
                "or freeze its parameters; autograd through the fused kernel is not provided")
        with torch.no_grad():
            canvas, occupancy, _ = self._encode(pts, want_voxels=False, raw_scan=raw_scan, out=out)
        return canvas, occupancy

    def forward(self, pcl_t0, img_t0=None, raw_scan: bool = False, out=None):
        """``raw_scan=True``: ``pcl_t0`` are raw scans (with ground); the ground points are dropped inside the kernel
        with the dataset's cone rule, which gives the same canvas as passing ``pcl_full_no_ground``."""
        return self.extract_pts_feat(pcl_t0, raw_scan=raw_scan, out=out)
