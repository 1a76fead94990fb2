"""dataset-side pre-processing of raw scans on the GPU (reference: ``liso/datasets/torch_dataset_commons.py``).

``preprocess_scans`` turns a batch of RAW LiDAR scans that already live on the device into the sample dictionary
``SLIM.forward`` consumes -- what ``LidarDataset.pillarize_points_remove_ground_add_bev_ghm_occupancy``
(``torch_dataset_commons.py:1061-1106``) and the collate function (``:380-401``) build on the CPU:

* ``pcl_full_w_ground_ta``  the raw scans themselves; the pillar encoder drops the ground points in-kernel
  (``PointsPillarFeatureNetWrapper.forward(..., raw_scan=True)``), so no ``pcl_full_no_ground`` copy is made
* ``pcl_ta`` = ``{"pcl", "pillar_coors", "pcl_is_valid"}``: non-ground points inside the BEV / height range in scan
  order, their pillar coordinates (fp64 truncation arithmetic of ``voxelize_pcl``), padded with NaN / -1 / False
* ``counts``: kept points per sample, on the device (nothing is synchronised)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def preprocess_params(cfg, c_in: int = 4, cone_angle_deg: float = 0.8) -> _lib.PreprocessParams:
    p = _lib.PreprocessParams()
    ghm = cfg.data.get("ground_height_map", None) if hasattr(cfg.data, "get") else getattr(cfg.data, "ground_height_map", None)
    cone_z = float(ghm["ground_threshold"]) if ghm is not None and "ground_threshold" in ghm else -1.5  # liso_config.yml:113-114
    p.cone_z_threshold = float(np.float32(cone_z))
    p.cone_tan = float(np.float32(np.tan(cone_angle_deg / 180.0 * np.pi)))
    # the dataset keeps bev_range_m as a float32 array (utils/bev_utils.py:42) before it is widened to float64
    p.range_x, p.range_y = float(np.float32(cfg.data.bev_range_m[0])), float(np.float32(cfg.data.bev_range_m[1]))
    p.grid_x, p.grid_y = int(cfg.data.img_grid_size[0]), int(cfg.data.img_grid_size[1])
    hr = cfg.data.get("pillar_height_range_m", (-2.0, 1.0)) if hasattr(cfg.data, "get") else (-2.0, 1.0)
    p.z_min, p.z_max = float(np.float32(hr[0])), float(np.float32(hr[1]))  # float32 array (torch_dataset_commons.py:499-501)
    p.c_in = c_in
    return p


@_lib.on_device_of_args
def preprocess_scans(scans: Sequence[torch.Tensor], cfg, ground_labels: Optional[Sequence[Optional[torch.Tensor]]] = None,
                     cap: Optional[int] = None) -> Dict:
    """``scans``: list of ``(N_i, 3|4)`` float32 CUDA tensors (raw, with ground).  Returns the sample dictionary."""
    lib = _lib.load()
    B = len(scans)
    if B < 1 or B > _lib.MAX_BATCH:
        raise ValueError("batch size must be in [1, %d]" % _lib.MAX_BATCH)
    c_in = int(scans[0].shape[1])
    pts: List[torch.Tensor] = []
    for t in scans:
        _lib.require_cuda(t)
        if t.dim() != 2 or t.shape[1] != c_in:
            raise ValueError("expected (N, %d) scans, got %s" % (c_in, tuple(t.shape)))
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous() or t.data_ptr() % 16:
            t = t.float().contiguous().clone()
        pts.append(t)
    dev = pts[0].device
    n = [int(t.shape[0]) for t in pts]
    cap = max(n) if cap is None else int(cap)
    p = preprocess_params(cfg, c_in)
    pcl = torch.empty((B, cap, c_in), dtype=torch.float32, device=dev)
    coors = torch.empty((B, cap, 2), dtype=torch.int32, device=dev)
    valid = torch.empty((B, cap), dtype=torch.uint8, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    ws = torch.empty(max(256, lib.slimb200_preprocess_workspace_bytes(B, cap, C.byref(p))), dtype=torch.uint8, device=dev)
    ptrs = (C.c_void_p * B)(*[t.data_ptr() for t in pts])
    gl = None
    keep = []
    if ground_labels is not None:
        arr = []
        for g, t in zip(ground_labels, pts):
            if g is None:
                arr.append(None)
                continue
            g = g.to(device=dev, dtype=torch.uint8).contiguous()
            if g.shape[0] != t.shape[0]:
                raise ValueError("ground label length does not match its scan")
            keep.append(g)
            arr.append(g.data_ptr())
        gl = (C.c_void_p * B)(*arr)
    cnt = (C.c_int32 * B)(*n)
    _lib.check(lib.slimb200_preprocess_points(
        C.cast(ptrs, C.POINTER(C.c_void_p)), C.cast(gl, C.POINTER(C.c_void_p)) if gl is not None else None, cnt, B, cap,
        C.byref(p), pcl.data_ptr(), coors.data_ptr(), valid.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(),
        _lib.current_stream_ptr()))
    return {"pcl_full_w_ground_ta": pts, "raw_scan": True,
            "pcl_ta": {"pcl": pcl, "pillar_coors": coors, "pcl_is_valid": valid.view(torch.bool)}, "counts": counts}
