"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference SLIM hot path (baurst/liso).

This module is the *checker*.  It may be imported only from ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.  Nothing under ``liso_b200/`` imports it; the product path is CUDA-only and
fails loudly when its extension is missing.

Parity status ("pinned" = checked against something the reference itself provides):

* hard voxelisation (a2)            PINNED by the reference's golden vector
  ``mmdetection3d/tests/test_models/test_voxel_encoder/test_voxel_generator.py:8-24`` and by
  running the reference numba kernel (``voxel_generator.py:137-208``) -> ``tests/golden/voxelize_*.npz``
* initialize_flow convention (a11)  PINNED by the comment ``raft_mod.py:134-135``
* PFN / scatter / correlation / pyramid / lookup / dataset pillar coords / full forward
  (a3-a10, a12): no reference test pins values ("parity unpinned" by reference tests);
  PINNED instead by executing the unmodified reference modules in the authoring container
  (``oracle/ref_shims.py`` + ``oracle/gen_golden.py``) -> ``tests/golden/*.npz``.

The one arithmetic dependency that is not vendored in the reference tree is
``mmcv-full==1.7.1`` (``docker/Dockerfile.base:68``) ``mmcv.ops.Voxelization`` ->
``hard_voxelize_forward``.  Its published algorithm is restated here following the in-tree
twin ``mmdet3d/core/voxel/voxel_generator.py:76-208`` with the deterministic point-index
order (always a legal outcome of mmcv's ``deterministic=False`` kernel, identical whenever no
pillar exceeds 20 points and at most 40000 pillars are occupied).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MAX_POINTS_PER_PILLAR = 20  # pcl_to_feature_grid.py:25
MAX_PILLARS = 40000  # pcl_to_feature_grid.py:27
BN_EPS = 1e-3  # pcl_to_feature_grid.py:45
BN_MOMENTUM = 0.01  # pcl_to_feature_grid.py:45


# ----------------------------------------------------------------------------------
# a1: geometry of the pillar grid                      pcl_to_feature_grid.py:11-30
# ----------------------------------------------------------------------------------
def pillar_geometry(bev_range_m: Sequence[float], img_grid_size: Sequence[int], z_cutoff: float):
    """Returns (pc_range f64[6], voxel_size f64[3]) exactly as the wrapper ctor computes them."""
    pc_range_half = np.append(np.array(bev_range_m) / 2.0, z_cutoff)
    pc_range = np.concatenate([-pc_range_half, pc_range_half], axis=0)
    voxel_size = np.array(bev_range_m) / np.array(img_grid_size)
    voxel_size = np.append(voxel_size, 2 * z_cutoff)
    return pc_range, voxel_size


# ----------------------------------------------------------------------------------
# f.4: dataset-side pre-processing      torch_dataset_commons.py:133-146, 975-987, 1061-1184
# ----------------------------------------------------------------------------------
def ground_label_cone_f32(pcl: np.ndarray, cone_z_threshold_m: float = -1.5, cone_angle_deg: float = 0.8) -> np.ndarray:
    """``infer_ground_label_using_cone`` (``torch_dataset_commons.py:133-146``) as the reference environment
    evaluates it: float32 cloud, and -- NumPy < 2 value-based casting -- the float64 scalars ``tan(angle)`` and the
    threshold are cast to float32, so every operation rounds to float32 (written out explicitly here because
    NumPy >= 2 would promote to float64)."""
    x, y, z = (pcl[..., k].astype(np.float32) for k in range(3))
    d_xy = np.sqrt((x * x + y * y).astype(np.float32)).astype(np.float32)
    tan32 = np.float32(np.tan(cone_angle_deg / 180.0 * np.pi))
    thr = (np.float32(cone_z_threshold_m) + (tan32 * d_xy).astype(np.float32)).astype(np.float32)
    return z < thr


def preprocess_scan(pcl: np.ndarray, bev_range_m, img_grid_size, ground_label=None, cone_z_threshold_m: float = -1.5,
                    height_range_m=(-2.0, 1.0)):
    """One frame of ``pillarize_points_remove_ground_add_bev_ghm_occupancy`` (``torch_dataset_commons.py:1061-1106``):
    returns (pcl_full_no_ground, pcl_ta, pillar_coors) -- ground = label | cone rule (``:1164-1184``), pcl_ta = in-range
    (``voxelize_sample`` ``:975-987``) non-ground points in scan order."""
    ground = ground_label_cone_f32(pcl, cone_z_threshold_m)
    if ground_label is not None:
        ground = ground | ground_label.astype(bool)
    coors, in_range = pillar_coors_f64(pcl, bev_range_m, img_grid_size, height_range_m)
    keep = in_range & ~ground
    return pcl[~ground], pcl[keep], coors[keep]


# ----------------------------------------------------------------------------------
# a2: hard voxelisation                  mmdet3d/core/voxel/voxel_generator.py:76-208
# ----------------------------------------------------------------------------------
def hard_voxelize_loop(points, voxel_size, coors_range, max_points=MAX_POINTS_PER_PILLAR, max_voxels=MAX_PILLARS):
    """Line-by-line pure-Python restatement of ``_points_to_voxel_reverse_kernel``
    (``voxel_generator.py:137-208``), arithmetic in ``points.dtype``.  Small inputs only."""
    dt = points.dtype
    voxel_size = np.asarray(voxel_size, dtype=dt)
    coors_range = np.asarray(coors_range, dtype=dt)
    grid_size = np.round((coors_range[3:] - coors_range[:3]) / voxel_size).astype(np.int32)
    N = points.shape[0]
    coor_to_voxelidx: Dict[Tuple[int, int, int], int] = {}
    voxels = np.zeros((max_voxels, max_points, points.shape[1]), dtype=dt)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    pt2voxel = -np.ones((N,), dtype=np.int32)
    voxel_num = 0
    for i in range(N):
        coor = [0, 0, 0]
        failed = False
        for j in range(3):
            c = np.floor((points[i, j] - coors_range[j]) / voxel_size[j])
            if c < 0 or c >= grid_size[j]:
                failed = True
                break
            coor[2 - j] = int(c)
        if failed:
            continue
        key = (coor[0], coor[1], coor[2])
        voxelidx = coor_to_voxelidx.get(key, -1)
        if voxelidx == -1:
            voxelidx = voxel_num
            if voxel_num >= max_voxels:
                continue
            voxel_num += 1
            coor_to_voxelidx[key] = voxelidx
            coors[voxelidx] = coor
        n = num[voxelidx]
        if n < max_points:
            voxels[voxelidx, n] = points[i]
            num[voxelidx] += 1
            pt2voxel[i] = voxelidx
    return voxels[:voxel_num], coors[:voxel_num], num[:voxel_num], pt2voxel


def hard_voxelize(points, voxel_size, coors_range, max_points=MAX_POINTS_PER_PILLAR, max_voxels=MAX_PILLARS):
    """Vectorised numpy restatement of the same kernel (``voxel_generator.py:137-208``).

    Returns (voxels (P,max_points,C), coors (P,3) int32 in (z,y,x), num_points (P,) int32,
    pt2voxel (N,) int32 = pillar ordinal of every point that was stored, else -1).
    Arithmetic is carried out in ``points.dtype`` (fp32 for the SLIM path: mmcv casts
    ``voxel_size`` / ``point_cloud_range`` to float, the numba twin keeps fp32 arrays fp32).
    NaN coordinates are treated as out of range.
    """
    dt = points.dtype
    voxel_size = np.asarray(voxel_size, dtype=dt)
    coors_range = np.asarray(coors_range, dtype=dt)
    grid = np.round((coors_range[3:] - coors_range[:3]) / voxel_size).astype(np.int64)
    N, C = points.shape
    c = np.floor((points[:, :3] - coors_range[:3]) / voxel_size)  # dtype dt, true division
    with np.errstate(invalid="ignore"):
        ok = np.all((c >= 0) & (c < grid.astype(dt)), axis=1)
    idx_ok = np.nonzero(ok)[0]
    ci = c[idx_ok].astype(np.int64)
    cell = (ci[:, 2] * grid[1] + ci[:, 1]) * grid[0] + ci[:, 0]
    uniq, first_pos, inverse = np.unique(cell, return_index=True, return_inverse=True)
    order = np.argsort(first_pos, kind="stable")  # unique cells by first appearance
    ordinal_of_uniq = np.empty_like(order)
    ordinal_of_uniq[order] = np.arange(order.size)
    vox_of_pt = ordinal_of_uniq[inverse]  # per in-range point
    P = min(order.size, max_voxels)
    # rank of each in-range point inside its voxel, by point index
    by_vox = np.argsort(vox_of_pt, kind="stable")
    sorted_vox = vox_of_pt[by_vox]
    start = np.searchsorted(sorted_vox, np.arange(order.size), side="left")
    rank = np.empty_like(by_vox)
    rank[by_vox] = np.arange(by_vox.size) - start[sorted_vox]
    keep = (vox_of_pt < max_voxels) & (rank < max_points)
    voxels = np.zeros((P, max_points, C), dtype=dt)
    voxels[vox_of_pt[keep], rank[keep]] = points[idx_ok[keep]]
    num = np.bincount(vox_of_pt[keep], minlength=P).astype(np.int32)[:P]
    coors = np.zeros((P, 3), dtype=np.int32)
    first_ci = ci[first_pos[order[:P]]]
    coors[:, 0], coors[:, 1], coors[:, 2] = first_ci[:, 2], first_ci[:, 1], first_ci[:, 0]
    pt2voxel = -np.ones((N,), dtype=np.int32)
    pt2voxel[idx_ok[keep]] = vox_of_pt[keep]
    return voxels, coors, num, pt2voxel


# ----------------------------------------------------------------------------------
# a3: batch voxelisation + coordinate reorder        pcl_to_feature_grid.py:56-84
# ----------------------------------------------------------------------------------
def voxelize_batch(points: List[np.ndarray], pc_range, voxel_size):
    """-> voxels (sumP,20,C) f32, num_points (sumP,) i32, coors (sumP,4) i32 = (b, z, xi, yi),
    pt2pillar: list[B] of (N_i,) i32 (global pillar row, -1 = dropped)."""
    vs32 = np.asarray(voxel_size, dtype=np.float32)
    rg32 = np.asarray(pc_range, dtype=np.float32)
    voxels, nums, coors, p2p = [], [], [], []
    base = 0
    for b, pts in enumerate(points):
        v, c, n, p2v = hard_voxelize(np.ascontiguousarray(pts, dtype=np.float32), vs32, rg32)
        c = c[:, [0, 2, 1]]  # (z,y,x) -> (z,x,y)            pcl_to_feature_grid.py:73
        c = np.concatenate([np.full((c.shape[0], 1), b, dtype=np.int32), c], axis=1)  # :79-83
        voxels.append(v)
        nums.append(n)
        coors.append(c)
        p2p.append(np.where(p2v >= 0, p2v + base, -1).astype(np.int32))
        base += v.shape[0]
    return np.concatenate(voxels, 0), np.concatenate(nums, 0), np.concatenate(coors, 0), p2p


# ----------------------------------------------------------------------------------
# a4 + a5: PillarFeatureNet + PFNLayer    pillar_encoder.py:93-159, voxel_encoders/utils.py:146-182
# ----------------------------------------------------------------------------------
def pfn_forward(
    voxels: torch.Tensor,
    num_points: torch.Tensor,
    coors: torch.Tensor,
    linear_weight: torch.Tensor,
    bn_weight: torch.Tensor,
    bn_bias: torch.Tensor,
    running_mean: torch.Tensor,
    running_var: torch.Tensor,
    pc_range,
    voxel_size,
    training: bool,
    eps: float = BN_EPS,
    momentum: float = BN_MOMENTUM,
):
    """(P,20,C) -> (P,64).  Returns (features, new_running_mean, new_running_var).

    Restates, op for op: cluster offset over all 20 slots divided by num_points
    (``pillar_encoder.py:108-113``); legacy *in-place* voxel-centre offset that also overwrites the
    raw xyz channels and reads ``coors[:,3]`` for x / ``coors[:,2]`` for y (``:129-139``);
    concat to [xc,yc,zc,(i),dx,dy,dz,xc,yc,zc] (``:147``); padded rows zeroed (``:151-154``);
    Linear(no bias) -> BatchNorm1d over (P,64,20) -> ReLU -> max over the 20 slots, padded rows
    included (``voxel_encoders/utils.py:161-169``).
    """
    features = voxels.clone().float()
    vx, vy, vz = float(voxel_size[0]), float(voxel_size[1]), float(voxel_size[2])
    x_offset = vx / 2 + float(pc_range[0])
    y_offset = vy / 2 + float(pc_range[1])
    z_offset = vz / 2 + float(pc_range[2])
    points_mean = features[:, :, :3].sum(dim=1, keepdim=True) / num_points.type_as(features).view(-1, 1, 1)
    f_cluster = features[:, :, :3] - points_mean
    f_center = features[:, :, :3]  # a view: the in-place writes below alias `features`
    f_center[:, :, 0] = f_center[:, :, 0] - (coors[:, 3].type_as(features).unsqueeze(1) * vx + x_offset)
    f_center[:, :, 1] = f_center[:, :, 1] - (coors[:, 2].type_as(features).unsqueeze(1) * vy + y_offset)
    f_center[:, :, 2] = f_center[:, :, 2] - (coors[:, 1].type_as(features).unsqueeze(1) * vz + z_offset)
    feats = torch.cat([features, f_cluster, f_center], dim=-1)
    slots = feats.shape[1]
    mask = num_points.int().unsqueeze(1) > torch.arange(slots, dtype=torch.int).view(1, -1)
    feats = feats * mask.unsqueeze(-1).type_as(feats)
    x = F.linear(feats, linear_weight)
    rm, rv = running_mean.clone(), running_var.clone()
    x = F.batch_norm(x.permute(0, 2, 1).contiguous(), rm, rv, bn_weight, bn_bias, training, momentum, eps)
    x = F.relu(x.permute(0, 2, 1).contiguous())
    return torch.max(x, dim=1, keepdim=True)[0].squeeze(1), rm, rv


# ----------------------------------------------------------------------------------
# a6: PointPillarsScatter.forward_batch              pillar_scatter.py:62-102
# ----------------------------------------------------------------------------------
def pillar_scatter(voxel_features: torch.Tensor, coors: torch.Tensor, batch_size: int, ny: int, nx: int):
    C = voxel_features.shape[1]
    out = []
    for b in range(batch_size):
        canvas = torch.zeros(C, nx * ny, dtype=voxel_features.dtype)
        m = coors[:, 0] == b
        this = coors[m, :]
        indices = (this[:, 2] * nx + this[:, 3]).long()  # row = x index, col = y index
        canvas[:, indices] = voxel_features[m, :].t()
        out.append(canvas)
    return torch.stack(out, 0).view(batch_size, C, ny, nx)


def pillar_encoder_forward(points: List[np.ndarray], params: Dict[str, torch.Tensor], bev_range_m, img_grid_size,
                           z_cutoff: float = 10.0, training: bool = False):
    """``PointsPillarFeatureNetWrapper.forward`` (``pcl_to_feature_grid.py:86-107``) on CPU.

    ``params`` keys: linear_weight, bn_weight, bn_bias, running_mean, running_var.
    Returns dict(canvas, occupancy, coors, num_points, pt2pillar, pillar_features, running_mean, running_var).
    """
    pc_range, voxel_size = pillar_geometry(bev_range_m, img_grid_size, z_cutoff)
    voxels, nums, coors, p2p = voxelize_batch(points, pc_range, voxel_size)
    tv, tn, tc = torch.from_numpy(voxels), torch.from_numpy(nums), torch.from_numpy(coors)
    feats, rm, rv = pfn_forward(
        tv, tn, tc, params["linear_weight"], params["bn_weight"], params["bn_bias"],
        params["running_mean"], params["running_var"], pc_range, voxel_size, training,
    )
    B = len(points)
    ny, nx = int(img_grid_size[0]), int(img_grid_size[1])
    canvas = pillar_scatter(feats, tc, B, ny, nx)
    occ = pillar_scatter(torch.ones_like(feats[:, [0]]), tc, B, ny, nx)
    return dict(canvas=canvas, occupancy=occ, coors=tc, num_points=tn, pt2pillar=p2p, pillar_features=feats,
                running_mean=rm, running_var=rv)


# ----------------------------------------------------------------------------------
# a12: dataset-side point -> pillar map       datasets/nuscenes/analyse_boxes.py:6-26,
#                                             torch_dataset_commons.py:975-987
# ----------------------------------------------------------------------------------
def pillar_coors_f64(pcl: np.ndarray, bev_range_m, img_grid_size, height_range_m=(-2.0, 1.0)):
    # (the dataset keeps bev_range_m as a float32 array, utils/bev_utils.py:42; np.append with the float64 1000.0 widens it)
    bev = np.append(np.asarray(bev_range_m, dtype=np.float32), np.array(1000.0))
    height_range_m = np.asarray(height_range_m, dtype=np.float32)  # torch_dataset_commons.py:499-501
    grid = np.append(np.asarray(img_grid_size), np.array(1))
    c = (pcl[:, :3] + 0.5 * bev) / bev  # fp32 + fp64 -> fp64
    c = (c * grid).astype(np.int32)  # truncation toward zero
    ok = ((0 <= c[:, 0]) & (0 <= c[:, 1]) & (0 <= c[:, 2])
          & (c[:, 0] < grid[0]) & (c[:, 1] < grid[1]) & (c[:, 2] < grid[2]))
    ok = ok & (height_range_m[0] < pcl[:, 2]) & (pcl[:, 2] < height_range_m[1])
    return c[:, 0:2], ok


# ----------------------------------------------------------------------------------
# a7 + a8: all-pairs correlation and pyramid                          corr.py:7-21,48-56
# ----------------------------------------------------------------------------------
def corr_volume(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    b, d, h, w = fmap1.shape
    f1 = fmap1.reshape(b, d, h * w)
    f2 = fmap2.reshape(b, d, h * w)
    corr = torch.matmul(f1.transpose(1, 2), f2).view(b, h, w, 1, h, w)
    return corr / torch.sqrt(torch.tensor(d).float())


def corr_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4) -> List[torch.Tensor]:
    corr = corr_volume(fmap1, fmap2)
    b, h1, w1, dim, h2, w2 = corr.shape
    corr = corr.reshape(b * h1 * w1, dim, h2, w2)
    pyr = [corr]
    for _ in range(num_levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)  # floor mode: 115 -> 57 -> 28 -> 14
        pyr.append(corr)
    return pyr


# ----------------------------------------------------------------------------------
# a9 + a10: pyramid lookup                             corr.py:23-46, raft_code/utils.py:15-29
# ----------------------------------------------------------------------------------
def bilinear_sampler(img: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    H, W = img.shape[-2:]
    xgrid, ygrid = coords.split([1, 1], dim=-1)
    xgrid = 2 * xgrid / (W - 1) - 1
    ygrid = 2 * ygrid / (H - 1) - 1
    return F.grid_sample(img, torch.cat([xgrid, ygrid], dim=-1), align_corners=True)


def corr_lookup(pyramid: List[torch.Tensor], coords: torch.Tensor, radius: int = 3) -> torch.Tensor:
    r = radius
    coords = coords.permute(0, 2, 3, 1)
    batch, h1, w1, _ = coords.shape
    out = []
    for i, corr in enumerate(pyramid):
        d = torch.linspace(-r, r, 2 * r + 1)
        # first window axis offsets x, second offsets y (RAFT's transposition, corr.py:31-41)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1)
        centroid = coords.reshape(batch * h1 * w1, 1, 1, 2) / 2**i
        samp = bilinear_sampler(corr.float(), centroid + delta.view(1, 2 * r + 1, 2 * r + 1, 2))
        out.append(samp.view(batch, h1, w1, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def coords_grid(batch: int, ht: int, wd: int) -> torch.Tensor:
    """``raft_code/utils.py:32-37``: channel 0 = x (column index), channel 1 = y (row index)."""
    ys, xs = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)
