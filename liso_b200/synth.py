"""Seeded synthetic spinning-LiDAR frame pairs with the shapes of the reference datasets.

There is no dataset access, so the bench and the parity tests use ray-cast scenes:
a ground plane plus seeded random boxes (cars, walls, poles), scanned by a spinning
sensor (K: 64 beams, N: 32 beams, A: 2x32 beams), frame t1 = same scene after ego motion
and object motion.  The per-sample dict layout mirrors what the reference's DataLoader
hands to ``SLIM.forward``:

* ``pcl_full_no_ground_ta``  list[B] of (N_i, 4) f32   -- network input (``kabsch/main_utils.py:247-261``)
* ``pcl_ta``: ``pcl`` (B, Nmax, 4) NaN-padded, ``pcl_is_valid`` (B, Nmax) bool,
  ``pillar_coors`` (B, Nmax, 2) int32 padded with -1      (collate: ``torch_dataset_commons.py:380-401``)
* ``gt.odom_ta_tb`` (B, 4, 4) f64

Ground removal follows ``infer_ground_label_using_cone`` (``torch_dataset_commons.py:133-146``,
threshold ``liso_config.yml:113-114``); ``pcl_ta`` / ``pillar_coors`` follow ``voxelize_sample``
(``torch_dataset_commons.py:975-987``) and are computed by :func:`pillar_coors_f64_numpy`, a
host-side helper kept bit-identical to ``voxelize_pcl`` (``datasets/nuscenes/analyse_boxes.py:6-26``).
"""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np

SENSOR_HEIGHT_M = 1.73
MAX_RANGE_M = 85.0

_BEAM_LAYOUT = {
    # (number of beams, elevation max deg, elevation min deg)
    64: (64, 2.0, -24.8),
    32: (32, 10.67, -30.67),
}


def _scene(rng: np.random.Generator, half_extent: float) -> Dict[str, np.ndarray]:
    """Random oriented boxes: centre (x,y,z), size (l,w,h), yaw, velocity (vx,vy)."""
    boxes = []

    def add(n, size_lo, size_hi, moving_frac):
        for _ in range(n):
            cx, cy = rng.uniform(-half_extent, half_extent, size=2)
            l, w, h = rng.uniform(size_lo, size_hi)
            yaw = rng.uniform(-np.pi, np.pi)
            # keep a 5 m disc around the sensor free: distance from the origin to the footprint
            c, s = np.cos(-yaw), np.sin(-yaw)
            ox, oy = -(c * cx - s * cy), -(s * cx + c * cy)
            dx, dy = max(abs(ox) - l / 2, 0.0), max(abs(oy) - w / 2, 0.0)
            if dx * dx + dy * dy < 25.0:
                continue
            speed = rng.uniform(2.0, 12.0) if rng.uniform() < moving_frac else 0.0
            boxes.append([cx, cy, -SENSOR_HEIGHT_M + h / 2, l, w, h, yaw, speed * np.cos(yaw), speed * np.sin(yaw)])

    add(90, (3.5, 1.6, 1.4), (5.5, 2.1, 2.0), 0.35)  # cars
    add(60, (8.0, 0.4, 2.5), (30.0, 1.0, 7.0), 0.0)  # walls / facades
    add(120, (0.2, 0.2, 2.0), (0.6, 0.6, 6.0), 0.0)  # poles / trunks
    add(80, (1.0, 1.0, 1.0), (3.5, 3.5, 4.0), 0.0)  # bushes / clutter
    # vegetation height field (1.5 m cells over +-120 m): ground hits inside a vegetated cell
    # return from a random height inside the canopy instead of from the ground
    veg = rng.uniform(0.3, 3.0, size=(160, 160)) * (rng.uniform(size=(160, 160)) < 0.45)
    return {"boxes": np.asarray(boxes, dtype=np.float64), "veg": veg}


def _ray_dirs(beams: int, n_azimuth: int, dual: bool) -> np.ndarray:
    nb, emax, emin = _BEAM_LAYOUT[beams]
    elev = np.deg2rad(np.linspace(emax, emin, nb))
    az = np.linspace(-np.pi, np.pi, n_azimuth, endpoint=False)
    if dual:  # AV2: two stacked 32-beam sensors, second one offset by half a step
        elev = np.concatenate([elev, elev + np.deg2rad(0.4)])
    ce, se = np.cos(elev), np.sin(elev)
    d = np.stack(
        [np.outer(np.cos(az), ce), np.outer(np.sin(az), ce), np.broadcast_to(se, (n_azimuth, elev.size))], axis=-1
    )
    return d.reshape(-1, 3)


def _cast(dirs: np.ndarray, n_rows: int, boxes: np.ndarray, veg: np.ndarray, veg_T: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Nearest hit of each ray (origin 0) with ground/vegetation and the boxes -> (N,4) f32.

    ``veg_T`` maps sensor-frame xy to the (static) world frame of the vegetation field.
    """
    n = dirs.shape[0]
    t_best = np.full(n, np.inf)
    dz = dirs[:, 2]
    down = dz < -1e-6
    t_ground = np.where(down, -SENSOR_HEIGHT_M / np.where(down, dz, -1.0), np.inf)
    gx = dirs[:, 0] * np.where(down, t_ground, 0.0)
    gy = dirs[:, 1] * np.where(down, t_ground, 0.0)
    wx = veg_T[0, 0] * gx + veg_T[0, 1] * gy + veg_T[0, 3]
    wy = veg_T[1, 0] * gx + veg_T[1, 1] * gy + veg_T[1, 3]
    ci = np.clip(((wx + 120.0) / 1.5).astype(np.int64), 0, veg.shape[0] - 1)
    cj = np.clip(((wy + 120.0) / 1.5).astype(np.int64), 0, veg.shape[1] - 1)
    hveg = veg[ci, cj] * rng.uniform(0.0, 1.0, size=n)
    # move the return back along the ray so that it sits hveg above the ground
    t_ground = np.where(down, (-SENSOR_HEIGHT_M + hveg) / np.where(down, dz, -1.0), np.inf)
    t_best = np.minimum(t_best, t_ground)
    # rays are laid out (azimuth-major, beam-minor); each box only sees a small azimuth interval
    n_rows = int(n_rows)
    n_az = n // n_rows
    for b in boxes:
        c, s = np.cos(b[6]), np.sin(b[6])
        hl, hw = b[3] / 2, b[4] / 2
        corners = np.array([[hl, hw], [hl, -hw], [-hl, -hw], [-hl, hw]])
        cxy = corners @ np.array([[c, s], [-s, c]]) + b[:2]
        ang = np.arctan2(cxy[:, 1], cxy[:, 0])
        a0 = np.arctan2(b[1], b[0])
        rel = np.mod(ang - a0 + np.pi, 2 * np.pi) - np.pi
        lo = a0 + rel.min() - 0.01
        hi = a0 + rel.max() + 0.01
        i0 = int(np.floor((lo + np.pi) / (2 * np.pi) * n_az))
        i1 = int(np.ceil((hi + np.pi) / (2 * np.pi) * n_az)) + 1
        az_idx = np.arange(i0, i1) % n_az
        ridx = (az_idx[:, None] * n_rows + np.arange(n_rows)[None, :]).reshape(-1)
        dsub = dirs[ridx]
        c, s = np.cos(-b[6]), np.sin(-b[6])
        # ray in box frame
        ox, oy, oz = -b[0], -b[1], -b[2]
        o = np.array([c * ox - s * oy, s * ox + c * oy, oz])
        d = np.stack([c * dsub[:, 0] - s * dsub[:, 1], s * dsub[:, 0] + c * dsub[:, 1], dsub[:, 2]], axis=-1)
        half = b[3:6] / 2
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            t1 = (-half - o) * inv
            t2 = (half - o) * inv
        tmin = np.nanmax(np.minimum(t1, t2), axis=-1)
        tmax = np.nanmin(np.maximum(t1, t2), axis=-1)
        hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.5)
        cur = t_best[ridx]
        t_best[ridx] = np.where(hit & (tmin < cur), tmin, cur)
    ok = np.isfinite(t_best) & (t_best < MAX_RANGE_M)
    t = t_best[ok] + rng.normal(0.0, 0.02, size=int(ok.sum()))
    pts = dirs[ok] * t[:, None]
    inten = rng.uniform(0.0, 1.0, size=(pts.shape[0], 1))
    return np.concatenate([pts, inten], axis=-1).astype(np.float32)


def ground_mask_cone(pcl: np.ndarray, cone_z_threshold_m: float = -1.5, cone_angle_deg: float = 0.8) -> np.ndarray:
    """``infer_ground_label_using_cone`` (``torch_dataset_commons.py:133-146``)."""
    cone_angle = cone_angle_deg / 180.0 * np.pi
    d_xy = np.linalg.norm(pcl[..., 0:2], axis=-1)
    return pcl[..., 2] < cone_z_threshold_m + np.tan(cone_angle) * d_xy


def pillar_coors_f64_numpy(pcl: np.ndarray, bev_range_m, grid_size, height_range_m=(-2.0, 1.0)):
    """Dataset-side point->pillar map, identical arithmetic to ``voxelize_pcl`` + ``voxelize_sample``.

    float32 points promoted against a float64 range array, ``astype(int32)`` truncation toward
    zero, strict ``zmin < z < zmax`` height filter.  Returns (coors (N,2) int32, valid (N,) bool).
    """
    rng_m = np.append(np.asarray(bev_range_m, dtype=np.float32), np.array(1000.0))  # float32 ranges widened (bev_utils.py:42)
    height_range_m = np.asarray(height_range_m, dtype=np.float32)
    gsz = np.append(np.asarray(grid_size, dtype=np.int64), 1)
    c = (pcl[:, :3] + 0.5 * rng_m) / rng_m
    c = (c * gsz).astype(np.int32)
    ok = (
        (0 <= c[:, 0]) & (0 <= c[:, 1]) & (0 <= c[:, 2]) & (c[:, 0] < gsz[0]) & (c[:, 1] < gsz[1]) & (c[:, 2] < gsz[2])
    )
    ok &= (height_range_m[0] < pcl[:, 2]) & (pcl[:, 2] < height_range_m[1])
    return c[:, :2], ok


def _ego_motion(rng: np.random.Generator) -> np.ndarray:
    yaw = np.deg2rad(rng.uniform(-2.0, 2.0))
    T = np.eye(4)
    T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    T[:2, 3] = [rng.uniform(0.6, 1.4), rng.uniform(-0.05, 0.05)]
    return T


def make_frame_pair(workload: Dict[str, Any], seed: int, n_frames: int = 2):
    """Return (full no-ground cloud t0, same t1, odom_t0_t1 4x4 f64) for one pair; with ``n_frames=3`` (t0, t1, t2,
    odom_t0_t1, odom_t0_t2) -- the first two frames are the pair's."""
    rng = np.random.default_rng(seed)
    bev = float(workload["bev_range_m"][0])
    beams = workload["beams"]
    dual = workload.get("dual", False)
    scene = _scene(rng, bev / 2 + 5.0)
    n_target = int(workload["n_points"])
    rows = _BEAM_LAYOUT[beams][0] * (2 if dual else 1)

    def frame(boxes, n_az, noise_rng, veg_T=np.eye(4)):
        pts = _cast(_ray_dirs(beams, n_az, dual), rows, boxes, scene["veg"], veg_T, noise_rng)
        return pts[~ground_mask_cone(pts)]

    # calibrate azimuth resolution so the ground-free cloud has about n_target points
    n_az = max(64, int(2.5 * n_target / rows))
    probe = frame(scene["boxes"], n_az // 8, np.random.default_rng(seed + 7))
    keep_per_az = max(probe.shape[0], 1) / (n_az // 8)
    n_az = max(64, int(round(n_target / keep_per_az)))
    pc0 = frame(scene["boxes"], n_az, rng)

    odom = _ego_motion(rng)  # pose of sensor at t1 expressed in t0
    boxes1 = scene["boxes"].copy()
    dt = 0.1
    boxes1[:, 0] += boxes1[:, 7] * dt
    boxes1[:, 1] += boxes1[:, 8] * dt
    inv = np.linalg.inv(odom)
    ctr = np.concatenate([boxes1[:, :3], np.ones((boxes1.shape[0], 1))], axis=-1) @ inv.T
    boxes1[:, :3] = ctr[:, :3]
    boxes1[:, 6] -= np.arctan2(odom[1, 0], odom[0, 0])
    pc1 = frame(boxes1, n_az, rng, odom)
    if n_frames == 2:
        return pc0, pc1, odom
    # a third frame (the KITTI / nuScenes export also runs t0 -> t2 and t1 -> t2): one more ego-motion / object step
    step = _ego_motion(rng)
    odom2 = odom @ step  # pose of the sensor at t2 expressed in t0
    boxes2 = scene["boxes"].copy()
    boxes2[:, 0] += boxes2[:, 7] * 2 * dt
    boxes2[:, 1] += boxes2[:, 8] * 2 * dt
    inv2 = np.linalg.inv(odom2)
    ctr2 = np.concatenate([boxes2[:, :3], np.ones((boxes2.shape[0], 1))], axis=-1) @ inv2.T
    boxes2[:, :3] = ctr2[:, :3]
    boxes2[:, 6] -= np.arctan2(odom2[1, 0], odom2[0, 0])
    pc2 = frame(boxes2, n_az, rng, odom2)
    return pc0, pc1, pc2, odom, odom2


def make_sample_dicts(workload: Dict[str, Any], seeds: List[int], as_torch: bool = True):
    """Batch of pairs in the reference's collated layout -> (sample_data_t0, sample_data_t1)."""
    import torch

    bev, grid = workload["bev_range_m"], workload["img_grid_size"]
    per_t: List[Dict[str, list]] = [dict(full=[], pcl=[], coors=[], odom=[]) for _ in range(2)]
    for s in seeds:
        pc0, pc1, odom = make_frame_pair(workload, s)
        for t, (pc, od) in enumerate(((pc0, odom), (pc1, np.linalg.inv(odom)))):
            coors, ok = pillar_coors_f64_numpy(pc, bev, grid)
            per_t[t]["full"].append(torch.from_numpy(pc))
            per_t[t]["pcl"].append(torch.from_numpy(pc[ok]))
            per_t[t]["coors"].append(torch.from_numpy(coors[ok]))
            per_t[t]["odom"].append(torch.from_numpy(od))
    out = []
    for t in range(2):
        pcl = torch.nn.utils.rnn.pad_sequence(per_t[t]["pcl"], batch_first=True, padding_value=float("nan"))
        coors = torch.nn.utils.rnn.pad_sequence(per_t[t]["coors"], batch_first=True, padding_value=-1)
        valid = torch.logical_not(torch.isnan(pcl).sum(-1))
        out.append(
            {
                "pcl_full_no_ground_ta": per_t[t]["full"],
                "pcl_ta": {"pcl": pcl, "pcl_is_valid": valid, "pillar_coors": coors},
                "gt": {"odom_ta_tb": torch.stack(per_t[t]["odom"], 0)},
            }
        )
    return out[0], out[1]


def _sample_of(pc: np.ndarray, bev, grid) -> Dict[str, Any]:
    """Un-batched sample dictionary of one frame (what ``liso_b200.slim.export.collate_pairs`` batches)."""
    import torch

    coors, ok = pillar_coors_f64_numpy(pc, bev, grid)
    return {"pcl_full_no_ground_ta": torch.from_numpy(pc),
            "pcl_ta": {"pcl": torch.from_numpy(pc[ok]), "pillar_coors": torch.from_numpy(coors[ok])}}


class ShiftedCloud:
    """A cloud that is a cast scene shifted as a whole, produced on demand: ``write_into`` adds the shift while copying into
    the caller's (pinned) buffer -- one pass over the points instead of transform + copy (``export.PinnedArena.pack``)."""

    def __init__(self, pc: np.ndarray, shift_xy):
        import torch

        self.pc = pc
        self.vec = np.array([shift_xy[0], shift_xy[1], 0.0, 0.0], dtype=np.float32)[:pc.shape[1]]
        self.shape, self.dtype = tuple(pc.shape), torch.float32

    def write_into(self, raw_u8: np.ndarray) -> None:
        np.add(self.pc, self.vec, out=raw_u8.view(np.float32).reshape(self.shape))

    def materialize(self):
        import torch

        return torch.from_numpy(self.pc + self.vec)


class SyntheticExportDataset:
    """Flow-export workload (BASELINE configs[4]): ``n_samples`` distinct samples of ``frames`` (2: t0, t1 | 3: t0, t1, t2)
    synthetic LiDAR frames each.  The ray caster makes a sample in ~2 s of CPU time, so the samples are drawn from a pool
    of ``pool`` cast scenes and made distinct by a per-sample yaw rotation + shift of the whole scene (every frame of a
    sample gets the same rigid motion: the relative motion between the frames -- what the network estimates -- is kept,
    every pillar map changes).  ``dataset[i] -> (sample_id, sample_t0, sample_t1[, sample_t2])`` with un-batched host
    tensors, deterministic in ``i``.

    ``raw=True``: the samples are RAW scans only (``{"pcl_full_w_ground_ta", "raw_scan": True}``): ground rule, height
    filter, pillar coordinates and compaction are left to the device (``liso_b200.datasets.preprocess_scans``, SURVEY
    8f.4), as ``ExportPipeline`` does for such samples -- the host then only moves bytes."""

    def __init__(self, workload: Dict[str, Any], n_samples: int, frames: int = 2, pool: int = 16, seed0: int = 5000,
                 raw: bool = False, motion: str = "rigid", lazy: bool = False):
        """``motion``: how a sample differs from the cast scene it is drawn from: "rigid" (yaw rotation + shift) or "shift"
        (shift only: one pass over the points, for hosts with few cores per GPU)."""
        assert frames in (2, 3) and motion in ("rigid", "shift")
        self.motion = motion
        self.lazy = bool(lazy)  # raw + shift-only samples as `ShiftedCloud`s (written straight into the upload buffer)
        self.workload, self.n, self.frames, self.pool, self.seed0 = workload, int(n_samples), frames, max(1, int(pool)), seed0
        self.raw = bool(raw)
        self._cache: Dict[int, tuple] = {}

    def prepare(self, workers: int = 1):
        """Cast the pool's scenes now (optionally on several threads) instead of lazily inside the export loop."""
        ks = [k for k in range(min(self.pool, self.n)) if k not in self._cache]
        if workers > 1 and len(ks) > 1:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(workers) as ex:
                list(ex.map(self._scene_frames, ks))
        else:
            for k in ks:
                self._scene_frames(k)
        return self

    def __len__(self) -> int:
        return self.n

    def _scene_frames(self, k: int):
        if k not in self._cache:
            out = make_frame_pair(self.workload, self.seed0 + k, n_frames=self.frames)
            self._cache[k] = tuple(out[:self.frames])
        return self._cache[k]

    def __getitem__(self, i: int):
        if not 0 <= i < self.n:
            raise IndexError(i)
        frames = self._scene_frames(i % self.pool)
        rng = np.random.default_rng(self.seed0 * 7919 + i)
        variant = i // self.pool
        yaw = 0.0 if variant == 0 or self.motion == "shift" else rng.uniform(-np.pi, np.pi)
        shift = np.zeros(2) if variant == 0 else rng.uniform(-1.5, 1.5, size=2)
        c, s_ = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
        bev, grid = self.workload["bev_range_m"], self.workload["img_grid_size"]
        samples = []
        for pc in frames:
            if yaw == 0.0 and self.raw and self.lazy:
                samples.append({"pcl_full_w_ground_ta": ShiftedCloud(pc, shift), "raw_scan": True})
                continue
            if yaw == 0.0:
                q = pc + np.array([shift[0], shift[1], 0.0, 0.0], dtype=np.float32)[:pc.shape[1]]
            else:
                q = pc.copy()
                q[:, 0] = c * pc[:, 0] - s_ * pc[:, 1] + np.float32(shift[0])
                q[:, 1] = s_ * pc[:, 0] + c * pc[:, 1] + np.float32(shift[1])
            if self.raw:
                import torch

                samples.append({"pcl_full_w_ground_ta": torch.from_numpy(q), "raw_scan": True})
                continue
            samples.append(_sample_of(q, bev, grid))
        return ("%06d" % i,) + tuple(samples)
