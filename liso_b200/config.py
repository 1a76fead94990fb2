"""Duck-typed configuration objects for the SLIM hot path.

The reference builds an OmegaConf ``DictConfig`` from ``liso/config/liso_config.yml``
(``config_helper/config.py:37-92``).  The hot-path constructors only use attribute
access, ``in`` and ``.setdefault`` on it (``pcl_to_feature_grid.py:14,31,37``), so a
small attribute dict is a faithful stand-in.  The values below restate the reference
defaults that the SLIM forward reads:

* ``SLIM.model.*``                 ``liso_config.yml:288-352`` overlaid by ``slim_RAFT`` (``:797-816``)
* ``data.*``                       ``liso_config.yml:47-49,115-119``
* grids ``slim_resolution``        ``liso_config.yml:524-531`` (640x640 over 70 m)
* grids ``slim_highest_resolution````liso_config.yml:542-549`` (920x920 over 120 m)
"""
from __future__ import annotations

import copy
from typing import Any, Dict


class AttrDict(dict):
    """dict with attribute access (what ``munch.Munch`` / OmegaConf give the reference)."""

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover - mirrors dict semantics
            raise AttributeError(k) from e

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = v

    def __delattr__(self, k: str) -> None:
        del self[k]

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_attr(d: Any) -> Any:
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return type(d)(to_attr(v) for v in d)
    return d


# name -> (bev_range_m, img_grid_size, approx points per frame, beams)
WORKLOADS: Dict[str, Dict[str, Any]] = {
    # KITTI-like, ``slim_resolution``
    "K": dict(bev_range_m=(70.0, 70.0), img_grid_size=(640, 640), n_points=120_000, beams=64),
    # nuScenes-like, same grid, sparse 32-beam
    "N": dict(bev_range_m=(70.0, 70.0), img_grid_size=(640, 640), n_points=35_000, beams=32),
    # AV2-like, ``slim_highest_resolution``
    "A": dict(bev_range_m=(120.0, 120.0), img_grid_size=(920, 920), n_points=100_000, beams=32, dual=True),
    # tiny grid for unit tests / smoke (not a reference config)
    "T": dict(bev_range_m=(28.0, 28.0), img_grid_size=(256, 256), n_points=20_000, beams=32),
}


def make_cfg(workload: str = "K", **overrides: Any) -> AttrDict:
    """Build the attr-dict the hot-path modules are constructed from."""
    w = WORKLOADS[workload]
    cfg = to_attr(
        {
            "data": {
                "bev_range_m": tuple(w["bev_range_m"]),
                "img_grid_size": tuple(w["img_grid_size"]),
                "use_lidar_intensity": True,
                "z_pillar_cutoff_value": 10.0,
                "pillar_height_range_m": (-2.0, 1.0),
                "use_ground_for_network": False,
            },
            "network": {"name": "slim", "centerpoint": {}},
            "SLIM": {
                "model": {
                    "name": "raft",
                    "dropout_rate": 0,
                    "raft_fnet_norm": "instance_affine",
                    "feature_downsampling_factor": 8,
                    "learn_upsampling": False,
                    "num_iters": 6,
                    "num_pred_iters": 6,
                    "flow_maps_archi": "single",
                    "corr_cfg": {
                        "module": "all",
                        "sampler": "bilinear",
                        "search_radius": 3,
                        "num_levels": 4,
                    },
                    "output_modification": {
                        "disappearing_logit": False,
                        "static_logit": "net",
                        "dynamic_logit": "net",
                        "ground_logit": False,
                        "dynamic_flow": "net",
                        "static_flow": "net",
                        "dynamic_flow_grad_scale": 1.0,
                    },
                    "predict_weight_for_static_aggregation": False,
                    "use_static_aggr_flow_for_aggr_flow": False,
                    "dynamic_flow_is_non_rigid_flow": False,
                    "point_pillars": {"nbr_point_feats": 64},
                    "u_net": {"final_scale": 1},
                },
                "losses": {"unsupervised": {"use_epsilon_for_weighted_pc_alignment": False}},
                "phases": {"train": {"dataset": "train", "mode": "unsupervised"}},
            },
        }
    )
    for k, v in overrides.items():
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = to_attr(v)
    return cfg
