"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference (baurst/liso) on CPU.

Used only in the authoring container (where ``/root/reference`` exists) by
``oracle/gen_golden.py`` to produce the golden fixtures under ``tests/golden/`` and by
the ``not gpu`` tests that cross-check the oracle restatement against the live reference
when it is present.  Nothing in ``liso_b200/``, ``bench.py`` or the ``-m gpu`` tests
imports this module; the GPU box has no ``/root/reference``.

The reference's SLIM forward depends on packages that are not installed here
(``mmcv``, ``munch``, the ``mmdet3d`` registry machinery, ``liso.kabsch.main_utils`` ->
matplotlib/skimage/...).  They are replaced by ``sys.modules`` stubs that are
behaviour-neutral for the fp32 forward:

* ``mmcv.runner.force_fp32 / auto_fp16``  -> pass-through decorators (no ``fp16_enabled``
  on the path: ``pillar_encoder.py:63``, ``voxel_encoders/utils.py:138``)
* ``mmcv.cnn.build_norm_layer``           -> ``nn.BatchNorm1d(C, eps, momentum)``
  (cfg ``{"type": "BN1d", "eps": 1e-3, "momentum": 0.01}``, ``pcl_to_feature_grid.py:45``)
* ``mmcv.ops.Voxelization``               -> deterministic hard voxelisation that calls the
  reference's own numba kernel ``mmdet3d/core/voxel/voxel_generator.py:76-208`` (the
  in-tree twin of mmcv-full==1.7.1 ``hard_voxelize_forward``, ``docker/Dockerfile.base:68``)
* ``munch.Munch``                         -> attribute dict
* ``liso.kabsch.main_utils.get_network_input_pcls`` -> the 15-line function restated
  (``main_utils.py:247-261``)
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REF_ROOT = os.environ.get("LISO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "liso", "slim", "model"))


class _Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _passthrough_decorator(*dargs, **dkwargs):
    def deco(fn):
        return fn

    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]
    return deco


def _build_norm_layer(cfg, num_features, postfix=""):
    assert cfg["type"] in ("BN1d", "BN"), cfg
    layer = nn.BatchNorm1d(num_features, eps=cfg.get("eps", 1e-5), momentum=cfg.get("momentum", 0.1))
    return "bn" + str(postfix), layer


def _load_by_path(mod_name: str, path: str):
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    spec.loader.exec_module(mod)
    return mod


_INSTALLED = False


def install() -> None:
    """Install the stubs and make ``import liso...`` resolve to the reference tree."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    # --- mmcv -------------------------------------------------------------------
    mmcv = types.ModuleType("mmcv")
    mmcv_runner = types.ModuleType("mmcv.runner")
    mmcv_runner.force_fp32 = _passthrough_decorator
    mmcv_runner.auto_fp16 = _passthrough_decorator
    mmcv_cnn = types.ModuleType("mmcv.cnn")
    mmcv_cnn.build_norm_layer = _build_norm_layer
    mmcv_ops = types.ModuleType("mmcv.ops")

    vg = _load_by_path(
        "_ref_voxel_generator",
        os.path.join(REF_ROOT, "mmdetection3d/mmdet3d/core/voxel/voxel_generator.py"),
    )

    class Voxelization(nn.Module):
        """Deterministic stand-in with mmcv's call signature (ctor: mmcv/ops/voxelize.py).

        fp32 arrays in, so the numba kernel computes ``floor((p - min) / vs)`` in fp32
        exactly like mmcv's CUDA/CPU kernels do.
        """

        def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, deterministic=True):
            super().__init__()
            self.voxel_size = np.asarray(voxel_size, dtype=np.float32)
            self.point_cloud_range = np.asarray(point_cloud_range, dtype=np.float32)
            self.max_num_points = max_num_points
            self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)

        def forward(self, points):
            max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
            pts = points.detach().cpu().numpy().astype(np.float32)
            voxels, coors, num = vg.points_to_voxel(
                pts, self.voxel_size, self.point_cloud_range, self.max_num_points, True, max_voxels
            )
            dev = points.device
            return (
                torch.from_numpy(voxels).to(dev),
                torch.from_numpy(coors).to(dev),
                torch.from_numpy(num).to(dev),
            )

    class DynamicScatter(nn.Module):  # imported, never used on this path
        def __init__(self, *a, **k):
            super().__init__()

    mmcv_ops.Voxelization = Voxelization
    mmcv_ops.DynamicScatter = DynamicScatter
    mmcv.runner, mmcv.cnn, mmcv.ops = mmcv_runner, mmcv_cnn, mmcv_ops
    sys.modules.update({"mmcv": mmcv, "mmcv.runner": mmcv_runner, "mmcv.cnn": mmcv_cnn, "mmcv.ops": mmcv_ops})

    # --- munch ------------------------------------------------------------------
    munch = types.ModuleType("munch")
    munch.Munch = _Munch
    sys.modules["munch"] = munch

    # --- mmdet3d (only the three vendored files on the path) -----------------------
    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    for name in ("mmdet3d", "mmdet3d.models", "mmdet3d.models.voxel_encoders", "mmdet3d.models.middle_encoders"):
        m = types.ModuleType(name)
        m.__path__ = []  # mark as package
        sys.modules[name] = m
    builder = types.ModuleType("mmdet3d.models.builder")
    builder.VOXEL_ENCODERS = _Registry()
    builder.MIDDLE_ENCODERS = _Registry()
    sys.modules["mmdet3d.models.builder"] = builder
    sys.modules["mmdet3d.models"].builder = builder
    base = os.path.join(REF_ROOT, "mmdetection3d/mmdet3d/models")
    _load_by_path("mmdet3d.models.voxel_encoders.utils", os.path.join(base, "voxel_encoders/utils.py"))
    _load_by_path("mmdet3d.models.voxel_encoders.pillar_encoder", os.path.join(base, "voxel_encoders/pillar_encoder.py"))
    _load_by_path("mmdet3d.models.middle_encoders.pillar_scatter", os.path.join(base, "middle_encoders/pillar_scatter.py"))

    # --- liso package rooted at the reference tree ---------------------------------
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    kabsch = types.ModuleType("liso.kabsch")
    kabsch.__path__ = []
    main_utils = types.ModuleType("liso.kabsch.main_utils")

    def get_network_input_pcls(cfg, sample_data_t0, time_key, to_device=None):
        key = ("pcl_full_w_ground_%s" if cfg.data.use_ground_for_network else "pcl_full_no_ground_%s") % time_key
        if to_device:
            # the reference hard-codes "cuda" (slim.py:56,62); CPU runs keep tensors where they are
            dev = to_device if torch.cuda.is_available() else "cpu"
            return [el.to(dev) for el in sample_data_t0[key]]
        return sample_data_t0[key]

    main_utils.get_network_input_pcls = get_network_input_pcls
    import liso  # noqa: F401  (namespace from the reference tree)

    sys.modules["liso.kabsch"] = kabsch
    sys.modules["liso.kabsch.main_utils"] = main_utils
    _INSTALLED = True


def ref_modules():
    """Return the reference classes used as ground truth."""
    install()
    from liso.networks.pcl_to_feature_grid.pcl_to_feature_grid import PointsPillarFeatureNetWrapper
    from liso.slim.model.raft_code.corr import CorrBlock
    from liso.slim.model.raft_code.utils import bilinear_sampler, coords_grid, initialize_flow
    from liso.slim.model.raft_mod import RAFT
    from liso.slim.model.slim import SLIM
    from liso.datasets.nuscenes.analyse_boxes import voxelize_pcl

    return _Munch(
        PointsPillarFeatureNetWrapper=PointsPillarFeatureNetWrapper,
        CorrBlock=CorrBlock,
        bilinear_sampler=bilinear_sampler,
        coords_grid=coords_grid,
        initialize_flow=initialize_flow,
        RAFT=RAFT,
        SLIM=SLIM,
        voxelize_pcl=voxelize_pcl,
    )


def ref_source_functions(rel_path: str, names, namespace=None):
    """Execute the UNMODIFIED source text of the named functions / methods of a reference file whose module cannot be
    imported here (``liso/datasets/torch_dataset_commons.py`` pulls in matplotlib, pykitti, nuscenes, ...).  Methods come
    back as plain functions taking ``self`` first.  ``namespace``: the globals the code sees (default: numpy as ``np``,
    torch)."""
    import ast
    import textwrap

    import numpy as np

    src = open(os.path.join(REF_ROOT, rel_path)).read()
    ns = {"np": np, "torch": torch}
    ns.update(namespace or {})
    found = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            node.decorator_list = []
            code = textwrap.dedent("\n".join(src.split("\n")[node.lineno - 1:node.end_lineno]))
            exec(compile(code, os.path.join(REF_ROOT, rel_path), "exec"), ns)
            found[node.name] = ns[node.name]
    missing = [n for n in names if n not in found]
    if missing:
        raise KeyError("not found in %s: %s" % (rel_path, missing))
    return found


def ref_preprocess_functions(legacy_numpy_promotion: bool = False):
    """``infer_ground_label_using_cone`` (``torch_dataset_commons.py:133-146``) and ``LidarDataset.voxelize_sample``
    (``:975-987``, calling ``voxelize_pcl``, ``datasets/nuscenes/analyse_boxes.py:6-26``) of the reference, executed from
    their source.  ``legacy_numpy_promotion``: the reference environment (NumPy < 2, ``docker/Dockerfile.base``) casts the
    float64 scalar ``np.tan(angle)`` to the float32 of the array it multiplies (value-based casting); NumPy >= 2 promotes
    the product to float64.  With the flag the code sees an ``np`` whose ``tan`` already returns that float32 -- the only
    difference between the two promotion rules in these functions -- and is otherwise untouched."""
    install()
    import numpy as np
    from liso.datasets.nuscenes.analyse_boxes import voxelize_pcl

    np_seen = np
    if legacy_numpy_promotion:
        class _LegacyNp:
            def __getattr__(self, k):
                return getattr(np, k)

            @staticmethod
            def tan(x):
                return np.float32(np.tan(x))

        np_seen = _LegacyNp()
    fns = ref_source_functions("liso/datasets/torch_dataset_commons.py", ["infer_ground_label_using_cone", "voxelize_sample"],
                               {"np": np_seen, "voxelize_pcl": voxelize_pcl})
    return fns["infer_ground_label_using_cone"], fns["voxelize_sample"], voxelize_pcl
