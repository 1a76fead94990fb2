"""CPU: the C-ABI library loads without a GPU and exports every symbol include/slimb200.h declares.
No compute call is made here."""
import ctypes
import os
import re

from liso_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "slimb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slimb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 10, names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(_lib.SYMBOLS) == names  # the Python binding covers exactly the header


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.slimb200_version() >= 103
    assert b"workspace" in lib.slimb200_strerror(-3)
    L = _lib.CorrLayout()
    assert lib.slimb200_corr_layout_init(2, 128, 80, 80, 4, ctypes.byref(L)) == 0
    assert (L.n_cols, L.n_panels, L.pitch, L.rows_padded) == (6400 + 1600 + 400 + 100, 67, 67 * 128, 6400)
    assert list(L.level_offset) == [0, 6400, 8000, 8400]
    assert lib.slimb200_corr_layout_init(1, 128, 115, 115, 4, ctypes.byref(L)) == 0
    assert list(L.level_h) == [115, 57, 28, 14] and L.n_cols == 13225 + 3249 + 784 + 196  # floor pooling
    assert L.rows_padded == 104 * 128  # 13225 source pixels in 128-row tiles
    assert lib.slimb200_corr_layout_init(1, 64, 80, 80, 4, ctypes.byref(L)) == -2  # D != 128 unsupported
    assert lib.slimb200_corr_pyramid_bytes(ctypes.byref(L), _lib.DTYPE_BF16) > 0
    p = _lib.PillarParams()
    p.grid[0], p.grid[1], p.grid[2] = 640, 640, 1
    p.max_points, p.max_voxels, p.c_in, p.c_out = 20, 40000, 4, 64
    assert lib.slimb200_pillar_workspace_bytes(8, 8 * 120000, ctypes.byref(p)) > 8 * 120000 * 24
    p.c_in = 5
    assert lib.slimb200_pillar_workspace_bytes(8, 1000, ctypes.byref(p)) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except RuntimeError as e:
        assert "no CPU" in str(e)
    else:
        raise AssertionError("load() must raise when the extension is missing")


def test_cpu_tensors_are_rejected():
    """The product path never falls back to CPU: a CPU tensor is an error, not a slow path."""
    import pytest
    import torch

    from liso_b200.config import make_cfg
    from liso_b200.networks.pcl_to_feature_grid import PointsPillarFeatureNetWrapper
    from liso_b200.slim.corr import CorrBlock

    m = PointsPillarFeatureNetWrapper(make_cfg("T")).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
        m([torch.zeros(10, 4)])
    with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
        CorrBlock(torch.zeros(1, 128, 8, 8), torch.zeros(1, 128, 8, 8))
    # ... and so is the output decoder (the stock-PyTorch restatement lives in tests/torch_decoder.py, not in the product)
    import numpy as np

    from liso_b200.slim.slim import HeadDecoder

    cfg = make_cfg("T")
    dec = HeadDecoder(cfg.SLIM, name="t", bev_extent=np.array([-7.0, -7.0, 7.0, 7.0]))
    with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
        dec(torch.zeros(1, 8, 8, 8), 0.5, pc=torch.zeros(1, 4, 4), pointwise_voxel_coordinates=torch.zeros(1, 4, 2, dtype=torch.int32),
            pointwise_valid_mask=torch.ones(1, 4, dtype=torch.bool), filled_pillar_mask=torch.ones(1, 8, 8, dtype=torch.bool))
    assert not hasattr(HeadDecoder, "_forward_torch")
    # switches that would change the decoder's arithmetic are refused, not ignored
    import copy

    for path in (("model", "use_static_aggr_flow_for_aggr_flow"), ("model", "dynamic_flow_is_non_rigid_flow"),
                 ("losses", "unsupervised", "use_epsilon_for_weighted_pc_alignment")):
        bad = copy.deepcopy(cfg.SLIM)
        node = bad
        for k in path[:-1]:
            node = node[k]
        node[path[-1]] = True
        with pytest.raises(NotImplementedError):
            HeadDecoder(bad, name="t", bev_extent=np.array([-7.0, -7.0, 7.0, 7.0]))


def test_glue_kernels_reject_cpu_tensors_and_bad_arguments():
    """SURVEY 8f.2 glue: same rule -- CPU tensors raise; the C entry points validate their arguments without a GPU."""
    import pytest
    import torch

    from liso_b200.slim import glue as G

    x = torch.zeros(1, 8, 4, 4).contiguous(memory_format=torch.channels_last)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.nhwc_cat([x, x])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.add_relu(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.gru_gate_zr(x, torch.zeros(8), x, x, 4)
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    one = (ctypes.c_void_p * 1)(ptr)
    # channel counts / offsets / pitches must be multiples of 4 (float4 rows); a slot must fit its destination
    assert lib.slimb200_nhwc_pack(one, (ctypes.c_int32 * 1)(6), None, 1, one, (ctypes.c_int32 * 1)(0), (ctypes.c_int32 * 1)(8), 1, 1, None) == -2
    assert lib.slimb200_nhwc_pack(one, (ctypes.c_int32 * 1)(8), None, 1, one, (ctypes.c_int32 * 1)(4), (ctypes.c_int32 * 1)(8), 1, 1, None) == -1
    assert lib.slimb200_nhwc_pack(one, (ctypes.c_int32 * 1)(8), None, 5, one, (ctypes.c_int32 * 1)(0), (ctypes.c_int32 * 1)(8), 1, 1, None) == -2
    assert lib.slimb200_gru_gate_zr(ptr, ptr, ptr, 8, ptr, ptr, 8, 6, 1, None) == -2
    assert lib.slimb200_gru_gate_out(None, ptr, ptr, ptr, 8, ptr, 8, 1, None) == -1
    assert lib.slimb200_add_relu(ptr, ptr, ptr, 6, None) == -2
    assert lib.slimb200_iter_update_taps(ptr, 2, ptr, ptr, 4, 1, 4, 4, ptr, ptr, ptr, None, 0, None) == -1  # even window
    assert lib.slimb200_add_relu(ptr, ptr, ptr, 0, None) == 0 and lib.slimb200_nhwc_pack(
        one, (ctypes.c_int32 * 1)(8), None, 1, one, (ctypes.c_int32 * 1)(0), (ctypes.c_int32 * 1)(8), 1, 0, None) == 0  # empty: no launch
