"""ctypes binding of ``libslimb200.so`` (the C ABI declared in ``include/slimb200.h``).

The library is the product path.  There is no fallback: if the shared object is missing, or a
call is made without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libslimb200.so")

MAX_BATCH = 64
MAX_LEVELS = 4
PANEL_COLS = 128
DTYPE_F32, DTYPE_BF16 = 0, 1


class PillarParams(C.Structure):
    _fields_ = [
        ("range_min", C.c_float * 3),
        ("voxel_size", C.c_float * 3),
        ("grid", C.c_int32 * 3),
        ("max_points", C.c_int32),
        ("max_voxels", C.c_int32),
        ("vx", C.c_float),
        ("vy", C.c_float),
        ("vz", C.c_float),
        ("x_offset", C.c_float),
        ("y_offset", C.c_float),
        ("z_offset", C.c_float),
        ("c_in", C.c_int32),
        ("c_out", C.c_int32),
        ("bn_training", C.c_int32),
        ("bn_eps", C.c_float),
        ("bn_momentum", C.c_float),
        ("ground_filter", C.c_int32),
        ("ground_cone_z", C.c_float),
        ("ground_cone_tan", C.c_float),
        ("canvas_layout", C.c_int32),
    ]


class CorrLayout(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("dim", C.c_int32),
        ("h", C.c_int32),
        ("w", C.c_int32),
        ("levels", C.c_int32),
        ("level_h", C.c_int32 * MAX_LEVELS),
        ("level_w", C.c_int32 * MAX_LEVELS),
        ("level_offset", C.c_int32 * MAX_LEVELS),
        ("n_cols", C.c_int32),
        ("pitch", C.c_int32),
        ("n_panels", C.c_int32),
        ("rows_padded", C.c_int32),
    ]


class PreprocessParams(C.Structure):
    _fields_ = [
        ("cone_z_threshold", C.c_float), ("cone_tan", C.c_float), ("range_x", C.c_double), ("range_y", C.c_double),
        ("grid_x", C.c_int32), ("grid_y", C.c_int32), ("z_min", C.c_float), ("z_max", C.c_float), ("c_in", C.c_int32),
        ("reserved", C.c_int32),
    ]


class DecodeParams(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_points", C.c_int32), ("pc_stride", C.c_int32),
        ("final_scale", C.c_int32), ("static_aggregation", C.c_int32), ("reserved", C.c_int32),
        ("ext_min_x", C.c_double), ("ext_min_y", C.c_double), ("ext_max_x", C.c_double), ("ext_max_y", C.c_double),
    ]


DECODE_BEV_CHANNELS, DECODE_POINT_CHANNELS = 16, 14


class DeflateMember(C.Structure):
    _fields_ = [("src", C.c_void_p), ("words_per_cell", C.c_int32), ("cell_stride", C.c_int32), ("n_words", C.c_uint32),
                ("first_chunk", C.c_uint32), ("crc_geo", C.c_uint32), ("reserved", C.c_uint32)]


DEFLATE_CHUNK_BYTES, DEFLATE_TABLE_BYTES = 8192, (256 + 2049 + 8192) * 4

# every symbol include/slimb200.h declares: (restype, argtypes)
SYMBOLS = {
    "slimb200_pillar_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64, C.POINTER(PillarParams)]),
    "slimb200_pillar_encode": (
        C.c_int,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32, C.POINTER(PillarParams)]
        + [C.c_void_p] * 12 + [C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "slimb200_pillar_coors_f64": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_float, C.c_float,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "slimb200_corr_layout_init": (C.c_int, [C.c_int32] * 5 + [C.POINTER(CorrLayout)]),
    "slimb200_corr_workspace_bytes": (C.c_size_t, [C.POINTER(CorrLayout)]),
    "slimb200_corr_pyramid_bytes": (C.c_size_t, [C.POINTER(CorrLayout), C.c_int32]),
    "slimb200_corr_build": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(CorrLayout), C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t,
         C.c_void_p],
    ),
    "slimb200_corr_lookup": (
        C.c_int,
        [C.c_void_p, C.c_int32, C.POINTER(CorrLayout), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p],
    ),
    "slimb200_corr_lookup_conv_packed_bytes": (C.c_size_t, [C.c_int32]),
    "slimb200_corr_lookup_conv_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "slimb200_corr_lookup_conv": (
        C.c_int,
        [C.c_void_p, C.c_int32, C.POINTER(CorrLayout), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
         C.c_void_p, C.c_int32, C.c_void_p],
    ),
    "slimb200_lookup_generation": (C.c_int, [C.c_int32]),
    "slimb200_lookup_conv_generation": (C.c_int, [C.c_int32]),
    "slimb200_preprocess_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.POINTER(PreprocessParams)]),
    "slimb200_preprocess_points": (
        C.c_int,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(PreprocessParams)]
        + [C.c_void_p] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "slimb200_head_decode_workspace_bytes": (C.c_size_t, [C.POINTER(DecodeParams)]),
    "slimb200_head_decode": (
        C.c_int,
        [C.c_void_p] * 7 + [C.POINTER(DecodeParams)] + [C.c_void_p] * 6 + [C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "slimb200_raft_output": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
         C.c_void_p],
    ),
    "slimb200_instnorm_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "slimb200_instnorm_nhwc": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "slimb200_instnorm_nhwc_slice": (
        C.c_int,
        [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "slimb200_bias_relu_slice": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "slimb200_nhwc_pack": (
        C.c_int,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p),
         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int64, C.c_void_p],
    ),
    "slimb200_gru_gate_zr": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p],
    ),
    "slimb200_gru_gate_out": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p],
    ),
    "slimb200_iter_update": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p],
    ),
    "slimb200_iter_update_taps": (
        C.c_int,
        [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p],
    ),
    "slimb200_add_relu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "slimb200_ctx_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "slimb200_add_bias_relu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "slimb200_deflate_plan": (C.c_int, [C.POINTER(DeflateMember), C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "slimb200_deflate_init": (C.c_int, [C.c_void_p, C.c_void_p]),
    "slimb200_deflate_encode": (
        C.c_int,
        [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p],
    ),
    "slimb200_strerror": (C.c_char_p, [C.c_int]),
    "slimb200_version": (C.c_int, []),
    "slimb200_profile_begin": (C.c_int, []),
    "slimb200_profile_end": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "slimb200_launch_count": (C.c_int64, [C.c_int32]),
    "slimb200_kernel_name": (C.c_char_p, [C.c_int32]),
}
N_KERNELS = 40
K_POINT_KEYS, K_SCAN_LOCAL, K_SCAN_GLOBAL, K_RANK_SCATTER = 0, 1, 2, 3
K_TILE_ENCODE, K_PILLAR_NHWC, K_FEAT_TRANSPOSE, K_FEAT_PACK, K_CORR_GEMM, K_CORR_LOOKUP = 6, 7, 8, 9, 10, 11
K_DECODE_BEV, K_DECODE_POINTS, K_DECODE_AGGR, K_RAFT_OUTPUT = 14, 15, 17, 18
K_IN_STATS, K_IN_FINALIZE, K_IN_APPLY = 24, 25, 26
K_LOOKUP_CONV = 32
K_DEFLATE_CHUNKS, K_DEFLATE_SCAN, K_DEFLATE_GATHER = 35, 36, 37
CANVAS_NCHW, CANVAS_NHWC = 0, 1

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libslimb200.so is missing (%s). Build it with `python -m liso_b200.build`; "
            "the SLIM hot path has no CPU or PyTorch fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    # tuning hooks (tools/, A/B runs of bench.py): SLIMB200_LOOKUP_GEN / SLIMB200_LOOKUP_CONV_GEN pick a kernel generation
    for env, fn in (("SLIMB200_LOOKUP_GEN", lib.slimb200_lookup_generation), ("SLIMB200_LOOKUP_CONV_GEN", lib.slimb200_lookup_conv_generation)):
        if os.environ.get(env, "").isdigit():
            fn(int(os.environ[env]))
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().slimb200_strerror(rc)
        raise RuntimeError("%s (code %d)" % (msg.decode() if msg else "slimb200 error", rc))


def current_stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_cuda_device(args):
    import torch

    for a in args:
        if torch.is_tensor(a):
            if a.is_cuda:
                return a.device
        elif isinstance(a, (list, tuple)):
            for b in a:
                if torch.is_tensor(b) and b.is_cuda:
                    return b.device
    return None


def on_device_of_args(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current: the C entry points launch on the CURRENT
    device's stream and keep per-device launch state, so a module living on cuda:1 must not be driven while cuda:0 is
    current (one process may drive several GPUs; bench.py uses one process per GPU + set_device)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        import torch

        dev = _first_cuda_device(args) or _first_cuda_device(tuple(kwargs.values()))
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "slimb200 kernels run on CUDA (sm_100a) only; got a %s tensor. There is no CPU fallback." % t.device
            )


# kernels of this library that ran as part of CUDA-graph replays (the library's own counter only sees the capture)
_graph_replayed_launches = 0


def note_graph_replay(n_kernels: int) -> None:
    global _graph_replayed_launches
    _graph_replayed_launches += int(n_kernels)


def total_launch_count() -> int:
    """Launches of slimb200 kernels so far: direct launches + kernels inside replayed CUDA graphs."""
    return int(load().slimb200_launch_count(-1)) + _graph_replayed_launches
