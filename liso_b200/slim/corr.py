"""B200 drop-in for ``liso/slim/model/raft_code/corr.py`` and ``raft_code/utils.py``.

``CorrBlock(fmap1, fmap2, num_levels=4, radius=4)`` is constructed once per direction per
forward (``raft_mod.py:160-165``) and called once per GRU iteration (``:196``), exactly like the
reference class.  Construction = one operand-pack launch + one persistent tcgen05 GEMM that
writes the whole 4-level pyramid (bf16) in a single pass; ``__call__`` = one gather launch that
emits the ``(B, L*(2r+1)^2, h, w)`` fp32 tensor directly.

Differences a caller can observe (documented in DESIGN.md):
* ``corr_pyramid[l]`` has the reference's shape ``(B*h*w, 1, h_l, w_l)`` but is bf16 and is gathered
  lazily (a copy) out of the tiled buffer the kernels use (``include/slimb200.h``: 16 KB half-tiles, 4 source pixels
  x 8 columns per 64-byte unit).
* values carry bf16 operand + storage rounding: |err| <= 2^-7 * ||f1_i|| * ||f2_j|| / sqrt(D).
* forward only.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List

import torch
import torch.nn.functional as F

from .. import _lib


def make_layout(batch: int, dim: int, h: int, w: int, levels: int) -> _lib.CorrLayout:
    L = _lib.CorrLayout()
    _lib.check(_lib.load().slimb200_corr_layout_init(batch, dim, h, w, levels, C.byref(L)))
    return L


def _tile_view(pyramid: torch.Tensor, L: _lib.CorrLayout) -> torch.Tensor:
    """The packed pyramid as (b, panel, m, half, g, cb, p, e): half-tiles of 128 rows x 64 columns in which 4 neighbouring
    source pixels (p) x 8 consecutive columns (e) form one unit (``include/slimb200.h``)."""
    return pyramid.view(L.batch, L.n_panels, L.rows_padded // 128, 2, 32, 8, 4, 8)


def unpack_rows(pyramid: torch.Tensor, L: _lib.CorrLayout) -> torch.Tensor:
    """(B, h*w, n_panels*128) row-major copy of the packed pyramid: row = source pixel, column = pyramid column."""
    t = _tile_view(pyramid, L).permute(0, 2, 4, 6, 1, 3, 5, 7)  # (b, m, g, p, panel, half, cb, e)
    return t.reshape(L.batch, L.rows_padded, L.n_panels * _lib.PANEL_COLS)[:, :L.h * L.w]


def unpack_level(pyramid: torch.Tensor, L: _lib.CorrLayout, level: int) -> torch.Tensor:
    """Level ``level`` with the reference's shape (B*h*w, 1, h_l, w_l), gathered out of the panel layout."""
    nf = L.h * L.w
    off, hl, wl = L.level_offset[level], L.level_h[level], L.level_w[level]
    return unpack_rows(pyramid, L)[:, :, off:off + hl * wl].reshape(L.batch * nf, 1, hl, wl)


class _LazyLevels:
    """``corr_pyramid`` of the reference is a list of dense tensors; here the levels live interleaved in one
    panel-tiled buffer, so each level is materialised (a copy) only when somebody indexes the list."""

    def __init__(self, pyramid: torch.Tensor, L: _lib.CorrLayout):
        self._pyramid, self._L = pyramid, L

    def __len__(self) -> int:
        return self._L.levels

    def __getitem__(self, level: int) -> torch.Tensor:
        if isinstance(level, slice):
            return [self[i] for i in range(*level.indices(len(self)))]
        if level < 0:
            level += len(self)
        if not 0 <= level < len(self):
            raise IndexError(level)
        return unpack_level(self._pyramid, self._L, level)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


@_lib.on_device_of_args
def lookup(pyramid: torch.Tensor, L: _lib.CorrLayout, coords: torch.Tensor, radius: int,
           channels_last: bool = False) -> torch.Tensor:
    """``slimb200_corr_lookup`` on a packed pyramid (bf16 or fp32).  ``channels_last`` (radius 3 only) returns the
    same ``(B, L*(2r+1)^2, h, w)`` tensor in channels-last memory format."""
    _lib.require_cuda(pyramid, coords)
    if coords.shape != (L.batch, 2, L.h, L.w):
        raise ValueError("coords must be (B,2,h,w) = %s, got %s" % ((L.batch, 2, L.h, L.w), tuple(coords.shape)))
    coords = coords.detach()
    if coords.dtype != torch.float32 or not coords.is_contiguous():
        coords = coords.float().contiguous()
    n_ch = L.levels * (2 * radius + 1) ** 2
    channels_last = bool(channels_last) and radius == 3
    out = torch.empty((L.batch, n_ch, L.h, L.w), dtype=torch.float32, device=coords.device,
                      memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    dt = _lib.DTYPE_BF16 if pyramid.dtype == torch.bfloat16 else _lib.DTYPE_F32
    if pyramid.dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("pyramid dtype must be bfloat16 or float32")
    _lib.check(_lib.load().slimb200_corr_lookup(pyramid.data_ptr(), dt, C.byref(L), coords.data_ptr(), radius,
                                                out.data_ptr(), _lib.CANVAS_NHWC if channels_last else _lib.CANVAS_NCHW,
                                                _lib.current_stream_ptr()))
    return out


class PackedLookupConv:
    """The weights of the 1x1 convolution behind the lookup, packed for ``slimb200_corr_lookup_conv`` (tf32 B operand in
    its shared-memory layout + biases; ``slimb200_corr_lookup_conv_pack``).  Built once per weight tensor; the owner
    re-packs when the source parameters change (:meth:`matches`)."""

    @_lib.on_device_of_args
    def __init__(self, weight: torch.Tensor, bias: torch.Tensor = None, levels: int = 4, radius: int = 3):
        _lib.require_cuda(weight, bias)
        lib = _lib.load()
        self.c_out = int(weight.shape[0])
        n_ch = levels * (2 * radius + 1) ** 2
        w2 = weight.detach().reshape(self.c_out, -1)
        if w2.shape[1] != n_ch or w2.dtype != torch.float32:
            raise ValueError("weight must be fp32 (C_out, %d[, 1, 1])" % n_ch)
        w2 = w2.contiguous()
        b1 = None
        if bias is not None:
            if bias.numel() != self.c_out or bias.dtype != torch.float32:
                raise ValueError("bias must be fp32 (C_out)")
            b1 = bias.detach().contiguous()
        n_bytes = int(lib.slimb200_corr_lookup_conv_packed_bytes(self.c_out))
        if n_bytes == 0:
            raise RuntimeError("slimb200_corr_lookup_conv: C_out must be 32, 64 or 96 (got %d)" % self.c_out)
        self.levels, self.radius = levels, radius
        self.packed = torch.empty(n_bytes, dtype=torch.uint8, device=weight.device)
        _lib.check(lib.slimb200_corr_lookup_conv_pack(w2.data_ptr(), b1.data_ptr() if b1 is not None else None, levels, radius,
                                                      self.c_out, self.packed.data_ptr(), _lib.current_stream_ptr()))
        self._src = (weakref.ref(weight), weight._version, None if bias is None else weakref.ref(bias),
                     None if bias is None else bias._version)

    def matches(self, weight: torch.Tensor, bias: torch.Tensor = None) -> bool:
        """Packed from exactly these tensors in their current version (identity, not address)."""
        w_ref, w_ver, b_ref, b_ver = self._src
        if w_ref() is not weight or weight._version != w_ver or weight.device != self.packed.device:
            return False
        if bias is None or b_ref is None:
            return bias is None and b_ref is None
        return b_ref() is bias and bias._version == b_ver


@_lib.on_device_of_args
def lookup_conv(pyramid: torch.Tensor, L: _lib.CorrLayout, coords: torch.Tensor, radius: int, packed: PackedLookupConv,
                relu: bool = True, out: torch.Tensor = None) -> torch.Tensor:
    """``slimb200_corr_lookup_conv``: ``act(conv1x1(lookup(coords)))`` in one kernel -- the lookup of
    ``corr.py:23-46`` fused with ``SmallMotionEncoder.conv_stat_corr1`` + ReLU (``update.py:49,71``); the
    ``(B, L*49, h, w)`` lookup tensor never reaches HBM.  tf32 operands, fp32 accumulation on the tensor cores.

    Returns ``(B, C_out, h, w)`` fp32 in channels-last memory format.  ``out``: optional destination -- a channels-last
    tensor ``(B, >= C_out, h, w)`` whose first ``C_out`` channels are written (its other channels are left untouched)."""
    _lib.require_cuda(pyramid, coords, out)
    if pyramid.dtype != torch.bfloat16:
        raise ValueError("lookup_conv needs the bf16 pyramid")
    if coords.shape != (L.batch, 2, L.h, L.w):
        raise ValueError("coords must be (B,2,h,w) = %s, got %s" % ((L.batch, 2, L.h, L.w), tuple(coords.shape)))
    if packed.levels != L.levels or packed.radius != radius or packed.packed.device != pyramid.device:
        raise ValueError("weights were packed for another lookup geometry / device")
    coords = coords.detach()
    if coords.dtype != torch.float32 or not coords.is_contiguous():
        coords = coords.float().contiguous()
    c_out = packed.c_out
    if out is None:
        out = torch.empty((L.batch, c_out, L.h, L.w), dtype=torch.float32, device=coords.device, memory_format=torch.channels_last)
    elif not (out.dtype == torch.float32 and out.dim() == 4 and out.shape[0] == L.batch and out.shape[2:] == (L.h, L.w)
              and out.shape[1] >= c_out and out.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("out must be a channels-last fp32 (B, >= C_out, h, w) tensor")
    _lib.check(_lib.load().slimb200_corr_lookup_conv(pyramid.data_ptr(), _lib.DTYPE_BF16, C.byref(L), coords.data_ptr(), radius,
                                                     packed.packed.data_ptr(), c_out, 1 if relu else 0, out.data_ptr(),
                                                     int(out.shape[1]), _lib.current_stream_ptr()))
    return out


def lookup_conv_supported(L: _lib.CorrLayout, radius: int, c_out: int) -> bool:
    return radius == 3 and L.levels == 4 and c_out in (32, 64, 96)


def pack_pyramid_f32(levels: List[torch.Tensor], L: _lib.CorrLayout) -> torch.Tensor:
    """Pack reference-shaped fp32 levels (B*h*w,1,h_l,w_l) into the library's panel layout (tests)."""
    nf, P, pw = L.h * L.w, L.n_panels, _lib.PANEL_COLS
    rows = torch.zeros((L.batch, L.rows_padded, P * pw), dtype=torch.float32, device=levels[0].device)
    for l, lv in enumerate(levels):
        rows[:, :nf, L.level_offset[l]:L.level_offset[l] + L.level_h[l] * L.level_w[l]] = lv.reshape(L.batch, nf, -1)
    t = rows.view(L.batch, L.rows_padded // 128, 32, 4, P, 2, 8, 8)  # (b, m, g, p, panel, half, cb, e)
    return t.permute(0, 4, 1, 5, 2, 6, 3, 7).contiguous().view(-1, pw)  # (b, panel, m, half, g, cb, p, e)


class CorrBlock:
    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, radius: int = 4):
        self.num_levels = num_levels
        self.radius = radius
        self.pyramid = None
        self._ws = None
        self.rebuild(fmap1, fmap2)

    @_lib.on_device_of_args
    def rebuild(self, fmap1: torch.Tensor, fmap2: torch.Tensor) -> "CorrBlock":
        """(Re)compute the pyramid for a new pair of feature maps INTO the buffers this block already owns (same
        shapes), so that a captured CUDA graph of the lookups keeps pointing at valid data."""
        _lib.require_cuda(fmap1, fmap2)
        if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
            raise RuntimeError("liso_b200 CorrBlock is forward-only (flow export): call it under torch.no_grad()")
        if fmap1.shape != fmap2.shape or fmap1.dim() != 4:
            raise ValueError("fmap1/fmap2 must both be (B, D, h, w)")
        B, D, h, w = fmap1.shape
        lib = _lib.load()
        L = make_layout(B, D, h, w, self.num_levels)
        f1, f2 = fmap1.detach().float(), fmap2.detach().float()
        # feed the kernel whatever layout the feature encoder produced: channels-last needs no transposition
        nhwc = f1.is_contiguous(memory_format=torch.channels_last) and f2.is_contiguous(memory_format=torch.channels_last) \
            and not (f1.is_contiguous() and f2.is_contiguous())
        if not nhwc:
            f1, f2 = f1.contiguous(), f2.contiguous()
        layout_flag = _lib.CANVAS_NHWC if nhwc else _lib.CANVAS_NCHW
        shape = (B * L.n_panels * L.rows_padded, _lib.PANEL_COLS)
        if self.pyramid is None:
            self.pyramid = torch.empty(shape, dtype=torch.bfloat16, device=f1.device)
            self._ws = torch.empty(lib.slimb200_corr_workspace_bytes(C.byref(L)), dtype=torch.uint8, device=f1.device)
            self.channels_last = nhwc  # answer lookups in the memory format the feature maps came in
        elif tuple(self.pyramid.shape) != shape or self.pyramid.device != f1.device or self.channels_last != nhwc:
            raise ValueError("CorrBlock.rebuild needs feature maps of the shape / layout the block was built for")
        self.layout = L
        _lib.check(lib.slimb200_corr_build(f1.data_ptr(), f2.data_ptr(), layout_flag, C.byref(L), _lib.DTYPE_BF16,
                                           self.pyramid.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                           _lib.current_stream_ptr()))
        self.corr_pyramid = _LazyLevels(self.pyramid, L)
        return self

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        return lookup(self.pyramid, self.layout, coords, self.radius, self.channels_last)

    def lookup_conv(self, coords: torch.Tensor, weight, bias: torch.Tensor = None, relu: bool = True,
                    out: torch.Tensor = None) -> torch.Tensor:
        """``act(conv1x1(self(coords)))`` without materialising the lookup tensor (see :func:`lookup_conv`).  ``weight``: a
        :class:`PackedLookupConv` (what a caller inside a loop keeps), or the conv's weight tensor (packed on the spot)."""
        if not isinstance(weight, PackedLookupConv):
            weight = PackedLookupConv(weight, bias, self.num_levels, self.radius)
        return lookup_conv(self.pyramid, self.layout, coords, self.radius, weight, relu, out)

    def lookup_conv_supported(self, c_out: int) -> bool:
        return self.pyramid.dtype == torch.bfloat16 and lookup_conv_supported(self.layout, self.radius, c_out)

    @staticmethod
    def corr(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
        """All-pairs volume with the reference's return shape (B, h, w, 1, h, w), fp32 values."""
        B, _, h, w = fmap1.shape
        blk = CorrBlock(fmap1, fmap2, num_levels=1, radius=0)
        return blk.corr_pyramid[0].float().reshape(B, h, w, 1, h, w)


# ---------------------------------------------------------------- raft_code/utils.py (stock PyTorch)
def bilinear_sampler(img, coords, mode="bilinear", mask=False):
    """Pixel-coordinate wrapper around ``F.grid_sample`` (``raft_code/utils.py:15-29``)."""
    H, W = img.shape[-2:]
    xg, yg = coords.split([1, 1], dim=-1)
    xg = 2 * xg / (W - 1) - 1
    yg = 2 * yg / (H - 1) - 1
    out = F.grid_sample(img, torch.cat([xg, yg], dim=-1), align_corners=True)
    if mask:
        inside = (xg > -1) & (yg > -1) & (xg < 1) & (yg < 1)
        return out, inside.float()
    return out


def coords_grid(batch, ht, wd, device):
    """channel 0 = x (column), channel 1 = y (row) (``raft_code/utils.py:32-37``, ``raft_mod.py:134-135``)."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def initialize_flow(img, downscale_factor=8):
    n, _, H, W = img.shape
    return coords_grid(n, H // downscale_factor, W // downscale_factor, device=img.device)


def upflow_n(flow, n=8, mode="bilinear"):
    return n * F.interpolate(flow, size=(n * flow.shape[2], n * flow.shape[3]), mode=mode, align_corners=True)


def uplogits_n(logits, n=8, mode="bilinear"):
    return F.interpolate(logits, size=(n * logits.shape[2], n * logits.shape[3]), mode=mode, align_corners=True)
