"""Python side of the SURVEY 8(f).2 glue kernels (``csrc/gru_glue.cu``): the element-wise work between the stock
convolutions of the ConvGRU update block (``liso/slim/model/update.py:23-38,70-93,130-150``) and the refinement loop's
coordinate / logit update (``raft_mod.py:188-212``) on channels-last fp32 CUDA tensors.

Every function takes and returns ordinary ``(B, C, h, w)`` tensors in channels-last memory format; the library sees
them as ``(B*h*w, C)`` rows.  No fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import torch

from .. import _lib


def is_nhwc(t: torch.Tensor) -> bool:
    """Packed channels-last (B, C, h, w) fp32 CUDA tensor."""
    return (t.dim() == 4 and t.is_cuda and t.dtype == torch.float32
            and t.is_contiguous(memory_format=torch.channels_last))


def as_nhwc(t: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(t)
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if is_nhwc(t) else t.contiguous(memory_format=torch.channels_last)


def _pixels(t: torch.Tensor) -> int:
    return t.shape[0] * t.shape[2] * t.shape[3]


def _nhwc_source(t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """(tensor, row pitch in floats) of a source: a packed channels-last tensor or a channel slice ``u[:, a:b]`` of one
    (channel offset and width multiples of 4); anything else is copied to a packed channels-last tensor first."""
    _lib.require_cuda(t)
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    if is_nhwc(t):
        return t, t.shape[1]
    B, Cn, h, w = t.shape
    pitch = t.stride(3)
    if (t.dim() == 4 and t.stride(1) == 1 and pitch >= Cn and pitch % 4 == 0 and t.stride(2) == w * pitch
            and t.stride(0) == h * w * pitch and t.data_ptr() % 16 == 0):
        return t, pitch
    return t.contiguous(memory_format=torch.channels_last), Cn


@_lib.on_device_of_args
def nhwc_pack_into(srcs: Sequence[torch.Tensor], dsts: Sequence[Tuple[torch.Tensor, int]]) -> None:
    """Channel-concatenate ``srcs`` (<= 4; packed channels-last tensors or channel slices of such) and store the result at
    channel offset ``off`` of every ``(dst, off)`` (<= 2 destinations, packed channels-last tensors with
    >= off + sum(C_src) channels)."""
    pairs = [_nhwc_source(s) for s in srcs]
    srcs = [p[0] for p in pairs]
    B, _, h, w = srcs[0].shape
    for s in srcs:
        if (s.shape[0], s.shape[2], s.shape[3]) != (B, h, w):
            raise ValueError("nhwc_pack: sources must share batch and spatial size")
    for d, _ in dsts:
        if not is_nhwc(d) or (d.shape[0], d.shape[2], d.shape[3]) != (B, h, w):
            raise ValueError("nhwc_pack: destinations must be packed channels-last fp32 CUDA tensors of the sources' size")
    n_s, n_d = len(srcs), len(dsts)
    sp = (C.c_void_p * n_s)(*[s.data_ptr() for s in srcs])
    sc = (C.c_int32 * n_s)(*[s.shape[1] for s in srcs])
    spitch = (C.c_int32 * n_s)(*[p[1] for p in pairs])
    dp = (C.c_void_p * n_d)(*[d.data_ptr() for d, _ in dsts])
    do = (C.c_int32 * n_d)(*[int(o) for _, o in dsts])
    dpitch = (C.c_int32 * n_d)(*[d.shape[1] for d, _ in dsts])
    _lib.check(_lib.load().slimb200_nhwc_pack(sp, sc, spitch, n_s, dp, do, dpitch, n_d, B * h * w, _lib.current_stream_ptr()))


def nhwc_cat(srcs: Sequence[torch.Tensor]) -> torch.Tensor:
    """``torch.cat(srcs, dim=1)`` for channels-last tensors, one launch."""
    s0 = srcs[0]
    out = torch.empty((s0.shape[0], sum(s.shape[1] for s in srcs), s0.shape[2], s0.shape[3]), dtype=torch.float32,
                      device=s0.device, memory_format=torch.channels_last)
    nhwc_pack_into(srcs, [(out, 0)])
    return out


@_lib.on_device_of_args
def gru_gate_zr(zr_raw: torch.Tensor, bias_zr: torch.Tensor, hx: torch.Tensor, rhx: torch.Tensor, hidden: int) -> torch.Tensor:
    """``z = sigmoid(zr_raw[:, :hidden] + b)`` (returned, packed) and ``rhx[:, :hidden] = sigmoid(zr_raw[:, hidden:] + b) *
    hx[:, :hidden]`` (``update.py:33-35``); ``zr_raw`` is the stacked update|reset convolution WITHOUT its bias."""
    zr_raw = as_nhwc(zr_raw)
    if not (is_nhwc(hx) and is_nhwc(rhx)) or zr_raw.shape[1] != 2 * hidden or bias_zr.numel() != 2 * hidden:
        raise ValueError("gru_gate_zr: bad buffers")
    z = torch.empty((zr_raw.shape[0], hidden, zr_raw.shape[2], zr_raw.shape[3]), dtype=torch.float32, device=zr_raw.device,
                    memory_format=torch.channels_last)
    _lib.check(_lib.load().slimb200_gru_gate_zr(zr_raw.data_ptr(), bias_zr.data_ptr(), hx.data_ptr(), hx.shape[1], z.data_ptr(),
                                                rhx.data_ptr(), rhx.shape[1], hidden, _pixels(zr_raw), _lib.current_stream_ptr()))
    return z


@_lib.on_device_of_args
def gru_gate_out(q_raw: torch.Tensor, bias_q: torch.Tensor, z: torch.Tensor, hx: torch.Tensor, hidden: int) -> torch.Tensor:
    """``h' = (1 - z) * h + z * tanh(q_raw + b)`` (``update.py:35-38``) written into ``hx[:, :hidden]`` in place and
    returned as a packed tensor for the heads."""
    q_raw, z = as_nhwc(q_raw), as_nhwc(z)
    if not is_nhwc(hx) or q_raw.shape[1] != hidden or z.shape != q_raw.shape or bias_q.numel() != hidden:
        raise ValueError("gru_gate_out: bad buffers")
    h = torch.empty_like(q_raw)
    _lib.check(_lib.load().slimb200_gru_gate_out(q_raw.data_ptr(), bias_q.data_ptr(), z.data_ptr(), hx.data_ptr(), hx.shape[1],
                                                 h.data_ptr(), hidden, _pixels(q_raw), _lib.current_stream_ptr()))
    return h


def _head_strides(t: torch.Tensor) -> Tuple[torch.Tensor, int, int, int]:
    """(tensor, batch stride, channel stride, pixel stride) of a (B, C, h, w) head output: NCHW- or channels-last-dense,
    or a channel slice of such a tensor (the two heads evaluated as one stacked convolution)."""
    _lib.require_cuda(t)
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    _, _, h, w = t.shape
    if t.stride(3) * w == t.stride(2) and (h == 1 or t.stride(2) > 0):  # pixels are evenly spaced: pix * stride(3)
        return t, t.stride(0), t.stride(1), t.stride(3)
    t = t.contiguous()
    return t, t.stride(0), t.stride(1), 1


def _is_nhwc(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def _stacked_channels(stacked) -> int:
    """The library's layout convention for the optional [flow | logits] copy: > 0 planar, < 0 channels-last."""
    if stacked is None:
        return 0
    if stacked.dtype != torch.float32 or not stacked.is_cuda:
        raise ValueError("stacked must be an fp32 CUDA tensor")
    return -int(stacked.shape[1]) if _is_nhwc(stacked) else int(stacked.shape[1])


@_lib.on_device_of_args
def iter_update(dflow_raw: torch.Tensor, bias_flow: torch.Tensor, dlogits_raw: torch.Tensor, bias_logits: torch.Tensor,
                coords1: torch.Tensor, flow: torch.Tensor, logits: torch.Tensor, stacked: torch.Tensor = None) -> None:
    """In place (``raft_mod.py:205-212``): ``coords1 += dflow_raw + b``; ``logits += dlogits_raw + b``;
    ``flow = coords1 - coords_grid``; optionally ``stacked = cat[flow, logits]``.  The raw tensors are the head
    convolutions without bias."""
    B, _, h, w = coords1.shape
    for t in (coords1, flow, logits) + ((stacked,) if stacked is not None and not _is_nhwc(stacked) else ()):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("iter_update: coords1 / flow / logits / stacked must be contiguous fp32 CUDA tensors")
    if tuple(dflow_raw.shape) != (B, 2, h, w) or tuple(dlogits_raw.shape) != tuple(logits.shape) or tuple(flow.shape) != (B, 2, h, w):
        raise ValueError("iter_update: shape mismatch")
    if stacked is not None and (stacked.shape[1] < 2 + logits.shape[1] or (stacked.shape[0], stacked.shape[2], stacked.shape[3]) != (B, h, w)):
        raise ValueError("iter_update: stacked must be (B, >= 2 + n_logits, h, w)")
    df, bs_f, cs_f, ps_f = _head_strides(dflow_raw)
    dl, bs_l, cs_l, ps_l = _head_strides(dlogits_raw)
    _lib.check(_lib.load().slimb200_iter_update(df.data_ptr(), bs_f, cs_f, ps_f, bias_flow.data_ptr(), dl.data_ptr(), bs_l, cs_l,
                                                ps_l, bias_logits.data_ptr(), logits.shape[1], B, h, w, coords1.data_ptr(),
                                                flow.data_ptr(), logits.data_ptr(),
                                                stacked.data_ptr() if stacked is not None else None,
                                                _stacked_channels(stacked), _lib.current_stream_ptr()))


@_lib.on_device_of_args
def iter_update_taps(taps: torch.Tensor, ksize: int, bias_flow: torch.Tensor, bias_logits: torch.Tensor, coords1: torch.Tensor,
                     flow: torch.Tensor, logits: torch.Tensor, stacked: torch.Tensor = None) -> None:
    """`iter_update` with the heads' k x k output convolution given as the 1x1 "tap" tensor (B, k*k*(2 + n_logits), h, w),
    channel = tap * (2 + n_logits) + c: the window sum of the taps is taken inside the kernel."""
    B, _, h, w = coords1.shape
    for t in (coords1, flow, logits) + ((stacked,) if stacked is not None and not _is_nhwc(stacked) else ()):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("iter_update_taps: coords1 / flow / logits / stacked must be contiguous fp32 CUDA tensors")
    nl = logits.shape[1]
    taps = as_nhwc(taps)
    if tuple(taps.shape) != (B, ksize * ksize * (2 + nl), h, w) or tuple(flow.shape) != (B, 2, h, w) or (
            stacked is not None and (stacked.shape[1] < 2 + nl or (stacked.shape[0], stacked.shape[2], stacked.shape[3]) != (B, h, w))):
        raise ValueError("iter_update_taps: shape mismatch")
    _lib.check(_lib.load().slimb200_iter_update_taps(taps.data_ptr(), ksize, bias_flow.data_ptr(), bias_logits.data_ptr(), nl, B,
                                                     h, w, coords1.data_ptr(), flow.data_ptr(), logits.data_ptr(),
                                                     stacked.data_ptr() if stacked is not None else None,
                                                     _stacked_channels(stacked), _lib.current_stream_ptr()))


@_lib.on_device_of_args
def add_relu(x: torch.Tensor, y: torch.Tensor, bias_x: torch.Tensor = None) -> torch.Tensor:
    """``relu(x + y)`` for two fp32 CUDA tensors of the same shape and memory layout (``extractor.py:57-68``); with
    ``bias_x`` (per channel, channels-last tensors) ``relu((x + bias_x) + y)``: ``x`` is then the projection shortcut
    convolved without its bias."""
    _lib.require_cuda(x, y, bias_x)
    if x.shape != y.shape or x.stride() != y.stride() or x.dtype != torch.float32 or y.dtype != torch.float32 \
            or not (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)) or x.numel() % 4:
        raise ValueError("add_relu: tensors must be dense fp32 with identical shape and strides, numel % 4 == 0")
    out = torch.empty_like(x)
    if bias_x is not None:
        C = int(x.shape[1])
        if not x.is_contiguous(memory_format=torch.channels_last) or C % 4 or bias_x.numel() != C or bias_x.dtype != torch.float32:
            raise ValueError("add_relu with bias_x: channels-last tensors with C % 4 == 0 and an fp32 bias of C elements")
        b = bias_x.detach().contiguous()
        _lib.check(_lib.load().slimb200_add_bias_relu(x.data_ptr(), b.data_ptr(), C, y.data_ptr(), out.data_ptr(), x.numel(),
                                                      _lib.current_stream_ptr()))
        return out
    _lib.check(_lib.load().slimb200_add_relu(x.data_ptr(), y.data_ptr(), out.data_ptr(), x.numel(), _lib.current_stream_ptr()))
    return out


@_lib.on_device_of_args
def ctx_split(raw: torch.Tensor, bias: torch.Tensor, hidden: int, context: int):
    """Context-encoder tail (``raft_mod.py:170-173``): ``(tanh(raw[:, :hidden] + b), relu(raw[:, hidden:] + b))`` from the
    bias-free output of cnet's last convolution (channels-last), as two packed channels-last tensors."""
    _lib.require_cuda(raw, bias)
    B, Cn, h, w = raw.shape
    if Cn != hidden + context or hidden % 4 or context % 4 or raw.dtype != torch.float32 or bias.numel() != Cn \
            or not raw.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("ctx_split: channels-last fp32 (B, hidden + context, h, w) with hidden % 4 == context % 4 == 0")
    net = torch.empty((B, hidden, h, w), dtype=torch.float32, device=raw.device, memory_format=torch.channels_last)
    inp = torch.empty((B, context, h, w), dtype=torch.float32, device=raw.device, memory_format=torch.channels_last)
    _lib.check(_lib.load().slimb200_ctx_split(raw.data_ptr(), bias.detach().contiguous().data_ptr(), hidden, context, B * h * w,
                                              net.data_ptr(), inp.data_ptr(), _lib.current_stream_ptr()))
    return net, inp


@_lib.on_device_of_args
def bias_relu_slice(x: torch.Tensor, c0: int, channels: int, bias: torch.Tensor) -> torch.Tensor:
    """``relu(x[:, c0 : c0 + channels] + bias)`` of a channels-last tensor as a packed channels-last tensor: one half of two
    parallel convolutions evaluated as one (without bias)."""
    _lib.require_cuda(x, bias)
    B, Cx, h, w = x.shape
    if c0 % 4 or channels % 4 or c0 + channels > Cx or x.dtype != torch.float32 or bias.numel() != channels \
            or not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("bias_relu_slice: channels-last fp32 tensor, slice bounds multiples of 4, bias of `channels` elements")
    out = torch.empty((B, channels, h, w), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    _lib.check(_lib.load().slimb200_bias_relu_slice(x.data_ptr() + 4 * c0, Cx, bias.detach().contiguous().data_ptr(), channels,
                                                    B * h * w, out.data_ptr(), _lib.current_stream_ptr()))
    return out


def add_relu_ok(x: torch.Tensor, y: torch.Tensor) -> bool:
    return (x.is_cuda and y.is_cuda and x.dtype == torch.float32 and y.dtype == torch.float32 and x.shape == y.shape
            and x.stride() == y.stride() and (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last))
            and x.numel() % 4 == 0 and not torch.is_grad_enabled())
