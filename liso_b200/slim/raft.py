"""RAFT wiring of SLIM around the two B200 kernels stages (reference ``liso/slim/model/raft_mod.py``).

Only the pillar encoder (``pp_layer``) and the correlation block are custom CUDA; the feature /
context encoders and the ConvGRU update block stay stock PyTorch (cuDNN) as the scope asks.  Their
module and parameter names follow the reference (``extractor.py``, ``update.py``) so that a
reference checkpoint loads with ``strict=True`` (``experiment.py:221-223``).
"""
from __future__ import annotations

import weakref
from typing import List

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..networks.pcl_to_feature_grid import PointsPillarFeatureNetWrapper
from .corr import CorrBlock, coords_grid, initialize_flow, uplogits_n, upflow_n


from .. import _lib as _lib_module  # noqa: E402


def _lib_mod():
    from .. import _lib

    return _lib


# Stock-op plumbing (no custom kernels): when True, Conv2d -> ReLU pairs go through cuDNN's fused
# conv + bias + activation (``torch.cudnn_convolution_relu``) instead of conv, bias-add and clamp as three
# launches, and affine InstanceNorm on channels-last tensors is computed with var_mean / addcmul so that the
# tensor is not copied to NCHW and back for cuDNN's batch-norm kernel.  Same math, fp32 / TF32 as configured.
FAST_STOCK_OPS = True


def _derived(owner: nn.Module, key, sources, build):
    """A tensor derived from parameters (a cast, a concatenation, stacked or tap weights), cached ON the module that owns
    the sources (inference only).  An entry is valid for exactly these tensor OBJECTS (weak references, not addresses: a
    freed model's parameters may be re-allocated at the same address with the same version) in their current version and
    storage; it dies with the module."""
    cache = owner.__dict__.setdefault("_slimb200_derived", {})
    sig = tuple((s._version, s.data_ptr(), s.device, tuple(s.stride())) for s in sources)
    ent = cache.get(key)
    if ent is not None and ent[1] == sig and len(ent[0]) == len(sources) and all(r() is s for r, s in zip(ent[0], sources)):
        return ent[2]
    val = build()
    cache[key] = (tuple(weakref.ref(s) for s in sources), sig, val)
    return val


def _derived_values(root: nn.Module):
    """Every derived tensor currently cached below `root` (a captured CUDA graph points at them)."""
    out = []
    for mod in root.modules():
        for ent in mod.__dict__.get("_slimb200_derived", {}).values():
            out.append(ent[2])
    return out


def _autocast_param(owner: nn.Module, name: str, t, dtype):
    """Low-precision copy of a parameter, cached like autocast caches its weight casts (inference only)."""
    if t is None:
        return None

    def build():
        hit = t.detach().to(dtype)
        if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
            hit = hit.contiguous(memory_format=torch.channels_last)
        return hit

    return _derived(owner, ("cast", name, dtype), (t,), build)


def _cat_params(owner: nn.Module, name: str, ts):
    """dim-0 concatenation of parameters, cached on `owner` (inference only)."""

    def build():
        hit = torch.cat([t.detach() for t in ts], dim=0)
        if hit.dim() == 4 and ts[0].is_contiguous(memory_format=torch.channels_last) and not ts[0].is_contiguous():
            hit = hit.contiguous(memory_format=torch.channels_last)
        return hit

    return _derived(owner, ("cat", name), tuple(ts), build)


def conv_relu(conv: nn.Conv2d, x: torch.Tensor, weight: torch.Tensor = None, bias: torch.Tensor = None) -> torch.Tensor:
    """relu(conv(x)); ``weight`` / ``bias`` override the module's own (stacked parallel convolutions, same geometry)."""
    w, b = (conv.weight, conv.bias) if weight is None else (weight, bias)
    if FAST_STOCK_OPS and x.is_cuda and conv.padding_mode == "zeros" and not torch.is_grad_enabled():
        if torch.is_autocast_enabled():  # the fused op is not on autocast's cast list
            dt = torch.get_autocast_dtype("cuda")
            tag = "own" if weight is None else "override"
            x, w, b = x.to(dt), _autocast_param(conv, tag + ".w", w, dt), _autocast_param(conv, tag + ".b", b, dt)
        if x.dtype == w.dtype:
            return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    return F.relu(F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups))


def _same_geometry(a: nn.Conv2d, b: nn.Conv2d) -> bool:
    return (a.kernel_size == b.kernel_size and a.stride == b.stride and a.padding == b.padding and a.dilation == b.dilation
            and a.groups == b.groups == 1 and a.padding_mode == b.padding_mode == "zeros")


def _stacked_params(a: nn.Conv2d, b: nn.Conv2d, shared_input: bool, pad_in_to: int = 0):
    """Weight / bias of ONE convolution that evaluates the parallel convolutions ``a`` and ``b`` (same geometry): outputs
    stacked [a | b]; with ``shared_input`` both read the same tensor, otherwise the input is the channel concatenation
    [in_a | in_b] (zero-padded to ``pad_in_to`` channels) and the weight is block-diagonal (the zero blocks add exact
    zeros).  Cached on ``a`` like `_cat_params`."""

    def build():
        wa, wb = a.weight.detach(), b.weight.detach()
        if shared_input:
            w = torch.cat([wa, wb], dim=0)
        else:
            w = wa.new_zeros((wa.shape[0] + wb.shape[0], max(wa.shape[1] + wb.shape[1], pad_in_to)) + tuple(wa.shape[2:]))
            w[:wa.shape[0], :wa.shape[1]] = wa
            w[wa.shape[0]:, wa.shape[1]:wa.shape[1] + wb.shape[1]] = wb
        if wa.is_contiguous(memory_format=torch.channels_last) and not wa.is_contiguous():
            w = w.contiguous(memory_format=torch.channels_last)
        return w, torch.cat([a.bias.detach(), b.bias.detach()], dim=0)

    return _derived(a, ("stacked", shared_input, pad_in_to), (a.weight, a.bias, b.weight, b.bias), build)


@_lib_module.on_device_of_args
def instance_norm_nhwc(norm: nn.InstanceNorm2d, x: torch.Tensor, relu: bool, residual: torch.Tensor = None,
                       inplace: bool = False, channel_slice=None) -> torch.Tensor:
    """Affine InstanceNorm2d (+ ReLU) of a channels-last CUDA tensor in the library's glue kernel
    (``slimb200_instnorm_nhwc``): 3 launches / 2 passes instead of copy-to-NCHW + cuDNN batch-norm + copy back + clamp.
    With ``residual`` the block's join is fused into the same pass: ``relu(residual + [relu](norm(x)))``.
    ``inplace`` overwrites ``x`` (a convolution output nobody else reads): half the L2 footprint of the apply pass.
    ``channel_slice=(c0, C)``: normalise channels [c0, c0 + C) of ``x`` only, into a packed C-channel tensor."""
    lib = _lib_mod().load()
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    B, Cx, H, W = x.shape
    c0, Cn = (0, Cx) if channel_slice is None else channel_slice
    if channel_slice is not None:
        assert not inplace and c0 % 4 == 0 and Cn % 4 == 0 and c0 + Cn <= Cx
        out = torch.empty((B, Cn, H, W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    else:
        out = x if inplace else torch.empty_like(x)  # (empty_like preserves the channels-last strides)
    ws = torch.empty(max(256, lib.slimb200_instnorm_workspace_bytes(B, Cn, H * W)), dtype=torch.uint8, device=x.device)
    flags = (1 if relu else 0) | (2 if residual is not None else 0)
    _lib_mod().check(lib.slimb200_instnorm_nhwc_slice(
        x.data_ptr() + 4 * c0, Cx, norm.weight.data_ptr(), norm.bias.data_ptr(), float(norm.eps), B, H, W, Cn, flags,
        residual.data_ptr() if residual is not None else None, out.data_ptr(), ws.data_ptr(), ws.numel(),
        _lib_mod().current_stream_ptr()))
    return out


def _fused_norm_ok(norm: nn.Module, x: torch.Tensor, channels: int = None) -> bool:
    """``x``: the tensor to normalise, or (with ``channels``) the channels-last input of the convolution that produces it."""
    c = x.shape[1] if channels is None else channels
    return (FAST_STOCK_OPS and isinstance(norm, nn.InstanceNorm2d) and norm.affine and not norm.track_running_stats
            and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and not torch.is_autocast_enabled()
            and c % 4 == 0 and c <= 256 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last))


def conv_norm(conv: nn.Conv2d, norm: nn.Module, x: torch.Tensor, relu: bool, residual: torch.Tensor = None) -> torch.Tensor:
    """[relu](norm(conv(x))), or with ``residual`` the whole tail of a residual block relu(residual + [relu](norm(conv(x))))
    (``extractor.py:57-68``).  In front of an InstanceNorm the convolution's bias is a per-channel constant that the
    normalisation subtracts again, so on the fused path the bias add (a full pass over the tensor) is skipped: equal
    in exact arithmetic, ~1 ulp different in fp32."""
    if _fused_norm_ok(norm, x, conv.out_channels):
        y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        if residual is None or (residual.shape == y.shape and residual.stride() == y.stride() and residual.dtype == y.dtype
                                and residual.is_cuda):
            return instance_norm_nhwc(norm, y, relu=relu, residual=residual, inplace=True)
        return add_relu(residual, instance_norm_nhwc(norm, y, relu=relu, inplace=True))
    y = norm(conv(x))
    y = F.relu(y) if relu else y
    return y if residual is None else add_relu(residual, y)


def add_relu(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """relu(x + y): one glue launch on dense fp32 CUDA tensors of equal layout, the two stock ops otherwise."""
    if FAST_STOCK_OPS and _glue().add_relu_ok(x, y):
        return _glue().add_relu(x, y)
    return F.relu(x + y)


def _glue():
    from . import glue

    return glue


def _is_identity(norm: nn.Module) -> bool:
    return isinstance(norm, nn.Sequential) and len(norm) == 0


def _make_norm(kind: str, channels: int) -> nn.Module:
    if kind == "instance_affine":
        return nn.InstanceNorm2d(channels, eps=1e-3, affine=True)
    if kind == "none":
        return nn.Sequential()
    if kind == "instance":
        return nn.InstanceNorm2d(channels)
    if kind == "batch":
        return nn.BatchNorm2d(channels)
    if kind == "group":
        return nn.GroupNorm(num_groups=channels // 8, num_channels=channels)
    raise ValueError("unknown norm %r" % kind)


class ResidualBlock(nn.Module):
    """``extractor.py:5-68``.  The projection shortcut is created from the *stage's* input width
    (``dummy_in_filters``), so the second block of a widening stage also gets one (and the
    reference registers its norm twice: ``norm3`` and ``downsample.1``)."""

    def __init__(self, in_filters, out_filters, dummy_in_filters, norm_fn="group", stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_filters, out_filters, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(out_filters, out_filters, kernel_size=3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = _make_norm(norm_fn, out_filters)
        self.norm2 = _make_norm(norm_fn, out_filters)
        self.downsample = None
        if not (stride == 1 and dummy_in_filters == out_filters):
            self.norm3 = _make_norm(norm_fn, out_filters)
            self.downsample = nn.Sequential(nn.Conv2d(in_filters, out_filters, kernel_size=1, stride=stride), self.norm3)

    def forward(self, x):
        if _is_identity(self.norm1):
            y = conv_relu(self.conv2, conv_relu(self.conv1, x))
            if self.downsample is not None:
                d = self.downsample[0]
                if (FAST_STOCK_OPS and d.bias is not None and d.out_channels % 4 == 0 and d.padding_mode == "zeros" and x.is_cuda
                        and x.dtype == torch.float32 and not torch.is_grad_enabled() and not torch.is_autocast_enabled()
                        and y.is_contiguous(memory_format=torch.channels_last) and not y.is_contiguous()):
                    # the shortcut's bias joins in the residual kernel instead of in ATen's separate bias pass over the tensor
                    xr = F.conv2d(x, d.weight, None, d.stride, d.padding, d.dilation, d.groups)
                    if _glue().add_relu_ok(xr, y):
                        return _glue().add_relu(xr, y, bias_x=d.bias)
                    x = xr + d.bias[None, :, None, None]
                else:
                    x = self.downsample(x)
            return add_relu(x, y)
        y = conv_norm(self.conv1, self.norm1, x, relu=True)
        if self.downsample is not None:
            x = conv_norm(self.downsample[0], self.downsample[1], x, relu=False)
        return conv_norm(self.conv2, self.norm2, y, relu=True, residual=x)  # relu(x + relu(norm2(conv2(y))))


class SmallEncoder(nn.Module):
    """``extractor.py:211-297``: 7x7/2 stem, three residual stages (32, 64/2, 96/2), 1x1 head."""

    def __init__(self, output_dim=128, norm_fn="batch", dropout=0.0):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = _make_norm(norm_fn, 32)
        self.conv1 = nn.Conv2d(64, 32, kernel_size=7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        self.layer1 = self._stage(32, 32, 1)
        self.layer2 = self._stage(32, 64, 2)
        self.layer3 = self._stage(64, 96, 2)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        self.conv2 = nn.Conv2d(96, output_dim, kernel_size=1)

    def _stage(self, cin, cout, stride):
        return nn.Sequential(
            ResidualBlock(cin, cout, dummy_in_filters=cin, norm_fn=self.norm_fn, stride=stride),
            ResidualBlock(cout, cout, dummy_in_filters=cin, norm_fn=self.norm_fn, stride=1),
        )

    def forward(self, x, head_bias: bool = True, stem_out: torch.Tensor = None):
        """``head_bias=False``: the output convolution is evaluated without its bias (the caller adds it in a kernel of
        its own, see ``RAFT._context``); not available with dropout.  ``stem_out``: the result of the stem
        (relu(norm1(conv1(x)))) when the caller has computed it already (``RAFT._stems``)."""
        if stem_out is not None:
            x = stem_out
        else:
            x = conv_relu(self.conv1, x) if _is_identity(self.norm1) else conv_norm(self.conv1, self.norm1, x, relu=True)
        x = self.layer3(self.layer2(self.layer1(x)))
        if not head_bias:
            assert not (self.training and self.dropout is not None)
            c = self.conv2
            return F.conv2d(x, c.weight, None, c.stride, c.padding, c.dilation, c.groups)
        x = self.conv2(x)
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return x


class FlowOrClassificationHead(nn.Module):
    def __init__(self, input_dim=128, hidden_dim=256, out_dims=2):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, out_dims, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(conv_relu(self.conv1, x))


class ConvGRU(nn.Module):
    def __init__(self, hidden_dim=96, input_dim=304):
        super().__init__()
        self.convz = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.convr = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.convq = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)

    def forward(self, h, x):
        hx = torch.cat([h, x], dim=1)
        if FAST_STOCK_OPS and hx.is_cuda and not torch.is_grad_enabled():
            # the update and reset gates read the same input: one 192-channel convolution instead of two
            w = _cat_params(self, "zr.weight", (self.convz.weight, self.convr.weight))
            b = _cat_params(self, "zr.bias", (self.convz.bias, self.convr.bias))
            z, r = torch.sigmoid(F.conv2d(hx, w, b, padding=1)).split(self.convz.out_channels, dim=1)
        else:
            z = torch.sigmoid(self.convz(hx))
            r = torch.sigmoid(self.convr(hx))
        q = torch.tanh(self.convq(torch.cat([r * h, x], dim=1)))
        return (1 - z) * h + z * q


class SmallMotionEncoder(nn.Module):
    """``update.py:41-93`` for ``flow_maps_archi != 'vanilla'`` and no static-aggregation weights."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        cc = cfg.model.corr_cfg
        self.conv_stat_corr1 = nn.Conv2d(cc.num_levels * (2 * cc.search_radius + 1) ** 2, 96, 1, padding=0)
        self.conv_flow1 = nn.Conv2d(2, 64, 7, padding=3)
        self.conv_flow2 = nn.Conv2d(64, 32, 3, padding=1)
        self.conv_class1 = nn.Conv2d(4, 64, 7, padding=3)
        self.conv_class2 = nn.Conv2d(64, 32, 3, padding=1)
        self.conv = nn.Conv2d(160, 80, 3, padding=1)

    def forward(self, flow, corr, logits):
        c = conv_relu(self.conv_stat_corr1, corr)
        f = conv_relu(self.conv_flow2, conv_relu(self.conv_flow1, flow))
        lg = conv_relu(self.conv_class2, conv_relu(self.conv_class1, logits))
        out = conv_relu(self.conv, torch.cat([c, f, lg], dim=1))
        return torch.cat([out, lg, f], dim=1)


class SmallUpdateBlock(nn.Module):
    def __init__(self, cfg, filters=96):
        super().__init__()
        if cfg.model.flow_maps_archi == "vanilla" or cfg.model.predict_weight_for_static_aggregation is not False:
            raise NotImplementedError("only the released SLIM configuration (single flow map, no aggregation weights)")
        self.cfg = cfg
        self.filters = filters
        self.motion_encoder = SmallMotionEncoder(cfg)
        self.gru = ConvGRU(hidden_dim=filters, input_dim=304)
        self.static_flow_head = FlowOrClassificationHead(filters, 128, 2)
        self.classification_head = FlowOrClassificationHead(filters, 128, 4)

    def forward(self, net, inp, corr, flow, logits, weight_logits_for_static_aggregation=None):
        motion = self.motion_encoder(flow, corr, logits)
        net = self.gru(net, torch.cat([inp, motion], dim=1))
        return net, self.static_flow_head(net), self.classification_head(net), None


@_lib_module.on_device_of_args
def raft_output_fused(flow, logits, n, res_rows, res_cols):
    """upflow_n + uplogits_n + flip / scale + concat2network_output in one kernel (``slimb200_raft_output``,
    SURVEY 8f.2); the returned (B, H, W, 8) tensor carries the logit minimum the decoder needs."""
    from .. import _lib

    lib = _lib.load()
    B, _, h, w = flow.shape
    flow = flow.detach().float().contiguous()
    logits = logits.detach().float().contiguous()
    out = torch.empty((B, h * n, w * n, 8), dtype=torch.float32, device=flow.device)
    min_key = torch.empty((1,), dtype=torch.int32, device=flow.device)
    _lib.check(lib.slimb200_raft_output(flow.data_ptr(), logits.data_ptr(), B, h, w, n, float(res_rows), float(res_cols),
                                        out.data_ptr(), min_key.data_ptr(), _lib.current_stream_ptr()))
    out._slimb200_min_key = min_key
    return out


def concat2network_output(logits, static_flow, dynamic_flow):
    """``HeadDecoder.concat2network_output`` (``head_decoder.py:36-64``): (B,H,W,8) channels-last."""
    return torch.cat([logits, static_flow, dynamic_flow], dim=1).permute(0, 2, 3, 1)


class RAFT(nn.Module):
    def __init__(self, cfg, head_decoder_fw=None, head_decoder_bw=None):
        super().__init__()
        self.cfg = cfg
        self.slim_cfg = cfg.SLIM
        m = self.slim_cfg.model
        self.head_decoder_fw = head_decoder_fw
        self.head_decoder_bw = head_decoder_bw
        self.iters = m.num_iters
        # "all": one (B,H,W,8) output per GRU iteration like the reference (raft_mod.py:216-257); "last": only the
        # final one is up-sampled / assembled (the flow export reads nothing else, experiment.py:391-399)
        self.output_iterations = "all"
        # capture everything between the pillar encoder and the decoder in a CUDA graph (inference on CUDA only; the same
        # kernels are launched eagerly when switched off)
        self.use_cuda_graph = True
        # glue-kernel version of the update block's element-wise work (channels-last fp32 inference only)
        self.fused_update_block = True
        # ... and its pairs of parallel convolutions (flow / logits branches of the motion encoder, the two heads) stacked
        # into one cuDNN launch each
        self.merge_parallel_convs = True
        self.tap_heads = True  # ... and the heads' 3x3 output convolution as a 1x1 convolution to taps + window sum
        # the lookup fused with the 1x1 convolution that consumes it (conv_stat_corr1 + ReLU, update.py:49,71) in one
        # tcgen05 kernel with tf32 operands (slimb200_corr_lookup_conv): True = whenever TF32 convolutions are
        # allowed (cuDNN's precision for that layer then), "always" = also with fp32 convolutions, False (default) = lookup
        # kernel + stock convolution.  Off by default because it does not pay yet: inside the step both variants run at the
        # same pairs/s (the fused kernel holds an SM per CTA and cannot overlap with the other
        # direction's kernels the way the small lookup CTAs do); bench.py measures both in every run.
        self.fuse_lookup_conv = False
        # forward and backward direction as two parallel branches of the CUDA graph
        self.concurrent_directions = True
        self.stacked_stems = True  # fnet / cnet 7x7 stems on the same canvas as one stacked convolution
        self.direction_branches = 2  # parallel branches of the CUDA graph the directions are dealt over
        self.pre_replay_event = None  # see forward_frames
        self.last_forward_was_graph = False
        self.batched_encoders = True  # fnet / cnet over all frames of a pass at once (canvases = slices of one buffer)
        self.batched_frame_encoding = True  # every frame of a pass through one pillar-encoder call (eval-mode BatchNorm)
        # consumer of the per-iteration network outputs, e.g. SLIM's output decoder: called as
        # output_sink(direction, iteration, net_out, occupancy) right behind the kernel that wrote `net_out`.  While the
        # CUDA graph is captured it runs on a forked stream, i.e. the decodes become side branches of the graph that
        # overlap the following GRU iterations; output_sink_begin() announces a new pass (eager call or capture).
        self.output_sink = None
        self.output_sink_begin = None
        self.graph_extra_key = None  # whatever else the captured graph depends on (the sink's static buffers)
        self._streams = {}
        self._graphs = {}
        rows = float(cfg.data.bev_range_m[0]) / cfg.data.img_grid_size[0] * m.u_net.final_scale
        cols = float(cfg.data.bev_range_m[1]) / cfg.data.img_grid_size[1] * m.u_net.final_scale
        assert rows == cols, "anisotropic BEV resolution is not supported (raft_mod.py:42-45)"
        self.bev_rows_res_meters_per_fs_pixel = rows
        self.bev_cols_res_meters_per_fs_pixel = cols
        if m.corr_cfg.module != "all" or m.feature_downsampling_factor != 8:
            raise ValueError("only the all-pairs correlation block at 1/8 resolution exists (raft_mod.py:50-58)")
        self.pp_layer = PointsPillarFeatureNetWrapper(cfg)
        self.hidden_dim, self.context_dim = 96, 64
        self.fnet = SmallEncoder(output_dim=128, norm_fn=m.raft_fnet_norm, dropout=m.dropout_rate)
        self.cnet = SmallEncoder(output_dim=self.hidden_dim + self.context_dim, norm_fn="none", dropout=m.dropout_rate)
        self.update_block = SmallUpdateBlock(cfg=self.slim_cfg, filters=self.hidden_dim)

    def forward(self, pcl_t0, pcl_t1, raw_scans: bool = False):
        """``raw_scans``: the clouds are raw scans with ground (``liso_b200.datasets.preprocess_scans``); the encoder
        applies the dataset's ground rule itself."""
        outs, occs = self.forward_frames([pcl_t0, pcl_t1], [(0, 1)], raw_scans)
        return outs[0], outs[1], {"t0": {"bev_net_input_dbg": occs[0]}, "t1": {"bev_net_input_dbg": occs[1]}}

    def forward_frames(self, pcls, pairs, raw_scans: bool = False):
        """Several frame pairs over a common set of frames in one pass: ``pcls[f]`` = the cloud batch of frame f (a list of
        (N_i, 3|4) tensors), ``pairs`` = [(a, b), ...].  Returns (outs, occupancies): ``outs[2p]`` = the network outputs of
        a -> b of pair p, ``outs[2p + 1]`` = b -> a (each a list over GRU iterations of (B, H, W, 8)), ``occupancies[f]``
        per frame.  Every frame is encoded ONCE -- pillar encoder, feature encoder, context encoder -- however many pairs
        it takes part in: the reference's KITTI / nuScenes export runs t0 -> t1, t0 -> t2 and t1 -> t2 as three model()
        calls, i.e. twelve encoder passes for three frames (``experiment.py:386-456``)."""
        kw = {"raw_scan": True} if raw_scans else {}  # (the reference signature is forward(pcl, img=None))
        pairs = [(int(a), int(b)) for a, b in pairs]
        dev = pcls[0][0].device
        B = len(pcls[0])
        # all frames through ONE pillar-encoder call (5 launches whatever the number of frames: the key / scan / rank
        # kernels in front of the canvas writer are latency-bound, so their cost is per call, not per frame).  Not in
        # train-mode BatchNorm, whose batch statistics are per reference call (pcl_to_feature_grid.py:86-107).
        one_call = (self.batched_frame_encoding and not self.pp_layer.training and all(len(p) == B for p in pcls)
                    and B * len(pcls) <= _lib_module.MAX_BATCH and hasattr(self.pp_layer, "empty_outputs"))
        if not self.will_use_graph(*pcls):
            self.last_forward_was_graph = False
            self._wait_output_readers(dev)
            if one_call and dev.type == "cuda":
                canvas, occ = self.pp_layer([t for p in pcls for t in p], **kw)
                enc = [(canvas[f * B:(f + 1) * B], occ[f * B:(f + 1) * B]) for f in range(len(pcls))]
            else:
                enc = [self.pp_layer(p, **kw) for p in pcls]
            outs = self._net_body_frames([e[0] for e in enc], [e[1] for e in enc], pairs)
            return outs, [e[1] for e in enc]

        # ---- everything between the pillar encoder and the decoder as ONE CUDA graph: the encoder writes
        # its canvases straight into the graph's static inputs; ~800 launches per step become one graph launch.
        # The captured kernels hold raw pointers to the weights (and to cached concatenations of them): any in-place
        # update or re-allocation of a parameter invalidates the graph.
        # (the module tree is walked once; the modules' own parameter dictionaries are read every call, so a re-assigned
        # or re-allocated parameter is still seen)
        pdicts = self.__dict__.get("_graph_param_dicts")
        if pdicts is None:
            pdicts = self.__dict__["_graph_param_dicts"] = [m._parameters for root in (self.fnet, self.cnet, self.update_block)
                                                            for m in root.modules() if m._parameters]
        wsig = tuple((p.data_ptr(), p._version) for d in pdicts for p in d.values() if p is not None)
        key = (B, self.output_iterations, self.pp_layer.canvas_memory_format, str(dev), torch.backends.cudnn.allow_tf32, self.training,
               self.fused_update_block, self.merge_parallel_convs, self.tap_heads, self.concurrent_directions, FAST_STOCK_OPS,
               self.fuse_lookup_conv, self.stacked_stems, self.batched_frame_encoding, self.batched_encoders, self.direction_branches,
               self.output_sink is not None, self.graph_extra_key, wsig)
        slot = "net" if (len(pcls), pairs) == (2, [(0, 1)]) else "net:%d:%s" % (len(pcls), pairs)
        self._last_graph_slot = slot
        st = self._graphs.get(slot)
        if st is not None and st["key"] != key:
            st = None  # (the old graph and its buffers are released when the slot is overwritten)
        if st is None:
            big = self.pp_layer.empty_outputs(B * len(pcls), dev)  # the frames' canvases are batch slices of one buffer
            st = {"key": key, "in_all": big, "in": [(big[0][f * B:(f + 1) * B], big[1][f * B:(f + 1) * B]) for f in range(len(pcls))],
                  "pairs": pairs, "slot": slot}
        if one_call:
            self.pp_layer([t for p in pcls for t in p], out=st["in_all"], **kw)
        else:
            for f, p in enumerate(pcls):
                self.pp_layer(p, out=st["in"][f], **kw)
        self._wait_output_readers(dev)
        if "graph" not in st:
            self._capture_net_graph(st, dev)
        self.last_forward_was_graph = True
        st["graph"].replay()
        _lib_mod().note_graph_replay(st["launches"])
        return st["outs"], [i[1] for i in st["in"]]

    def _wait_output_readers(self, dev):
        """A consumer on another stream may still be reading the static outputs of the previous graph replay
        (``ExportPipeline``'s encoder sets ``pre_replay_event``): the pillar encoder in front did not have to wait for it,
        everything that may overwrite or release those buffers does."""
        if self.pre_replay_event is not None:
            if dev.type == "cuda":
                torch.cuda.current_stream(dev).wait_event(self.pre_replay_event)
            self.pre_replay_event = None

    def will_use_graph(self, pcl_t0, pcl_t1, *more) -> bool:
        dev = pcl_t0[0].device
        return (self.use_cuda_graph and FAST_STOCK_OPS and dev.type == "cuda" and not torch.is_grad_enabled()
                and not (self.training and self._graphed_part_depends_on_training_mode())
                and not torch.cuda.is_current_stream_capturing()
                and all(len(p) == len(pcl_t0) for p in (pcl_t1,) + more) and hasattr(self.pp_layer, "empty_outputs"))

    def _graphed_part_depends_on_training_mode(self) -> bool:
        """The reference's flow export never calls ``model.eval()`` (``experiment.py:164-198,225-361``): it runs in train
        mode under ``no_grad``.  The only train-mode-dependent layer of the released configuration is the pillar encoder's
        BatchNorm1d (batch statistics + running-stat update) -- and the pillar encoder runs in front of the graph.  The
        graph itself (encoders, pyramids, GRU loops) is mode-independent unless it holds an active Dropout or a norm with
        running statistics."""
        for mod in (self.fnet, self.cnet, self.update_block):
            for sub in mod.modules():
                if isinstance(sub, (nn.Dropout, nn.Dropout2d, nn.Dropout3d)) and sub.p > 0:
                    return True
                if isinstance(sub, nn.modules.batchnorm._NormBase) and sub.track_running_stats:
                    return True
        return False

    def _net_body(self, img_t0, img_t1, occupancies=(None, None)):
        """Feature encoders, correlation pyramids, context encoders and both refinement loops (raft_mod.py:82-257)."""
        outs = self._net_body_frames([img_t0, img_t1], list(occupancies), [(0, 1)])
        return outs[0], outs[1]

    def _net_body_frames(self, imgs, occupancies, pairs):
        """`_net_body` for several pairs over common frames: fnet and cnet run once per frame; direction 2p = a -> b of
        pair p, 2p + 1 = b -> a."""
        self._occupancies = [occupancies[f] for a, b in pairs for f in (a, b)]  # per direction: the SOURCE frame's occupancy
        if self.output_sink_begin is not None:
            self.output_sink_begin()
        sources = sorted({f for a, b in pairs for f in (a, b)})
        fmaps, ctx = [None] * len(imgs), {}
        joint = self._joint_frames(imgs) if len(sources) == len(imgs) else None
        if joint is not None:
            # every frame needs both encoders and the canvases are batch slices of one buffer: ONE pass of each encoder
            # over all frames (InstanceNorm is per sample, so only cuDNN's batch-size-dependent summation order differs
            # from per-frame calls: ~1e-6) -- half (a third) of the launches, fuller waves for the 80 x 80 layers
            B = imgs[0].shape[0]
            stems = self._stems(joint)
            fmap_all = self.fnet(joint, stem_out=stems[0]) if stems is not None else self.fnet(joint)
            net_all, inp_all = self._context(joint, stem_out=stems[1] if stems is not None else None)
            for f in range(len(imgs)):
                fmaps[f] = fmap_all[f * B:(f + 1) * B]
                ctx[f] = (net_all[f * B:(f + 1) * B], inp_all[f * B:(f + 1) * B])
        else:
            for f, img in enumerate(imgs):
                stems = self._stems(img) if f in sources else None  # a frame that needs both encoders: one stacked stem convolution
                fmaps[f] = self.fnet(img, stem_out=stems[0]) if stems is not None else self.fnet(img)
                if f in sources:  # context encoder of every frame a direction starts from (raft_mod.py:170-173)
                    ctx[f] = self._context(img, stem_out=stems[1] if stems is not None else None)
        dirs = [(a, b) for a, b in pairs for a, b in ((a, b), (b, a))]
        outs = [None] * len(dirs)

        def run(k):
            a, b = dirs[k]
            outs[k] = self.predict_single_flow_map_and_classes(imgs[a], fmaps[a], fmaps[b], direction=k, context=ctx[a])

        if self.concurrent_directions and imgs[0].is_cuda and torch.cuda.is_current_stream_capturing():
            # the forward and the backward directions (pyramid, refinement loop) share nothing but the read-only feature
            # maps, context tensors and weights: inside the CUDA graph they are two parallel branches, so the many short
            # kernels of one loop fill the gaps of the other.  Fork / join with stream waits (captured as graph
            # dependencies); every tensor a branch allocates stays on its own stream, the shared inputs outlive the join.
            # A pair has two directions = two branches; the six directions of a triple are dealt over `direction_branches`
            # branches (branch j runs directions j, j + n, ...; branch 0 is the capturing stream itself).
            main = torch.cuda.current_stream(imgs[0].device)
            n_br = max(1, min(len(dirs), int(self.direction_branches)))
            sides = [self._branch_stream(imgs[0].device, "bw" if j == 1 else "br%d" % j) for j in range(1, n_br)]
            for j, side in enumerate(sides, start=1):
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    for k in range(j, len(dirs), n_br):
                        run(k)
            for k in range(0, len(dirs), n_br):
                run(k)
            for side in sides:
                main.wait_stream(side)
            return outs
        for k in range(len(dirs)):
            run(k)
        return outs

    def _joint_frames(self, imgs):
        """The frames' canvases as ONE batch when they are consecutive batch slices of one buffer (how `forward_frames`
        lays them out), else None."""
        if not (self.batched_encoders and FAST_STOCK_OPS and len(imgs) > 1 and imgs[0].is_cuda and not torch.is_grad_enabled()
                and not self.fnet.training and not self.cnet.training):
            return None
        first = imgs[0]
        B = first.shape[0]
        step = B * first.stride(0) * first.element_size()
        for f, t in enumerate(imgs):
            if (t.shape != first.shape or t.stride() != first.stride() or t.dtype != first.dtype
                    or t.data_ptr() != first.data_ptr() + f * step or t.untyped_storage().data_ptr() != first.untyped_storage().data_ptr()):
                return None
        return first.as_strided((B * len(imgs),) + tuple(first.shape[1:]), first.stride(), first.storage_offset())

    def _stems(self, img):
        """The 7x7 / 2 stems of the feature and the context encoder on the same canvas (``extractor.py:262-266``, called from
        ``raft_mod.py:140-173``) as ONE stock convolution with the output channels stacked [fnet | cnet]: both read the
        105 MB canvas, and 32 output channels fill only half of the tensor-core tile cuDNN picks -- the stacked convolution
        costs about what one of them does.  The feature half goes through the sliced InstanceNorm + ReLU kernel (its bias
        is a per-channel constant the normalisation removes), the context half through bias + ReLU.  Returns (stem output
        for fnet, stem output for cnet) or None when the layout / module types do not allow it."""
        fn, cn = self.fnet, self.cnet
        if not (self.stacked_stems and FAST_STOCK_OPS and img.is_cuda and img.dtype == torch.float32 and not torch.is_grad_enabled()
                and not torch.is_autocast_enabled() and isinstance(fn, SmallEncoder) and isinstance(cn, SmallEncoder)
                and _same_geometry(fn.conv1, cn.conv1) and fn.conv1.in_channels == cn.conv1.in_channels
                and _is_identity(cn.norm1) and cn.conv1.bias is not None and fn.conv1.out_channels % 4 == 0
                and cn.conv1.out_channels % 4 == 0 and img.is_contiguous(memory_format=torch.channels_last)
                and _fused_norm_ok(fn.norm1, img, fn.conv1.out_channels)):
            return None
        c = fn.conv1
        w = _cat_params(self, "stem.weight", (fn.conv1.weight, cn.conv1.weight))
        raw = F.conv2d(img, w, None, c.stride, c.padding, c.dilation, c.groups)
        if not raw.is_contiguous(memory_format=torch.channels_last):
            raw = raw.contiguous(memory_format=torch.channels_last)
        nf, nc = fn.conv1.out_channels, cn.conv1.out_channels
        xf = instance_norm_nhwc(fn.norm1, raw, relu=True, channel_slice=(0, nf))
        xc = _glue().bias_relu_slice(raw, nf, nc, cn.conv1.bias)
        return xf, xc

    def _context(self, img, stem_out=None):
        """(tanh(net), relu(inp)) of the context encoder (``raft_mod.py:170-173``).  On the channels-last inference path the
        bias of cnet's output convolution, the split, tanh and relu are ONE glue launch (``slimb200_ctx_split``) instead of
        ATen's bias pass + tanh + clamp."""
        cn = self.cnet
        if (FAST_STOCK_OPS and img.is_cuda and img.dtype == torch.float32 and not torch.is_grad_enabled()
                and not torch.is_autocast_enabled() and isinstance(cn, SmallEncoder) and cn.conv2.bias is not None
                and not (cn.training and cn.dropout is not None) and self.hidden_dim % 4 == 0 and self.context_dim % 4 == 0):
            raw = cn(img, head_bias=False, stem_out=stem_out)
            if raw.dtype == torch.float32 and raw.is_contiguous(memory_format=torch.channels_last):
                return _glue().ctx_split(raw, cn.conv2.bias, self.hidden_dim, self.context_dim)
            raw = raw + cn.conv2.bias[None, :, None, None]
        else:
            raw = cn(img, stem_out=stem_out)
        net, inp = torch.split(raw, [self.hidden_dim, self.context_dim], dim=1)
        return torch.tanh(net), torch.relu(inp)

    def _branch_stream(self, device, name="bw", priority=-1):
        """Streams of the graph's branches.  The two refinement loops are captured on high-priority streams, the sink
        (decoder) branches on default-priority ones: the node priorities are captured with the graph, so the decoders'
        large grids fill the SMs only where the latency-bound loops leave them idle."""
        st = self._streams.get((str(device), name))
        if st is None:
            st = self._streams[(str(device), name)] = torch.cuda.Stream(device=device, priority=priority)
        return st

    def _emit(self, direction, it, out):
        """Hand a finished network output to the sink: inline in eager mode, on a forked stream (= a side branch of the
        graph, joined by `_join_sink` at the end of the loop) while capturing."""
        if self.output_sink is None:
            return
        occ = self._occupancies[direction]
        if out.is_cuda and torch.cuda.is_current_stream_capturing():
            cur = torch.cuda.current_stream(out.device)
            ds = self._branch_stream(out.device, "sink%d" % direction, priority=0)
            ds.wait_stream(cur)
            with torch.cuda.stream(ds):
                self.output_sink(direction, it, out, occ)
        else:
            self.output_sink(direction, it, out, occ)

    def _join_sink(self, direction, device):
        if self.output_sink is not None and device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            torch.cuda.current_stream(device).wait_stream(self._branch_stream(device, "sink%d" % direction, priority=0))

    def _capture_net_graph(self, st, dev):
        lib = _lib_mod()
        imgs, occ = [i[0] for i in st["in"]], [i[1] for i in st["in"]]
        # warm-up on a side stream (cuDNN autotuning, lazy kernel attributes), as torch.cuda.graphs asks
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._net_body_frames(imgs, occ, st["pairs"])
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        n0 = lib.load().slimb200_launch_count(-1)
        with torch.cuda.graph(graph, stream=self._branch_stream(dev, "fw")):
            st["outs"] = self._net_body_frames(imgs, occ, st["pairs"])
        st["launches"] = int(lib.load().slimb200_launch_count(-1) - n0)  # library kernels inside one replay
        st["graph"] = graph
        st["keepalive"] = _derived_values(self) + [getattr(self, "_packed_c1", None)]  # derived weights the captured kernels point at
        self._graphs[st["slot"]] = st
        self.n_graph_captures = getattr(self, "n_graph_captures", 0) + 1

    def _gru_loop(self, correlation, net, inp, img_hw, batch, device, direction=0) -> List[torch.Tensor]:
        """The refinement loop of ``raft_mod.py:188-257``: lookup -> update block -> coordinate / logit update, and the
        per-iteration network output.  Pure function of (pyramid, net, inp): this is what gets captured in a CUDA graph."""
        m = self.slim_cfg.model
        ds = m.feature_downsampling_factor
        h, w = img_hw[0] // ds, img_hw[1] // ds
        if self._fused_update_ok(correlation, net, inp):
            return self._gru_loop_fused(correlation, net, inp, h, w, batch, device, direction)
        coords0 = coords_grid(batch, h, w, device)
        coords1 = coords_grid(batch, h, w, device)
        logits = torch.zeros((batch, 4, h, w), dtype=torch.float32, device=device)
        outs = []
        for it in range(m.num_iters):
            coords1 = coords1.detach()
            logits = logits.detach()
            corr = correlation(coords1)
            net, dflow, dlogits, _ = self.update_block(net, inp, corr, coords1 - coords0, logits, None)
            coords1 = coords1 + dflow
            logits = logits + dlogits
            if self.output_iterations == "last" and it != m.num_iters - 1:
                continue
            if FAST_STOCK_OPS and coords1.is_cuda:
                outs.append(raft_output_fused(coords1 - coords0, logits, ds, self.bev_rows_res_meters_per_fs_pixel,
                                              self.bev_cols_res_meters_per_fs_pixel))
                self._emit(direction, it, outs[-1])
                continue
            # RAFT (x, y) pixel flow -> (row, col) metres (raft_mod.py:262-266)
            res = torch.tensor([self.bev_rows_res_meters_per_fs_pixel, self.bev_cols_res_meters_per_fs_pixel],
                               device=device, dtype=torch.float32)[None, :, None, None]
            flow_m = torch.flip(upflow_n(coords1 - coords0, n=ds), dims=[1]) * res
            outs.append(concat2network_output(uplogits_n(logits, n=ds), flow_m, flow_m))
            self._emit(direction, it, outs[-1])
        self._join_sink(direction, device)
        return outs

    def _fused_update_ok(self, correlation, net, inp) -> bool:
        """The glue-kernel version of the loop needs the channels-last fp32 inference setting of the export."""
        return (FAST_STOCK_OPS and self.fused_update_block and not torch.is_grad_enabled() and not torch.is_autocast_enabled()
                and net.is_cuda and net.dtype == torch.float32 and inp.dtype == torch.float32
                and getattr(correlation, "channels_last", False)
                and isinstance(self.update_block, SmallUpdateBlock) and self.update_block.gru.convz.padding_mode == "zeros"
                and net.shape[1] % 4 == 0 and inp.shape[1] % 4 == 0)

    def _gru_loop_fused(self, correlation, net, inp, h, w, batch, device, direction=0) -> List[torch.Tensor]:
        """Same loop as `_gru_loop` (raft_mod.py:188-257, update.py:23-38,70-93,130-150) with the element-wise work
        between the stock convolutions in the library's glue kernels: the two 304-channel GRU inputs
        [h | inp | motion] and [r*h | inp | motion] are persistent channels-last buffers whose slots the producers
        write directly (no torch.cat), the gates are two launches, and the coordinate / logit update is one."""
        g = _glue()
        m = self.slim_cfg.model
        ds = m.feature_downsampling_factor
        ub = self.update_block
        me, gru = ub.motion_encoder, ub.gru
        Ch, Cx = net.shape[1], inp.shape[1]
        n_in = gru.convz.in_channels
        hx = torch.empty((batch, n_in, h, w), dtype=torch.float32, device=device, memory_format=torch.channels_last)
        rhx = torch.empty_like(hx)
        # [h | inp] into both GRU input buffers with one launch (the first Ch channels of rhx are r * h, rewritten by the gate
        # kernel before the q convolution reads them: what lands there now does not matter)
        g.nhwc_pack_into([net, inp], [(hx, 0), (rhx, 0)])
        w_zr = _cat_params(gru, "zr.weight", (gru.convz.weight, gru.convr.weight))
        b_zr = _cat_params(gru, "zr.bias", (gru.convz.bias, gru.convr.bias))
        fh, lh = ub.static_flow_head, ub.classification_head
        coords1 = coords_grid(batch, h, w, device).contiguous()
        flow = torch.zeros((batch, 2, h, w), dtype=torch.float32, device=device)
        logits = torch.zeros((batch, lh.conv2.out_channels, h, w), dtype=torch.float32, device=device)

        def raw(conv, x):  # the convolution without its bias (the consumer adds it)
            return F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)

        # parallel convolutions evaluated as one launch each (stock cuDNN, like the stacked update | reset gates):
        # [conv_flow1 | conv_class1] on [flow | logits], [conv_flow2 | conv_class2] block-diagonal, both heads' conv1 on
        # the shared hidden state, both heads' conv2 block-diagonal
        merge = (self.merge_parallel_convs and _same_geometry(me.conv_flow1, me.conv_class1)
                 and _same_geometry(me.conv_flow2, me.conv_class2) and _same_geometry(fh.conv1, lh.conv1)
                 and _same_geometry(fh.conv2, lh.conv2) and me.conv_flow2.out_channels % 4 == 0
                 and me.conv_class2.out_channels % 4 == 0)
        if merge:
            # ([flow | logits] is padded from 6 to 8 channels: a tensor-core friendly width whatever cuDNN's autotuner sees)
            n_st = (2 + logits.shape[1] + 7) // 8 * 8
            w_m1, b_m1 = _stacked_params(me.conv_flow1, me.conv_class1, shared_input=False, pad_in_to=n_st)
            w_m2, b_m2 = _stacked_params(me.conv_flow2, me.conv_class2, shared_input=False)
            w_h1, b_h1 = _stacked_params(fh.conv1, lh.conv1, shared_input=True)
            w_h2, _ = _stacked_params(fh.conv2, lh.conv2, shared_input=False)
            # [flow | logits | 0], channels-last like the stock convolution that reads it (no layout copy per iteration)
            stacked = torch.zeros((batch, n_st, h, w), dtype=torch.float32, device=device).contiguous(memory_format=torch.channels_last)
            n_f = me.conv_flow2.out_channels
            # the k x k output convolution of the heads (6 channels: as slow in cuDNN as one to 256) as a 1x1 convolution
            # to k*k taps x 6 channels; the window sum of the taps is part of the update kernel
            k = fh.conv2.kernel_size[0]
            w_taps = None
            if (self.tap_heads and fh.conv2.kernel_size == (k, k) and fh.conv2.stride == (1, 1) and fh.conv2.dilation == (1, 1)
                    and fh.conv2.padding == (k // 2, k // 2) and k % 2 == 1 and k <= 7):
                # W1[(ky*k + kx)*6 + c][cin] = W[c][cin][ky][kx]
                w_taps = _derived(fh.conv2, ("taps",), (fh.conv2.weight, lh.conv2.weight),
                                  lambda: w_h2.permute(2, 3, 0, 1).reshape(k * k * w_h2.shape[0], w_h2.shape[1], 1, 1).contiguous())
        c1 = me.conv_stat_corr1
        fuse_lookup = (bool(self.fuse_lookup_conv) and (self.fuse_lookup_conv == "always" or torch.backends.cudnn.allow_tf32)
                       and c1.kernel_size == (1, 1) and c1.stride == (1, 1) and c1.padding == (0, 0) and c1.groups == 1
                       and c1.bias is not None and correlation.lookup_conv_supported(c1.out_channels))
        if fuse_lookup:  # tf32 weights in the kernel's shared-memory layout: packed once per weight version, kept on the module
            from .corr import PackedLookupConv

            packed_c1 = getattr(self, "_packed_c1", None)
            if packed_c1 is None or not packed_c1.matches(c1.weight, c1.bias):
                packed_c1 = self._packed_c1 = PackedLookupConv(c1.weight, c1.bias, m.corr_cfg.num_levels, m.corr_cfg.search_radius)
        outs = []
        for it in range(m.num_iters):
            if fuse_lookup:
                c = correlation.lookup_conv(coords1, packed_c1, relu=True)
            else:
                c = conv_relu(c1, correlation(coords1))
            if merge:
                flg = conv_relu(me.conv_flow2, conv_relu(me.conv_flow1, stacked, w_m1, b_m1), w_m2, b_m2)  # [flow | logits] features
                f, lg = flg[:, :n_f], flg[:, n_f:]
                out = conv_relu(me.conv, g.nhwc_cat([c, flg]))
            else:
                f = conv_relu(me.conv_flow2, conv_relu(me.conv_flow1, flow))
                lg = conv_relu(me.conv_class2, conv_relu(me.conv_class1, logits))
                out = conv_relu(me.conv, g.nhwc_cat([c, f, lg]))
            g.nhwc_pack_into([out, lg, f], [(hx, Ch + Cx), (rhx, Ch + Cx)])  # x = [inp | out | logits | flow]
            z = g.gru_gate_zr(F.conv2d(hx, w_zr, None, gru.convz.stride, gru.convz.padding), b_zr, hx, rhx, Ch)
            net = g.gru_gate_out(raw(gru.convq, rhx), gru.convq.bias, z, hx, Ch)
            if merge and w_taps is not None:
                g.iter_update_taps(F.conv2d(conv_relu(fh.conv1, net, w_h1, b_h1), w_taps), k, fh.conv2.bias, lh.conv2.bias,
                                   coords1, flow, logits, stacked)
            elif merge:
                d = F.conv2d(conv_relu(fh.conv1, net, w_h1, b_h1), w_h2, None, fh.conv2.stride, fh.conv2.padding)
                g.iter_update(d[:, :2], fh.conv2.bias, d[:, 2:], lh.conv2.bias, coords1, flow, logits, stacked)
            else:
                g.iter_update(raw(fh.conv2, conv_relu(fh.conv1, net)), fh.conv2.bias, raw(lh.conv2, conv_relu(lh.conv1, net)),
                              lh.conv2.bias, coords1, flow, logits)
            if self.output_iterations == "last" and it != m.num_iters - 1:
                continue
            outs.append(raft_output_fused(flow, logits, ds, self.bev_rows_res_meters_per_fs_pixel,
                                          self.bev_cols_res_meters_per_fs_pixel))
            self._emit(direction, it, outs[-1])
        self._join_sink(direction, device)
        return outs

    def predict_single_flow_map_and_classes(self, img_t0, fmap_t0, fmap_t1, decoder=None, direction=0, context=None) -> List[torch.Tensor]:
        """``context``: (tanh(net), relu(inp)) of ``cnet(img_t0)`` when the caller has already run the context encoder."""
        m = self.slim_cfg.model
        b, _, H, W = img_t0.shape
        correlation = CorrBlock(fmap_t0, fmap_t1, num_levels=m.corr_cfg.num_levels, radius=m.corr_cfg.search_radius)
        if context is None:
            context = self._context(img_t0)
        return self._gru_loop(correlation, context[0], context[1], (H, W), b, img_t0.device, direction)
