"""SURVEY 8(f).2 glue kernels (``csrc/gru_glue.cu``, the residual mode of ``slimb200_instnorm_nhwc``) against the stock
PyTorch ops they replace (``liso/slim/model/update.py:23-38,70-93,130-150``, ``raft_mod.py:188-212``,
``extractor.py:57-68``).  The kernels round like the PyTorch expressions, so the comparisons are (nearly) exact."""
import pytest
import torch
import torch.nn.functional as F

from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.slim import glue as G
from liso_b200.slim import raft as R
from liso_b200.slim.corr import coords_grid
from liso_b200.slim.slim import SLIM
from liso_b200.synth import make_sample_dicts
from liso_b200.weights import synth_weights_like

pytestmark = pytest.mark.gpu


def _nhwc(B, C, h, w, seed, device):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, C, h, w, generator=g).to(device).contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("B,h,w", [(1, 5, 7), (2, 16, 24), (8, 80, 80)])
def test_nhwc_pack_equals_cat(cuda, B, h, w):
    c, f, lg, o = (_nhwc(B, C, h, w, 10 + C, cuda) for C in (96, 32, 36, 80))
    got = G.nhwc_cat([c, f, lg])
    assert got.is_contiguous(memory_format=torch.channels_last) and torch.equal(got, torch.cat([c, f, lg], dim=1))
    assert torch.equal(G.nhwc_cat([c]), c) and torch.equal(G.nhwc_cat([c, f, lg, o]), torch.cat([c, f, lg, o], dim=1))
    # slot writes into two wider buffers; everything outside the slot stays untouched
    hx = _nhwc(B, 304, h, w, 1, cuda)
    rhx = _nhwc(B, 312, h, w, 2, cuda)
    hx0, rhx0 = hx.clone(), rhx.clone()
    G.nhwc_pack_into([o, lg, f], [(hx, 156), (rhx, 160)])
    cat = torch.cat([o, lg, f], dim=1)
    assert torch.equal(hx[:, 156:], cat) and torch.equal(hx[:, :156], hx0[:, :156])
    assert torch.equal(rhx[:, 160:308], cat) and torch.equal(rhx[:, :160], rhx0[:, :160]) and torch.equal(rhx[:, 308:], rhx0[:, 308:])
    # channel slices of a wider channels-last tensor as sources (the stacked flow | logits branch of the motion encoder)
    wide = _nhwc(B, 64, h, w, 3, cuda)
    G.nhwc_pack_into([o, wide[:, 32:], wide[:, :32]], [(hx, 160)])
    assert torch.equal(hx[:, 160:], torch.cat([o, wide[:, 32:], wide[:, :32]], dim=1)) and torch.equal(hx[:, :156], hx0[:, :156])
    with pytest.raises(RuntimeError):
        G.nhwc_pack_into([o, lg, f], [(hx, 160)])  # does not fit
    with pytest.raises(RuntimeError):
        G.nhwc_cat([c.cpu()])


@pytest.mark.parametrize("B,h,w", [(1, 5, 7), (8, 80, 80)])
def test_gru_gates_equal_stock_expressions(cuda, B, h, w):
    Ch = 96
    hx = _nhwc(B, 304, h, w, 3, cuda)
    rhx = _nhwc(B, 304, h, w, 4, cuda)
    zr = _nhwc(B, 2 * Ch, h, w, 5, cuda) * 3.0
    q = _nhwc(B, Ch, h, w, 6, cuda) * 2.0
    bzr = torch.randn(2 * Ch, device=cuda)
    bq = torch.randn(Ch, device=cuda)
    hprev = hx[:, :Ch].clone()
    x_part = rhx[:, Ch:].clone()
    # update.py:33-35 with the bias added the way F.conv2d does (a separate fp32 add)
    zr_ref = torch.sigmoid(zr + bzr[None, :, None, None])
    z_ref, r_ref = zr_ref[:, :Ch], zr_ref[:, Ch:]
    z = G.gru_gate_zr(zr, bzr, hx, rhx, Ch)
    assert torch.allclose(z, z_ref, rtol=0, atol=2e-7)
    assert torch.allclose(rhx[:, :Ch], r_ref * hprev, rtol=1e-6, atol=1e-7)
    assert torch.equal(rhx[:, Ch:], x_part) and torch.equal(hx[:, :Ch], hprev)
    h_ref = (1 - z) * hprev + z * torch.tanh(q + bq[None, :, None, None])
    hnew = G.gru_gate_out(q, bq, z, hx, Ch)
    assert hnew.is_contiguous(memory_format=torch.channels_last)
    assert torch.allclose(hnew, h_ref, rtol=1e-6, atol=3e-7)
    assert torch.equal(hx[:, :Ch], hnew)


@pytest.mark.parametrize("fmt", ["contiguous", "channels_last"])
@pytest.mark.parametrize("B,h,w", [(1, 5, 7), (8, 80, 80)])
def test_iter_update_equals_stock_ops(cuda, B, h, w, fmt):
    g = torch.Generator().manual_seed(7)
    mf = torch.channels_last if fmt == "channels_last" else torch.contiguous_format
    df = torch.randn(B, 2, h, w, generator=g).to(cuda).contiguous(memory_format=mf)
    dl = torch.randn(B, 4, h, w, generator=g).to(cuda).contiguous(memory_format=mf)
    bf, bl = torch.randn(2, device=cuda), torch.randn(4, device=cuda)
    coords0 = coords_grid(B, h, w, cuda)
    coords1 = (coords0 + torch.randn(B, 2, h, w, generator=g).to(cuda)).contiguous()
    logits = torch.randn(B, 4, h, w, generator=g).to(cuda)
    flow = torch.full((B, 2, h, w), float("nan"), device=cuda)
    c_ref = coords1 + (df + bf[None, :, None, None])
    l_ref = logits + (dl + bl[None, :, None, None])
    G.iter_update(df, bf, dl, bl, coords1, flow, logits)
    assert torch.equal(coords1, c_ref) and torch.equal(logits, l_ref) and torch.equal(flow, c_ref - coords0)
    # both head outputs as channel slices of ONE stacked tensor, plus the stacked [flow | logits] copy
    d6 = torch.cat([df, dl], dim=1).contiguous(memory_format=mf)
    stacked = torch.full((B, 8, h, w), float("nan"), device=cuda)  # 2 padding channels stay untouched
    c2_ref = c_ref + (df + bf[None, :, None, None])
    l2_ref = l_ref + (dl + bl[None, :, None, None])
    G.iter_update(d6[:, :2], bf, d6[:, 2:], bl, coords1, flow, logits, stacked)
    assert torch.equal(coords1, c2_ref) and torch.equal(logits, l2_ref) and torch.equal(flow, c2_ref - coords0)
    assert torch.equal(stacked[:, :6], torch.cat([flow, logits], dim=1)) and bool(torch.isnan(stacked[:, 6:]).all())
    # the heads' 3x3 output convolution given as 1x1 "taps": window sum inside the kernel == F.conv2d (summation order aside)
    x = torch.randn(B, 32, h, w, generator=g).to(cuda)
    w6 = (0.1 * torch.randn(6, 32, 3, 3, generator=g)).to(cuda)
    ref6 = F.conv2d(x, w6, None, padding=1)
    taps = F.conv2d(x, w6.permute(2, 3, 0, 1).reshape(54, 32, 1, 1).contiguous()).contiguous(memory_format=mf)
    c3_ref = coords1 + (ref6[:, :2] + bf[None, :, None, None])
    l3_ref = logits + (ref6[:, 2:] + bl[None, :, None, None])
    G.iter_update_taps(taps, 3, bf, bl, coords1, flow, logits, stacked)
    assert torch.allclose(coords1, c3_ref, rtol=0, atol=2e-5) and torch.allclose(logits, l3_ref, rtol=0, atol=2e-5)
    assert torch.equal(flow, coords1 - coords0) and torch.equal(stacked[:, :6], torch.cat([flow, logits], dim=1))
    # channels-last copy (what the stacked motion-encoder convolution reads)
    st_cl = torch.zeros_like(stacked).contiguous(memory_format=torch.channels_last)
    G.iter_update_taps(taps, 3, bf, bl, coords1.clone(), flow.clone(), logits.clone(), st_cl)
    G.iter_update_taps(taps, 3, bf, bl, coords1, flow, logits, stacked)
    assert torch.equal(st_cl[:, :6], stacked[:, :6]) and not st_cl.is_contiguous() and float(st_cl[:, 6:].abs().sum()) == 0.0


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 40, 56), (3, 96, 17, 23), (8, 32, 320, 320)])
def test_residual_joins(cuda, B, C, H, W):
    x, y = _nhwc(B, C, H, W, 8, cuda), _nhwc(B, C, H, W, 9, cuda)
    assert torch.equal(R.add_relu(x, y), F.relu(x + y))
    assert torch.equal(R.add_relu(x.contiguous(), y.contiguous()), F.relu(x + y).contiguous())
    from liso_b200.slim import glue as G

    bias = torch.randn(x.shape[1], device=x.device)
    xc, yc = x.contiguous(memory_format=torch.channels_last), y.contiguous(memory_format=torch.channels_last)
    assert torch.equal(G.add_relu(xc, yc, bias_x=bias), F.relu((xc + bias[None, :, None, None]) + yc))
    norm = torch.nn.InstanceNorm2d(C, eps=1e-3, affine=True).to(cuda)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
        for relu in (True, False):
            plain = R.instance_norm_nhwc(norm, y, relu)
            fused = R.instance_norm_nhwc(norm, y, relu, residual=x)
            assert fused.is_contiguous(memory_format=torch.channels_last)
            assert torch.equal(fused, F.relu(x + plain))


def _slim(cfg, device):
    m = SLIM(cfg).eval()
    m.load_state_dict(synth_weights_like(m.state_dict(), 0), strict=True)
    return m.to(device).to(memory_format=torch.channels_last)


def test_fused_update_block_equals_stock_loop(cuda):
    """Whole forward with the glue kernels (fused update block, fused residual joins) vs the same network on stock
    element-wise ops (metres / logits of the (B, H, W, 8) network output)."""
    cfg = make_cfg("T")
    model = _slim(cfg, cuda)
    s0, s1 = make_sample_dicts(WORKLOADS["T"], [41, 42])
    net = model.raft_network
    net.use_cuda_graph = False
    p0 = [p.to(cuda) for p in s0["pcl_full_no_ground_ta"]]
    p1 = [p.to(cuda) for p in s1["pcl_full_no_ground_ta"]]
    with torch.no_grad():
        fused = [o.clone() for o in net(p0, p1)[0]]
        net.tap_heads = False
        untapped = [o.clone() for o in net(p0, p1)[0]]
        net.merge_parallel_convs = False
        unmerged = [o.clone() for o in net(p0, p1)[0]]
        net.merge_parallel_convs = net.tap_heads = True
        net.fused_update_block = False
        stock_loop = [o.clone() for o in net(p0, p1)[0]]
        R.FAST_STOCK_OPS = False
        try:
            stock = [o.clone() for o in net(p0, p1)[0]]
        finally:
            R.FAST_STOCK_OPS = True
            net.fused_update_block = True
    assert len(fused) == len(stock_loop) == len(stock) == 6
    for a, t in zip(fused, untapped):  # 1x1 taps + window sum vs the 3x3 convolution: another summation order
        assert float((a - t).abs().max()) <= 2e-4, float((a - t).abs().max())
    for a, u, b, c in zip(fused, unmerged, stock_loop, stock):
        assert a.shape == b.shape == c.shape == u.shape
        # stacked parallel convolutions: same products, possibly another summation order inside cuDNN
        assert float((a - u).abs().max()) <= 2e-4, float((a - u).abs().max())
        # same convolutions, same element-wise rounding: the glue kernels reproduce the stock loop
        assert float((u - b).abs().max()) <= 1e-4, float((u - b).abs().max())
        assert float((a - c).abs().max()) <= 5e-3, float((a - c).abs().max())


def test_context_split(cuda):
    """slimb200_ctx_split == tanh / relu of the split of (raw + bias) (raft_mod.py:170-173), to the last bit."""
    from liso_b200.slim import glue as G

    raw = _nhwc(3, 160, 20, 28, 4, cuda)
    bias = torch.randn(160, device=cuda)
    net, inp = G.ctx_split(raw, bias, 96, 64)
    full = raw + bias[None, :, None, None]
    assert torch.equal(net, torch.tanh(full[:, :96])) and torch.equal(inp, torch.relu(full[:, 96:]))
    assert net.is_contiguous(memory_format=torch.channels_last) and inp.is_contiguous(memory_format=torch.channels_last)


def test_sliced_instance_norm_and_bias_relu(cuda):
    """The two halves of a stacked stem convolution: InstanceNorm + ReLU of a channel slice == the packed kernel on a copy of
    the slice (bit for bit), bias + ReLU of the other slice == the stock expression."""
    from liso_b200.slim import glue as G

    raw = _nhwc(3, 64, 40, 56, 11, cuda)
    norm = torch.nn.InstanceNorm2d(32, eps=1e-3, affine=True).to(cuda)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
        a = R.instance_norm_nhwc(norm, raw, relu=True, channel_slice=(0, 32))
        b = R.instance_norm_nhwc(norm, raw[:, :32].contiguous(memory_format=torch.channels_last), relu=True)
        ref = F.relu(norm(raw[:, :32]))
    assert torch.equal(a, b) and a.is_contiguous(memory_format=torch.channels_last)
    assert float((a - ref).abs().max()) < 1e-5
    bias = torch.randn(32, device=cuda)
    c = G.bias_relu_slice(raw, 32, 32, bias)
    assert torch.equal(c, F.relu(raw[:, 32:] + bias[None, :, None, None]))


def test_stacked_stems_equal_separate_encoders(cuda):
    """RAFT._stems: fnet / cnet stems as one stacked convolution -> the encoders' outputs equal the separate runs to fp32
    round-off (the stacked convolution may pick another cuDNN algorithm; fp32 convolutions here)."""
    from liso_b200.config import make_cfg
    from liso_b200.slim.slim import SLIM
    from liso_b200.weights import synth_weights_like

    torch.backends.cudnn.allow_tf32 = False
    cfg = make_cfg("T")
    model = SLIM(cfg).eval()
    model.load_state_dict(synth_weights_like(model.state_dict(), 0))
    net = model.to(cuda).to(memory_format=torch.channels_last).raft_network
    img = torch.randn(2, 64, 256, 256, device=cuda).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        stems = net._stems(img)
        assert stems is not None
        f1, (n1, i1) = net.fnet(img, stem_out=stems[0]), net._context(img, stem_out=stems[1])
        f0, (n0, i0) = net.fnet(img), net._context(img)
    for a, b in ((f1, f0), (n1, n0), (i1, i0)):
        assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(b.abs().max()))
