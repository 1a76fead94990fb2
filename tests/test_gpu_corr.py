"""Stage-2 parity on the B200: tcgen05 correlation pyramid + lookup kernel vs the CPU oracle.

Stated tolerances
* correlation / pyramid entries (bf16 operands, fp32 accumulate, bf16 storage):
    |got - fp64| <= 2^-7 * ||f1_i||_2 * ||f2_j||_2 / sqrt(D)          
* lookup on a *given* pyramid: <= 1e-5 relative to the pyramid's max magnitude (fp32 interpolation)
"""
import numpy as np
import pytest
import torch

from liso_b200.slim import corr as C
from oracle import slim_oracle as O

pytestmark = pytest.mark.gpu


def _fmaps(B, h, w, seed, device):
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(B, 128, h, w, generator=g)
    f2 = torch.randn(B, 128, h, w, generator=g)
    # spatial structure so pooled levels are not just noise
    f2 = f2 + 0.5 * torch.nn.functional.avg_pool2d(f2, 3, stride=1, padding=1)
    return f1, f2, f1.to(device), f2.to(device)


def _bound(f1, f2_level):
    n1 = f1.double().flatten(2).norm(dim=1)  # (B, Nf)
    n2 = f2_level.double().flatten(2).norm(dim=1)  # (B, Nl)
    return (2.0 ** -7) * n1[:, :, None] * n2[:, None, :] / np.sqrt(128.0)


@pytest.mark.parametrize("fmt", ["contiguous", "channels_last"])
@pytest.mark.parametrize("B,h,w", [(1, 16, 16), (2, 32, 32), (1, 80, 80), (1, 23, 23), (2, 12, 20), (1, 115, 115)])
def test_pyramid_within_bf16_bound(cuda, B, h, w, fmt):
    f1, f2, d1, d2 = _fmaps(B, h, w, 0, cuda)
    if fmt == "channels_last":  # what a channels-last fnet hands over: no transposition kernel on this path
        d1, d2 = d1.contiguous(memory_format=torch.channels_last), d2.contiguous(memory_format=torch.channels_last)
    blk = C.CorrBlock(d1, d2, num_levels=4, radius=3)
    torch.cuda.synchronize()
    ref = O.corr_pyramid(f1, f2, 4)
    f2_l = f2
    for l in range(4):
        got = blk.corr_pyramid[l].float().cpu()
        assert got.shape == ref[l].shape, (got.shape, ref[l].shape)
        exact = torch.einsum("bdi,bdj->bij", f1.double().flatten(2), f2_l.double().flatten(2)) / np.sqrt(128.0)
        err = (got.double().reshape(B, h * w, -1) - exact).abs()
        bound = _bound(f1, f2_l)
        assert bool((err <= bound).all()), (l, float(err.max()), float((err / bound).max()))
        # and close to the oracle's fp32 pyramid in aggregate (rms error ~ 2^-9 of the entry scale)
        rel_rms = float((got - ref[l]).pow(2).mean().sqrt() / ref[l].pow(2).mean().sqrt())
        assert rel_rms < 6e-3, (l, rel_rms)
        if l < 3:
            f2_l = torch.nn.functional.avg_pool2d(f2_l, 2, stride=2)


def test_corr_static_method_shape(cuda):
    f1, f2, d1, d2 = _fmaps(1, 16, 24, 3, cuda)
    vol = C.CorrBlock.corr(d1, d2)
    ref = O.corr_volume(f1, f2)
    assert vol.shape == ref.shape == (1, 16, 24, 1, 16, 24)
    assert float((vol.cpu() - ref).abs().max()) < 0.25


def _coords(B, h, w, seed, spread):
    g = torch.Generator().manual_seed(seed)
    c = O.coords_grid(B, h, w) + spread * torch.randn(B, 2, h, w, generator=g)
    # exact integers, negatives and far out-of-range positions
    c[:, :, 0, 0] = torch.tensor([3.0, 2.0])
    c[:, :, 0, 1] = torch.tensor([-2.5, 1.25])
    c[:, :, 1, 0] = torch.tensor([w + 9.0, h + 7.0])
    c[:, :, 1, 1] = torch.tensor([w - 1.0, h - 1.0])
    c[:, :, 2, 2] = torch.tensor([-40.0, -40.0])
    return c


@pytest.mark.parametrize("B,h,w", [(1, 16, 16), (2, 24, 40), (1, 80, 80), (1, 23, 29)])
def test_lookup_fp32_pyramid_matches_oracle(cuda, B, h, w):
    """T8: same pyramid in, lookup out: isolates the gather kernel (fp32 storage path)."""
    f1, f2, _, _ = _fmaps(B, h, w, 1, cuda)
    ref_pyr = O.corr_pyramid(f1, f2, 4)
    L = C.make_layout(B, 128, h, w, 4)
    packed = C.pack_pyramid_f32([p.to(cuda) for p in ref_pyr], L)
    for spread in (0.0, 0.7, 6.0):
        coords = _coords(B, h, w, 2, spread)
        got = C.lookup(packed, L, coords.to(cuda), 3).cpu()
        ref = O.corr_lookup(ref_pyr, coords, 3)
        assert got.shape == ref.shape == (B, 196, h, w)
        scale = float(ref_pyr[0].abs().max())
        assert float((got - ref).abs().max()) <= 1e-5 * scale, (spread, float((got - ref).abs().max()), scale)


@pytest.mark.parametrize("fmt", ["contiguous", "channels_last"])
@pytest.mark.parametrize("B,h,w", [(2, 32, 32), (1, 80, 80), (1, 115, 115)])
def test_lookup_on_bf16_pyramid(cuda, B, h, w, fmt):
    """End of stage 2: kernel lookup on the kernel's own bf16 pyramid == oracle lookup on the same values.
    With channels-last feature maps the lookup is answered channels-last too (same logical tensor)."""
    f1, f2, d1, d2 = _fmaps(B, h, w, 4, cuda)
    if fmt == "channels_last":
        d1, d2 = d1.contiguous(memory_format=torch.channels_last), d2.contiguous(memory_format=torch.channels_last)
    blk = C.CorrBlock(d1, d2, num_levels=4, radius=3)
    same_values = [lv.float().cpu().contiguous() for lv in blk.corr_pyramid]
    scale = float(same_values[0].abs().max())
    ref_pyr32 = O.corr_pyramid(f1, f2, 4)
    # spread 0: the pixel grid itself (what the first GRU iteration asks for) -- every position sits on an integer and
    # the per-offset floors scatter around it ("shifted" windows of the kernel); 0.5: half-integers at level 0
    for spread, seed in ((1.5, 5), (0.0, 6), (0.5, 7)):
        coords = _coords(B, h, w, seed, spread if spread != 0.5 else 0.0)
        if spread == 0.5:
            coords = coords + 0.5
        got_dev = blk(coords.to(cuda))
        assert got_dev.is_contiguous(memory_format=torch.channels_last if fmt == "channels_last" else torch.contiguous_format)
        got = got_dev.cpu()
        ref = O.corr_lookup(same_values, coords, 3)
        assert float((got - ref).abs().max()) <= 1e-5 * scale, (spread, float((got - ref).abs().max()), scale)
        # and against the full-precision oracle: bounded by the bf16 storage/operand rounding
        ref32 = O.corr_lookup(ref_pyr32, coords, 3)
        assert float((got - ref32).abs().max()) <= 2.0 ** -6 * scale


def test_window_channel_order(cuda):
    """Q6: channel k = l*49 + i*7 + j samples (x + i - 3, y + j - 3): first window axis offsets x."""
    B, h, w = 1, 16, 16
    L = C.make_layout(B, 128, h, w, 1)
    lvl = torch.zeros(h * w, 1, h, w)
    lvl[:, 0] = torch.arange(h * w, dtype=torch.float32).view(h, w)  # value = 16*row + col for every source pixel
    packed = C.pack_pyramid_f32([lvl.to(cuda)], L)
    coords = torch.zeros(1, 2, h, w)
    coords[0, 0] = 8.0  # x (column)
    coords[0, 1] = 5.0  # y (row)
    out = C.lookup(packed, L, coords.to(cuda), 3).cpu()
    for i in range(7):
        for j in range(7):
            # (the normalise / un-normalise round trip of grid_sample is not exact: ~1e-6 relative)
            assert abs(float(out[0, i * 7 + j, 0, 0]) - (16.0 * (5 + j - 3) + (8 + i - 3))) < 1e-3


# ----------------------------------------------------------------------------------------------------------------------
# second-generation gather (csrc/corr_lookup2.cu) and the lookup fused with the 1x1 convolution that consumes it
# ----------------------------------------------------------------------------------------------------------------------
def _wild_coords(B, h, w, seed, spread):
    c = _coords(B, h, w, seed, spread)
    c[:, :, 3, 3] = torch.tensor([1e9, -1e9])          # absurd but finite: every tap outside
    c[:, :, 3, 4] = torch.tensor([float("inf"), 2.0])  # non-finite: the kernel's fully predicated path
    c[:, :, 4, 3] = torch.tensor([0.0, h - 1.0])       # corners
    c[:, :, 4, 4] = torch.tensor([w - 1.0, 0.0])
    return c


@pytest.mark.parametrize("fmt", ["contiguous", "channels_last"])
@pytest.mark.parametrize("B,h,w", [(3, 16, 16), (1, 80, 80), (1, 115, 115), (2, 12, 20)])
def test_lookup_generations_agree(cuda, B, h, w, fmt):
    """The (pixel, level)-per-thread and the row-per-thread gathers answer like the first-generation kernel (same blend
    association: equal up to the sign of zero) on regular, integer ("shifted"), out-of-range and non-finite coordinates."""
    from liso_b200 import _lib

    f1, f2, d1, d2 = _fmaps(B, h, w, 11, cuda)
    if fmt == "channels_last":
        d1, d2 = d1.contiguous(memory_format=torch.channels_last), d2.contiguous(memory_format=torch.channels_last)
    blk = C.CorrBlock(d1, d2, num_levels=4, radius=3)
    lib = _lib.load()
    for spread in (1.5, 0.0, 6.0):
        coords = _wild_coords(B, h, w, 12, spread).to(cuda)
        default = lib.slimb200_lookup_generation(-1)
        assert default in (1, 2)
        try:
            lib.slimb200_lookup_generation(0)
            old = blk(coords).cpu()
            scale = float(old.abs().max())
            for generation in (1, 2):
                lib.slimb200_lookup_generation(generation)
                new = blk(coords)
                assert new.is_contiguous(memory_format=torch.channels_last if fmt == "channels_last" else torch.contiguous_format)
                new = new.cpu()
                assert torch.isfinite(new).all()
                assert float(new[:, :, 3, 3].abs().max()) == 0.0 and float(new[:, :, 3, 4].abs().max()) == 0.0
                assert float((new - old).abs().max()) <= 2e-6 * scale, (generation, spread, float((new - old).abs().max()), scale)
        finally:
            lib.slimb200_lookup_generation(default)


def _conv_ref(look, weight, bias, relu):
    """fp64 reference of act(conv1x1(look)) and the tf32 operand bound 2^-10 * sum_k |w_nk| |v_k| per output."""
    B, K, h, w = look.shape
    v = look.double().permute(0, 2, 3, 1).reshape(-1, K)
    y = v @ weight.double().t() + (bias.double() if bias is not None else 0.0)
    bound = (2.0 ** -10) * (v.abs() @ weight.double().abs().t())
    if relu:
        y = y.clamp_min(0.0)
    n = weight.shape[0]
    return y.reshape(B, h, w, n).permute(0, 3, 1, 2), bound.reshape(B, h, w, n).permute(0, 3, 1, 2)


@pytest.fixture(params=[3, 4])
def conv_generation(request):
    """Both fused kernels: 3 = row-per-thread gather, A tile in shared memory; 4 = cp.async landing slots, A tile in TMEM."""
    from liso_b200 import _lib

    lib = _lib.load()
    prev = lib.slimb200_lookup_conv_generation(request.param)
    yield request.param
    lib.slimb200_lookup_conv_generation(prev)


@pytest.mark.parametrize("B,h,w,n_out", [(3, 16, 16, 96), (1, 80, 80, 96), (1, 115, 115, 96), (2, 24, 40, 32), (2, 16, 20, 64),
                                         (1, 23, 29, 64), (5, 80, 80, 96)])
def test_lookup_conv_fused_matches_oracle(cuda, conv_generation, B, h, w, n_out):
    """slimb200_corr_lookup_conv == relu(conv_stat_corr1(CorrBlock(...)(coords))) of the oracle
    (corr.py:23-46 + update.py:49,71) within the tf32 operand bound: the window values and the weights are rounded to
    tf32 (rel. 2^-11 each), products accumulate in fp32:
        |got - fp64| <= 2^-10 * sum_k |w_nk| |v_k|  +  1e-5 * scale * sum_k |w_nk|  (the lookup's own fp32 interpolation)"""
    f1, f2, d1, d2 = _fmaps(B, h, w, 21, cuda)
    d1, d2 = d1.contiguous(memory_format=torch.channels_last), d2.contiguous(memory_format=torch.channels_last)
    blk = C.CorrBlock(d1, d2, num_levels=4, radius=3)
    assert blk.lookup_conv_supported(n_out)
    same_values = [lv.float().cpu().contiguous() for lv in blk.corr_pyramid]
    scale = float(same_values[0].abs().max())
    g = torch.Generator().manual_seed(22)
    weight = torch.randn(n_out, 196, 1, 1, generator=g) / 14.0
    bias = torch.randn(n_out, generator=g)
    wd, bd = weight.to(cuda), bias.to(cuda)
    for spread, relu in ((1.5, True), (0.0, True), (6.0, False)):
        coords = _wild_coords(B, h, w, 23, spread)
        got_dev = blk.lookup_conv(coords.to(cuda), wd, bd, relu=relu)
        assert got_dev.shape == (B, n_out, h, w) and got_dev.is_contiguous(memory_format=torch.channels_last)
        got = got_dev.cpu()
        look = O.corr_lookup(same_values, coords, 3)
        # non-finite coordinates: grid_sample answers NaN, the kernels define "every tap outside" (= 0, the old kernel too)
        assert not torch.isfinite(look[:, :, 3, 4]).all()
        look[:, :, 3, 4] = 0.0
        ref, bound = _conv_ref(look, weight.reshape(n_out, 196), bias, relu)
        tol = bound + 1e-5 * scale * float(weight.abs().sum(dim=1).max()) + 1e-6
        err = (got.double() - ref).abs()
        assert bool((err <= tol).all()), (spread, float(err.max()), float((err / tol).max()))
        # aggregate accuracy: rms error two orders below the bound's scale (tf32 rounding is unbiased)
        assert float(err.pow(2).mean().sqrt()) < 2.0 ** -11 * float(ref.abs().max()) + 1e-6
        # and against the library's own unfused path (lookup kernel -> fp32 conv)
        assert torch.isfinite(got).all()
        unf = torch.nn.functional.conv2d(blk(coords.to(cuda)), wd, bd)
        unf = torch.relu(unf) if relu else unf
        assert float((got_dev - unf).abs().max()) <= float(tol.max())


def test_lookup_conv_writes_channel_slice(cuda, conv_generation):
    """`out` may be a wider channels-last tensor: only its first C_out channels are written (no torch.cat afterwards)."""
    B, h, w, n_out = 2, 24, 40, 96
    _, _, d1, d2 = _fmaps(B, h, w, 31, cuda)
    blk = C.CorrBlock(d1.contiguous(memory_format=torch.channels_last), d2.contiguous(memory_format=torch.channels_last), 4, 3)
    g = torch.Generator().manual_seed(32)
    wd = (torch.randn(n_out, 196, generator=g) / 14.0).to(cuda)
    coords = _coords(B, h, w, 33, 1.5).to(cuda)
    ref = blk.lookup_conv(coords, wd, None, relu=False)
    wide = torch.full((B, 160, h, w), 7.0, device=cuda).contiguous(memory_format=torch.channels_last)
    ret = blk.lookup_conv(coords, wd, None, relu=False, out=wide)
    assert ret.data_ptr() == wide.data_ptr()
    assert torch.equal(wide[:, :n_out], ref) and bool((wide[:, n_out:] == 7.0).all())
    with pytest.raises(RuntimeError):
        blk.lookup_conv(coords, torch.zeros(48, 196, device=cuda), None)  # C_out must be a multiple of 32
    packed = C.PackedLookupConv(wd, None, 4, 3)  # what a loop keeps: packed once, re-packed when the parameter changes
    assert packed.matches(wd, None) and not packed.matches(wd.clone(), None)
    assert torch.equal(blk.lookup_conv(coords, packed, relu=False), ref)
    wd.mul_(2.0)
    assert not packed.matches(wd, None)
