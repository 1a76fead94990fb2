"""GPU .npz writer: the stream format (oracle restatement vs zlib), the zip / npy framing read back by
``np.load`` like ``torch_dataset_commons.py:614-616`` does, and -- on the GPU -- the kernel's bytes against the oracle."""
import io
import os
import zlib

import numpy as np
import pytest
import torch

from liso_b200.slim import npz_stream
from oracle import npz_deflate_oracle as DO


def _cases():
    rng = np.random.default_rng(7)
    out = {}
    a = np.zeros(3 * 2048 + 77, np.float32)  # ragged last chunk, runs across chunk borders
    a[[5, 6, 2047, 2048, 4100, 6200]] = rng.normal(size=6).astype(np.float32)
    out["sparse_ragged"] = a
    out["all_zero"] = np.zeros(2 * 2048, np.float32)
    out["one_word"] = np.array([1.5], np.float32)
    out["one_zero"] = np.zeros(1, np.float32)
    out["dense"] = rng.normal(size=5000).astype(np.float32)  # no zero at all, 9-bit literals included
    r = []  # every run length 1 .. 140 words (every remainder of the match rule), a different word between
    for n in range(1, 141):
        r += [0.0] * n + [float(n)]
    out["all_run_lengths"] = np.array(r, np.float32)
    r = []  # the same with runs of a NON-zero word (the denormal the softmax leaves in empty pillars)
    for n in range(1, 141):
        r += [3.8e-44] * n + [float(n)]
    out["all_run_lengths_denormal"] = np.array(r, np.float32)
    b = np.zeros(2048 * 2, np.float32)  # run bytes = 258 k + 2: the remainder that borrows from the last full match
    b[66] = 1.0  # 66 zero words: a run of 65 words = 260 bytes = 258 + 2
    b[67 + 195] = 2.0  # a run of 194 words = 776 bytes = 3 * 258 + 2
    b[67 + 196 + 130] = 3.0  # a run of 129 words = 516 bytes = 2 * 258 exactly
    out["remainders"] = b
    dyn = np.full((96, 96), 3.8e-44, np.float32)  # dynamicness-like: constant denormal, 6 % occupied
    m2 = rng.random((96, 96)) < 0.06
    dyn[m2] = rng.random(int(m2.sum())).astype(np.float32)
    out["bev_dynamicness"] = dyn
    bev = np.zeros((96, 96, 2), np.float32)  # BEV-like: 6 % occupied cells
    m = rng.random((96, 96)) < 0.06
    bev[m] = rng.normal(size=(int(m.sum()), 2)).astype(np.float32)
    out["bev_flow"] = bev
    neg = np.zeros(300, np.float32)
    neg[10] = -0.0  # negative zero is a non-zero WORD (0x80000000): must survive bit for bit
    out["negative_zero"] = neg
    return out


CASES = _cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_stream_is_valid_deflate(name):
    a = CASES[name]
    stream = DO.encode_member(a)
    assert zlib.decompress(stream, -15) == a.tobytes()
    d = zlib.decompressobj(-15)  # the stream ends exactly with its final block: nothing left over, nothing missing
    assert d.decompress(stream) == a.tobytes() and d.eof and d.unused_data == b""


def test_fixed_code_tables():
    """RFC 1951 3.2.6 / 3.2.5 spot values: literal 0 = 00110000, literal 144 = 110010000, length 3 = symbol 257 (0000001),
    length 258 = symbol 285 (11000101), length 11..12 share symbol 265 with one extra bit."""
    assert DO.lit_code(0) == (0b00001100, 8) and DO.lit_code(144) == (0b000010011, 9) and DO.lit_code(255) == (0b111111111, 9)
    assert DO.match_code(3) == (0b1000000, 12) and DO.match_code(258) == (0b10100011, 13)
    assert DO.match_code(11)[1] == 13 and DO.match_code(12)[0] == DO.match_code(11)[0] | (1 << 7)
    assert DO.match_code(257)[1] == 8 + 5 + 5
    assert DO.match_code(258, DO.DIST_4) == (0b10100011 | (0b11000 << 8), 13)  # distance symbol 3 = 00011, sent MSB first
    # the two codes the kernel carries as constants (csrc/npz_deflate.cu::run_tokens)
    assert DO.match_code(258, DO.DIST_4) == (0x18A3, 13) and DO.match_code(255, DO.DIST_4) == (0x31C23, 18)


def test_npz_framing_reads_back_with_np_load(tmp_path):
    """build_npz: stored-block .npy header + device stream + CRC completed from the remainder; np.load checks the CRC."""
    members = [("static_threshold", np.array(0.5, np.float32))]
    for name in ("bev_flow", "sparse_ragged", "dense"):
        a = CASES[name]
        members.append((name, a.shape, a.dtype, DO.encode_member(a), DO.crc_remainder(a)))
    members.append(("bev_range_m", np.array([70.0, 70.0])))
    blob = npz_stream.build_npz(members)
    p = tmp_path / "x.npz"
    p.write_bytes(blob)
    z = np.load(p)
    assert z.files == ["static_threshold", "bev_flow", "sparse_ragged", "dense", "bev_range_m"]
    for name in ("bev_flow", "sparse_ragged", "dense"):
        assert z[name].dtype == np.float32 and z[name].shape == CASES[name].shape
        assert np.array_equal(z[name].view(np.uint32), CASES[name].view(np.uint32))
    assert float(z["static_threshold"]) == 0.5 and z["bev_range_m"].tolist() == [70.0, 70.0]
    import zipfile

    assert zipfile.ZipFile(io.BytesIO(blob)).testzip() is None  # every member's CRC-32 verifies
    # the header np.save writes == the one framed here
    f = io.BytesIO()
    np.save(f, CASES["bev_flow"])
    assert f.getvalue().startswith(npz_stream.npy_header(CASES["bev_flow"].shape))
    # a wrong remainder must be caught by the reader (the CRC is live, not decorative)
    bad = npz_stream.build_npz([("a", (300,), np.float32, DO.encode_member(CASES["negative_zero"]), 1)])
    with pytest.raises(Exception):
        np.load(io.BytesIO(bad))["a"]


def _gf_mul(a, b):  # GF(2)[x] / P in zlib's reflected representation (bit 31 = x^0)
    p = 0
    for i in range(32):
        if a & (0x80000000 >> i):
            p ^= b
        b = (b >> 1) ^ (0xEDB88320 if b & 1 else 0)
    return p


def test_plan_geometric_factor_gives_the_crc_of_a_constant_fill():
    """Host-only part of the C ABI: slimb200_deflate_plan cuts members into chunks and fills crc_geo = sum_i x^(32 i), with
    which the kernel turns the remainder of ONE background word into the remainder of the whole constant fill."""
    import ctypes as C

    from liso_b200 import _lib

    lib = _lib.load()
    sizes = [1, 2, 3, 2048, 2049, 5000, 640 * 640, 640 * 640 * 2, 920 * 920 * 2]
    members = (_lib.DeflateMember * len(sizes))()
    for m, n in zip(members, sizes):
        m.src, m.words_per_cell, m.cell_stride, m.n_words = 4096, 1, 16, n  # (a fake, aligned device address: nothing is launched)
    total, ws, bound = C.c_int64(), C.c_size_t(), C.c_size_t()
    assert lib.slimb200_deflate_plan(members, len(sizes), C.byref(total), C.byref(ws), C.byref(bound)) == 0
    assert total.value == sum((n + 2047) // 2048 for n in sizes) and [m.first_chunk for m in members][:4] == [0, 1, 2, 3]
    assert ws.value >= total.value * 9232 and bound.value >= total.value * 9000
    word = np.array([3.8e-44], np.float32)  # the softmax denormal of an empty pillar
    r_word = zlib.crc32(word.tobytes()) ^ zlib.crc32(bytes(4))
    for m, n in zip(members, sizes):
        fill = np.full(n, word[0], np.float32).tobytes()
        assert _gf_mul(m.crc_geo, r_word) == zlib.crc32(fill) ^ zlib.crc32(bytes(len(fill))), n
    members[0].n_words = 0
    assert lib.slimb200_deflate_plan(members, 1, C.byref(total), C.byref(ws), C.byref(bound)) == -1  # empty member


def test_cell_stride_detection():
    enc = npz_stream.DeflateEncoder._cells
    bev = torch.zeros(2, 6, 5, 16)
    assert enc(bev[..., 7:9]) == (2, 16) and enc(bev[..., 5]) == (1, 16)
    assert enc(torch.zeros(2, 6, 5, 2)) == (60, 60) and enc(torch.zeros(3, 7)) == (7, 7)
    assert enc(bev[:, ::2, :, 5]) is None  # rows not evenly spaced with the cells


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_streams_equal_the_oracle_bytes():
    dev = torch.device("cuda:0")
    enc = npz_stream.DeflateEncoder(dev, slots=2)
    for name, a in sorted(CASES.items()):
        t = torch.from_numpy(a.reshape(1, -1).copy()).to(dev)
        enc.encode([t], slot=0)
        enc.start_download(0)
        got = enc.fetch(0)
        shape, stream, r = got.member(0, 0)
        assert stream == DO.encode_member(a), name
        assert r == DO.crc_remainder(a), name
        assert got.total_bytes == len(stream)


@pytest.mark.gpu
def test_gpu_strided_views_batches_and_files(tmp_path):
    """Members taken as channel slices of a packed channels-last buffer (how the decoder hands out the BEV maps), several
    samples and several views per call; the framed file is read by np.load bit for bit."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    B, H, W = 3, 80, 72
    bev = np.zeros((B, H, W, 16), np.float32)
    m = rng.random((B, H, W)) < 0.07
    bev[m] = rng.normal(size=(int(m.sum()), 16)).astype(np.float32)
    bev_d = torch.from_numpy(bev).to(dev)
    views = [bev_d[..., 7:9], bev_d[..., 5], bev_d[..., 10:12].contiguous()]
    enc = npz_stream.DeflateEncoder(dev, slots=2)
    side = torch.cuda.Stream(device=dev)
    for slot in (0, 1, 0):  # buffers are reused
        enc.encode(views, slot=slot)
        enc.start_download(slot, side)
        got = enc.fetch(slot)
        for vi, ref in enumerate((bev[..., 7:9], bev[..., 5], bev[..., 10:12])):
            for b in range(B):
                shape, stream, r = got.member(vi, b)
                assert shape == ref.shape[1:]
                assert zlib.decompress(stream, -15) == np.ascontiguousarray(ref[b]).tobytes()
                assert stream == DO.encode_member(np.ascontiguousarray(ref[b])) and r == DO.crc_remainder(ref[b])
    def framed(key, member):
        shape, stream, r = member
        return (key, shape, np.float32, stream, r)

    blob = npz_stream.build_npz([("static_threshold", np.array(0.25, np.float32))] +
                                [framed("flow_%d" % b, got.member(0, b)) for b in range(B)] + [framed("dyn_0", got.member(1, 0))])
    p = os.path.join(str(tmp_path), "s.npz")
    open(p, "wb").write(blob)
    z = np.load(p)
    for b in range(B):
        assert np.array_equal(z["flow_%d" % b].view(np.uint32), bev[b, :, :, 7:9].view(np.uint32))
    assert np.array_equal(z["dyn_0"], bev[0, :, :, 5])
    assert got.total_bytes < 0.2 * sum(v.numel() * 4 for v in views)


@pytest.mark.gpu
def test_gpu_full_size_round_trip():
    """BASELINE-size maps (8 x 640 x 640 x 2 and 8 x 640 x 640): zlib decodes every member to the input; CRC remainders
    combine to zlib's crc32 of header | data."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    B, H, W = 8, 640, 640
    occ = torch.rand(B, H, W, generator=g) < 0.05
    flow = torch.where(occ[..., None], torch.randn(B, H, W, 2, generator=g), torch.zeros(()))  # (x * False would leave -0.0 words)
    dyn = torch.where(occ, torch.rand(B, H, W, generator=g), torch.full((), 3.8e-44))  # empty pillars: the softmax denormal
    enc = npz_stream.DeflateEncoder(dev)
    enc.encode([flow.to(dev), dyn.to(dev)], slot=0)
    enc.start_download(0)
    got = enc.fetch(0)
    for vi, ref in enumerate((flow, dyn)):
        for b in range(B):
            shape, stream, r = got.member(vi, b)
            raw = ref[b].numpy().tobytes()
            assert zlib.decompress(stream, -15) == raw
            hdr = npz_stream.npy_header(shape)
            assert zlib.crc32(hdr + raw) == npz_stream.base_crc(shape) ^ r
    assert got.total_bytes < (flow.numel() + dyn.numel()) * 4 / 10  # D2H bytes down by more than 10x at 5 % occupancy
