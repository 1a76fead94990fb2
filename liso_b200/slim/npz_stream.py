"""``.npz`` files whose big members were DEFLATE-compressed on the GPU.

The reference writes one ``np.savez_compressed`` file per sample (``liso/slim/experiment.py:459-471``) and reads it back
with ``np.load`` (``torch_dataset_commons.py:614-616``).  An ``.npz`` is a zip archive of ``<key>.npy`` members, each a
``.npy`` header followed by the raw array bytes, deflated.  Here

* the device turns every exported map into a complete raw DEFLATE stream and the CRC-32 remainder of its bytes
  (``slimb200_deflate_encode``, ``csrc/npz_deflate.cu``) -- :class:`DeflateEncoder`; only those bytes are downloaded;
* the host puts a *stored* block with the ``.npy`` header in front of each stream, finishes the CRC with a per-shape
  constant, and writes the zip records around it -- :func:`build_npz`.  No zlib call touches the maps.
"""
from __future__ import annotations

import ctypes as C
import io
import struct
import zlib
from typing import Dict, List, Sequence, Tuple

import numpy as np

_DOS_TIME, _DOS_DATE = 0, (1 << 5) | 1  # 1980-01-01 00:00 (zip members need a valid date, nobody reads it)


def npy_header(shape, dtype=np.float32) -> bytes:
    """The bytes ``np.save`` puts in front of the data of a C-ordered array (format 1.0, what ``np.savez`` writes)."""
    f = io.BytesIO()
    np.lib.format.write_array_header_1_0(
        f, {"descr": np.lib.format.dtype_to_descr(np.dtype(dtype)), "fortran_order": False, "shape": tuple(int(s) for s in shape)})
    return f.getvalue()


_BASE_CRC: Dict[Tuple[tuple, str], int] = {}


def base_crc(shape, dtype=np.float32) -> int:
    """crc32(npy header | zeros of the array's size): the device's CRC remainder R of the data completes it,
    crc32(header | data) = base ^ R (CRC-32 is affine; the remainder of leading zeros is zero)."""
    key = (tuple(int(s) for s in shape), np.dtype(dtype).str)
    if key not in _BASE_CRC:
        n = int(np.prod(key[0], dtype=np.int64)) * np.dtype(dtype).itemsize
        _BASE_CRC[key] = zlib.crc32(bytes(n), zlib.crc32(npy_header(shape, dtype)))
    return _BASE_CRC[key]


def stored_block(payload: bytes, final: bool = False) -> bytes:
    """One stored DEFLATE block (RFC 1951 3.2.4) starting on a byte boundary."""
    assert len(payload) < 65536
    return struct.pack("<BHH", 1 if final else 0, len(payload), len(payload) ^ 0xFFFF) + payload


def _local_record(name: bytes, method: int, crc: int, csize: int, usize: int) -> bytes:
    return struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0, method, _DOS_TIME, _DOS_DATE, crc & 0xFFFFFFFF, csize, usize, len(name), 0) + name


def _central_record(name: bytes, method: int, crc: int, csize: int, usize: int, offset: int) -> bytes:
    return struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 20, 20, 0, method, _DOS_TIME, _DOS_DATE, crc & 0xFFFFFFFF, csize, usize,
                       len(name), 0, 0, 0, 0, 0, offset) + name


def build_npz(members: Sequence[tuple]) -> bytes:
    """The bytes of an ``.npz`` file.  ``members``: ``(key, ndarray)`` -- a small array, stored as is -- or
    ``(key, shape, dtype, stream, crc_remainder)`` -- an array whose data the device has deflated: ``stream`` is the
    complete raw DEFLATE stream of the array's bytes, ``crc_remainder`` the device's R."""
    parts: List[bytes] = []
    central: List[bytes] = []
    offset = 0
    for m in members:
        name = (m[0] + ".npy").encode()
        if len(m) == 2:
            f = io.BytesIO()
            np.lib.format.write_array(f, np.asanyarray(m[1]), allow_pickle=False)
            body = f.getvalue()
            method, crc, usize = 0, zlib.crc32(body), len(body)
            chunks = [body]
        else:
            _, shape, dtype, stream, r = m
            hdr = npy_header(shape, dtype)
            method, crc = 8, base_crc(shape, dtype) ^ int(r)
            usize = len(hdr) + int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
            chunks = [stored_block(hdr), stream]
        csize = sum(len(c) for c in chunks)
        rec = _local_record(name, method, crc, csize, usize)
        central.append(_central_record(name, method, crc, csize, usize, offset))
        parts.append(rec)
        parts.extend(chunks)
        offset += len(rec) + csize
    cd = b"".join(central)
    parts.append(cd)
    parts.append(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, len(central), len(central), len(cd), offset, 0))
    return b"".join(parts)


class EncodedBatch:
    """What :meth:`DeflateEncoder.fetch` hands out: the streams of all members of one encode call, on the host."""

    def __init__(self, table: np.ndarray, data, shapes, batch: int, rank=None):
        self.table, self.data, self.shapes, self.batch = table, data, shapes, batch
        self.rank = list(rank) if rank is not None else list(range(len(shapes)))
        self.total_bytes = int(table[-1, 0])

    def member(self, view: int, b: int):
        """(shape, stream bytes, crc remainder) of sample ``b`` of the ``view``-th tensor given to ``encode``."""
        off, n, r, _ = (int(v) for v in self.table[b * len(self.shapes) + self.rank[view]])
        return self.shapes[view], bytes(self.data[off:off + n]), r


class DeflateEncoder:
    """Device side of the GPU ``.npz`` writer for one device: CRC tables, member plans (cached per set of views), the
    shared workspace, and ``slots`` output buffers so that the download of one batch overlaps the encoding of the next.

    ``encode(views, slot)`` enqueues the kernels on the current stream; every ``views[v][b]`` (fp32, batch first, the
    cells of a sample uniformly strided -- e.g. a channel slice of a packed channels-last buffer) becomes one member.
    ``start_download(slot, stream)`` / ``fetch(slot)`` bring the streams to the host: first the member table, then --
    once the sizes are known -- exactly the compressed bytes."""

    def __init__(self, device, slots: int = 3):
        import torch

        from .. import _lib

        self.lib, self._lib, self.device, self.torch = _lib.load(), _lib, torch.device(device), torch
        if self.device.type != "cuda":
            raise RuntimeError("DeflateEncoder runs on CUDA only: the export writer has no CPU fallback")
        self.tables = torch.empty(_lib.DEFLATE_TABLE_BYTES // 4, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.slimb200_deflate_init(self.tables.data_ptr(), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        self._plans: Dict[tuple, dict] = {}
        self._ws = None
        self.slots = [dict(out=None, table_dev=None, table_host=None, host=None, plan=None, evt=None) for _ in range(slots)]

    # ---- planning --------------------------------------------------------------------------------------------------
    @staticmethod
    def _cells(v):
        """(words per cell, cell stride) of one sample of a batched view, or None when the cells are not uniformly strided."""
        shape, stride = list(v.shape[1:]), list(v.stride()[1:])
        dims = [(n, s) for n, s in zip(shape, stride) if n > 1]
        if not dims:
            return 1, 1
        wpc = 1
        while dims and dims[-1][1] == wpc:  # innermost contiguous run of words = one cell
            wpc *= dims.pop()[0]
        if not dims:
            return wpc, wpc
        cs = dims[-1][1]
        span = cs
        for n, s in reversed(dims):  # the cells themselves must be evenly spaced
            if s != span:
                return None
            span *= n
        return wpc, cs

    def _plan(self, views):
        torch = self.torch
        key = tuple((v.data_ptr(), tuple(v.shape), tuple(v.stride())) for v in views)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        B = int(views[0].shape[0])
        V = len(views)
        # member order = sample-major, and inside a sample the views that are slices of the same buffer next to each other
        # (flow and dynamicness of one direction live in the same packed decoder rows): the second one finds the rows in L2
        order = sorted(range(V), key=lambda i: (views[i].untyped_storage().data_ptr(), i))
        rank = [0] * V
        for r, i in enumerate(order):
            rank[i] = r
        members = (self._lib.DeflateMember * (V * B))()
        for vi, v in enumerate(views):
            if v.dtype != torch.float32 or not v.is_cuda or int(v.shape[0]) != B:
                raise ValueError("DeflateEncoder.encode takes fp32 CUDA tensors with a common batch size")
            cells = self._cells(v)
            if cells is None:
                raise ValueError("view with irregular strides %r: pass a contiguous tensor" % (tuple(v.stride()),))
            n_words = int(np.prod(v.shape[1:], dtype=np.int64))
            for b in range(B):
                m = members[b * V + rank[vi]]
                m.src, m.words_per_cell, m.cell_stride, m.n_words = v.data_ptr() + 4 * b * v.stride(0), cells[0], cells[1], n_words
        total, ws_bytes, bound = C.c_int64(), C.c_size_t(), C.c_size_t()
        self._lib.check(self.lib.slimb200_deflate_plan(members, len(members), C.byref(total), C.byref(ws_bytes), C.byref(bound)))
        host = torch.frombuffer(bytearray(bytes(members)), dtype=torch.uint8)
        plan = dict(n=len(members), B=B, total_chunks=int(total.value), ws_bytes=int(ws_bytes.value), bound=int(bound.value),
                    members_dev=host.to(self.device), shapes=[tuple(v.shape[1:]) for v in views], rank=rank,
                    raw_bytes=sum(4 * int(np.prod(v.shape, dtype=np.int64)) for v in views))
        if len(self._plans) > 16:
            self._plans.clear()
        self._plans[key] = plan
        return plan

    # ---- device work -----------------------------------------------------------------------------------------------
    def encode(self, views, slot: int = 0):
        torch = self.torch
        plan = self._plan(views)
        s = self.slots[slot]
        if self._ws is None or self._ws.numel() < plan["ws_bytes"]:
            self._ws = torch.empty(plan["ws_bytes"], dtype=torch.uint8, device=self.device)
        if s["out"] is None or s["out"].numel() < plan["bound"]:
            s["out"] = torch.empty(plan["bound"], dtype=torch.uint8, device=self.device)
        if s["table_dev"] is None or s["table_dev"].shape[0] != plan["n"] + 1:
            s["table_dev"] = torch.empty((plan["n"] + 1, 4), dtype=torch.int32, device=self.device)
            s["table_host"] = torch.empty((plan["n"] + 1, 4), dtype=torch.int32).pin_memory()
        with torch.cuda.device(self.device):
            self._lib.check(self.lib.slimb200_deflate_encode(
                plan["members_dev"].data_ptr(), plan["n"], plan["total_chunks"], self.tables.data_ptr(), self._ws.data_ptr(),
                self._ws.numel(), s["out"].data_ptr(), s["out"].numel(), s["table_dev"].data_ptr(),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        s["plan"] = plan
        return plan

    def start_download(self, slot: int, stream=None):
        """Enqueue the download of the member table on ``stream`` (default: current) behind the encode kernels."""
        torch = self.torch
        s = self.slots[slot]
        cur = torch.cuda.current_stream(self.device)
        stream = stream or cur
        if stream != cur:
            stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            s["table_host"].copy_(s["table_dev"], non_blocking=True)
            s["evt"] = torch.cuda.Event()
            s["evt"].record(stream)
        s["stream"] = stream

    def fetch_begin(self, slot: int, stream=None):
        """Wait for the member table, then enqueue the download of exactly the compressed bytes (on ``stream``, default:
        the stream the table came on)."""
        torch = self.torch
        s = self.slots[slot]
        if stream is not None:
            s["stream"] = stream
        s["evt"].synchronize()
        table = s["table_host"].numpy().view(np.uint32)
        if int(table[-1, 1]):
            raise RuntimeError("DeflateEncoder: output buffer overflow (%d bytes needed)" % int(table[-1, 0]))
        total = int(table[-1, 0])
        if s["host"] is None or s["host"].numel() < total:
            s["host"] = torch.empty(max(total + total // 2, 1 << 20), dtype=torch.uint8).pin_memory()
        with torch.cuda.stream(s["stream"]):
            s["host"][:total].copy_(s["out"][:total], non_blocking=True)
            s["evt2"] = torch.cuda.Event()
            s["evt2"].record(s["stream"])
        return total

    def fetch_end(self, slot: int) -> EncodedBatch:
        s = self.slots[slot]
        s["evt2"].synchronize()
        table = s["table_host"].numpy().view(np.uint32).copy()
        return EncodedBatch(table, memoryview(s["host"].numpy()), s["plan"]["shapes"], s["plan"]["B"], s["plan"]["rank"])

    def fetch(self, slot: int = 0) -> EncodedBatch:
        self.fetch_begin(slot)
        return self.fetch_end(slot)
