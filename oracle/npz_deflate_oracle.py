"""TEST INFRASTRUCTURE (not a product path): CPU restatement of the stream ``csrc/npz_deflate.cu`` emits, byte for byte.

What is being replaced is the zlib pass of ``np.savez_compressed`` in ``liso/slim/experiment.py:459-471``; zlib itself
(not in /root/reference: CPython's bundled zlib) is the independent checker -- ``zlib.decompress(stream, -15)`` must
return the array's bytes and ``np.load`` must read the framed file.  This module restates the *encoder's* format
(RFC 1951 3.2.6 fixed Huffman code; chunks of 8 KB, each a non-final fixed block + an empty stored block; zero-word
runs as ``literal 0`` + distance-1 matches) sequentially, so that the GPU bytes can be compared exactly.

Pinned by ``tests/test_npz_stream.py``: every stream decodes with zlib to the input, for the edge cases of the token
rule (runs of 1..600 words, runs across chunk borders, remainders 1 and 2, ragged last chunk, all-zero, no zero).
"""
from __future__ import annotations

import zlib

import numpy as np

CHUNK_WORDS = 2048


def _rev(code: int, n: int) -> int:
    return int(format(code, "0%db" % n)[::-1], 2)


def lit_code(v: int):
    return (_rev(0x30 + v, 8), 8) if v < 144 else (_rev(0x190 + v - 144, 9), 9)


def match_code(length: int):
    """<length, distance 1>: Huffman code of the length symbol, its extra bits, five zero bits for distance symbol 0."""
    assert 3 <= length <= 258
    extra, ne = 0, 0
    if length == 258:
        sym = 285
    elif length <= 10:
        sym = 254 + length
    else:
        l = length - 3
        ne = l.bit_length() - 3
        sym = 257 + 4 * ne + (l >> ne)
        extra = l & ((1 << ne) - 1)
    bits, n = (_rev(sym - 256, 7), 7) if sym < 280 else (_rev(0xC0 + sym - 280, 8), 8)
    return bits | (extra << n), n + ne + 5


class _Bits:
    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, code):
        self.acc |= code[0] << self.n
        self.n += code[1]


def run_tokens(out: _Bits, run_bytes: int):
    """Tokens of a run of ``run_bytes`` zero bytes: literal, matches of 258, remainder as a match (>= 3) or literals."""
    out.put(lit_code(0))
    m = run_bytes - 1
    for _ in range(m // 258):
        out.put(match_code(258))
    r = m % 258
    if r >= 3:
        out.put(match_code(r))
    else:
        for _ in range(r):
            out.put(lit_code(0))


def encode_chunk(words: np.ndarray, last: bool) -> bytes:
    out = _Bits()
    out.put((2, 3))  # BFINAL 0, BTYPE 01
    i, n = 0, len(words)
    while i < n:
        if words[i]:
            v = int(words[i])
            for k in range(4):
                out.put(lit_code((v >> (8 * k)) & 255))
            i += 1
        else:
            a = i
            while i < n and words[i] == 0:
                i += 1
            run_tokens(out, 4 * (i - a))
    out.put((0, 7))  # end of block
    out.put((1 if last else 0, 3))  # empty stored block: byte-aligns the stream; BFINAL on the member's last chunk
    nbytes = (out.n + 7) // 8
    return out.acc.to_bytes(nbytes, "little") + b"\x00\x00\xff\xff"


def encode_member(array: np.ndarray) -> bytes:
    words = np.ascontiguousarray(array).view(np.uint32).reshape(-1)
    n_chunks = (len(words) + CHUNK_WORDS - 1) // CHUNK_WORDS
    return b"".join(encode_chunk(words[c * CHUNK_WORDS:(c + 1) * CHUNK_WORDS], c == n_chunks - 1) for c in range(n_chunks))


def crc_remainder(array: np.ndarray) -> int:
    """R(data): crc32(prefix | data) ^ crc32(prefix | zeros) for any prefix (here: none)."""
    raw = np.ascontiguousarray(array).tobytes()
    return zlib.crc32(raw) ^ zlib.crc32(bytes(len(raw)))
