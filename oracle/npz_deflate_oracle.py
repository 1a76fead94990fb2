"""TEST INFRASTRUCTURE (not a product path): CPU restatement of the stream ``csrc/npz_deflate.cu`` emits, byte for byte.

What is being replaced is the zlib pass of ``np.savez_compressed`` in ``liso/slim/experiment.py:459-471``; zlib itself
(not in /root/reference: CPython's bundled zlib) is the independent checker -- ``zlib.decompress(stream, -15)`` must
return the array's bytes and ``np.load`` must read the framed file.  This module restates the *encoder's* format
(RFC 1951 3.2.6 fixed Huffman code; chunks of 8 KB, each a non-final fixed block + an empty stored block; a word equal to
the word before it continues a run, runs are distance-4 matches; a zero word that starts a run is ``literal 0`` +
``<distance 1, length 3>``) sequentially, so that the GPU bytes can be compared exactly.

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


def match_code(length: int, dist_sym: int = 0):
    """<length, distance>: Huffman code of the length symbol, its extra bits, five bits of the distance symbol (0..3 =
    distances 1..4, no extra bits)."""
    assert 3 <= length <= 258 and 0 <= dist_sym <= 3
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
    bits |= extra << n
    n += ne
    return bits | (_rev(dist_sym, 5) << n), n + 5


DIST_1, DIST_4 = 0, 3


class _Bits:
    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, code):
        self.acc |= code[0] << self.n
        self.n += code[1]


def run_tokens(out: _Bits, run_bytes: int):
    """Tokens of a run of ``run_bytes`` bytes that repeat the word in front of the run: matches of 258 at distance 4, the
    remainder as one more match; a remainder of 1 or 2 bytes borrows 3 bytes from the last full match."""
    nfull, r = divmod(run_bytes, 258)
    d = 3 if r in (1, 2) else 0
    for k in range(nfull):
        out.put(match_code(258 - d if k == nfull - 1 else 258, DIST_4))
    if r + d >= 3:
        out.put(match_code(r + d, DIST_4))


def encode_chunk(words: np.ndarray, last: bool) -> bytes:
    out = _Bits()
    out.put((2, 3))  # BFINAL 0, BTYPE 01
    i, n = 0, len(words)
    while i < n:
        if i == 0 or words[i] != words[i - 1]:
            v = int(words[i])
            if v == 0:
                out.put(lit_code(0))
                out.put(match_code(3, DIST_1))
            else:
                for k in range(4):
                    out.put(lit_code((v >> (8 * k)) & 255))
            i += 1
        else:
            a = i
            while i < n and words[i] == words[i - 1]:
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
