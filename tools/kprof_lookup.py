#!/usr/bin/env python
"""A handful of launches of the lookup kernels on the bench shapes (K, B=8), for `ncu --set full` captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from liso_b200 import _lib  # noqa: E402
from liso_b200.slim.corr import CorrBlock, PackedLookupConv, coords_grid  # noqa: E402


def main():
    B, h, w = 8, 80, 80
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    f1 = torch.randn(B, 128, h, w, generator=g).to(dev)
    f2 = torch.randn(B, 128, h, w, generator=g).to(dev)
    coords = (coords_grid(B, h, w, dev) + 1.5 * torch.randn(B, 2, h, w, generator=g).to(dev)).contiguous()
    blk_c = CorrBlock(f1, f2, num_levels=4, radius=3)
    blk_l = CorrBlock(f1.contiguous(memory_format=torch.channels_last), f2.contiguous(memory_format=torch.channels_last), 4, 3)
    wconv = (torch.randn(96, 196, generator=g) / 14.0).to(dev)
    packed = PackedLookupConv(wconv, torch.randn(96, generator=g).to(dev), 4, 3)
    which = sys.argv[1:] or ["nchw", "nhwc", "fused"]
    with torch.no_grad():
        for _ in range(3):
            if "nchw" in which:
                blk_c(coords)
            if "nhwc" in which:
                blk_l(coords)
            if "fused" in which:
                blk_l.lookup_conv(coords, packed, relu=True)
            if "gen0" in which:
                lib.slimb200_lookup_generation(0)
                blk_l(coords)
                lib.slimb200_lookup_generation(1)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
