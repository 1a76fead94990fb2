#!/usr/bin/env python
"""Per-kernel timing of the slimb200 library on the bench workload (K, B=8) without the stock PyTorch part.
Measurement tool for kernel iteration; the judged numbers come from bench.py."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from liso_b200 import _lib  # noqa: E402
from liso_b200.config import WORKLOADS, make_cfg  # noqa: E402
from liso_b200.networks.pcl_to_feature_grid import PointsPillarFeatureNetWrapper  # noqa: E402
from liso_b200.slim.corr import CorrBlock, coords_grid  # noqa: E402
from liso_b200.synth import make_sample_dicts  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="K")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--skip-pillar", action="store_true")
    ap.add_argument("--skip-corr", action="store_true")
    ap.add_argument("--train-bn", action="store_true")
    ap.add_argument("--skip-deflate", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    W = WORKLOADS[args.workload]
    cfg = make_cfg(args.workload)
    H, Wd = W["img_grid_size"]
    B = args.batch
    runs = []
    if not args.skip_pillar:
        s0, _ = make_sample_dicts(W, [1000 + i for i in range(B)])
        clouds = [t.to(dev) for t in s0["pcl_full_no_ground_ta"]]
        for fmt in ("channels_last", "contiguous"):
            m = PointsPillarFeatureNetWrapper(cfg, canvas_memory_format=fmt).to(dev)
            m.train(args.train_bn)
            runs.append(("pillar " + fmt[:8], (lambda mm: (lambda: mm(clouds)))(m)))
    if not args.skip_pillar:
        from liso_b200.datasets import preprocess_scans

        raw = [torch.cat([c, c[: c.shape[0] // 3] * torch.tensor([1.0, 1.0, 0.0, 1.0], device=dev)
                          + torch.tensor([0.0, 0.0, -1.73, 0.0], device=dev)]) for c in clouds]
        runs.append(("preprocess raw", lambda: preprocess_scans(raw, cfg)))
    if not args.skip_corr:
        g = torch.Generator(device="cpu").manual_seed(0)
        h, w = H // 8, Wd // 8
        f1 = torch.randn(B, 128, h, w, generator=g).to(dev)
        f2 = torch.randn(B, 128, h, w, generator=g).to(dev)
        coords = (coords_grid(B, h, w, dev) + 1.5 * torch.randn(B, 2, h, w, generator=g).to(dev)).contiguous()
        state = {}

        f1c, f2c = f1.contiguous(memory_format=torch.channels_last), f2.contiguous(memory_format=torch.channels_last)

        def build():
            state["blk"] = CorrBlock(f1c, f2c, num_levels=4, radius=3)

        def build_nchw():
            state["blk"] = CorrBlock(f1, f2, num_levels=4, radius=3)

        def look():
            for _ in range(6):
                state["out"] = state["blk"](coords)

        runs.append(("corr_build nchw", build_nchw))
        runs.append(("corr_build nhwc", build))
        runs.append(("corr_lookup x6", look))
        grid = coords_grid(B, h, w, dev).contiguous()

        def look_int():
            for _ in range(6):
                state["out"] = state["blk"](grid)

        runs.append(("corr_lookup integer coords", look_int))

        # generations / layouts of the lookup, the fused lookup + 1x1 convolution and what it replaces
        import torch.nn.functional as F

        blk_l = CorrBlock(f1c, f2c, num_levels=4, radius=3)
        blk_c = CorrBlock(f1, f2, num_levels=4, radius=3)
        wconv = (torch.randn(96, 196, 1, 1, generator=g) / 14.0).to(dev).contiguous(memory_format=torch.channels_last)
        bconv = torch.randn(96, generator=g).to(dev)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cudnn.benchmark = True

        def gen(n, fn):
            def run():
                lib.slimb200_lookup_generation(n)
                try:
                    for _ in range(6):
                        state["out"] = fn()
                finally:
                    lib.slimb200_lookup_generation(1)
            return run

        runs.append(("lookup gen0 nhwc x6", gen(0, lambda: blk_l(coords))))
        runs.append(("lookup gen1 nhwc x6", gen(1, lambda: blk_l(coords))))
        runs.append(("lookup gen0 nchw x6", gen(0, lambda: blk_c(coords))))
        runs.append(("lookup gen1 nchw x6", gen(1, lambda: blk_c(coords))))
        runs.append(("lookup gen1 nchw int x6", gen(1, lambda: blk_c(grid))))
        runs.append(("lookup gen2 nhwc x6", gen(2, lambda: blk_l(coords))))
        runs.append(("lookup gen2 nchw x6", gen(2, lambda: blk_c(coords))))
        runs.append(("lookup gen2 nchw int x6", gen(2, lambda: blk_c(grid))))
        # a smooth flow field (what the network asks for): neighbouring pixels look up neighbouring windows
        smooth = (coords_grid(B, h, w, dev) + 4.0 * F.interpolate(torch.randn(B, 2, 5, 5, generator=g).to(dev), size=(h, w),
                                                                    mode="bicubic", align_corners=True)).contiguous()
        runs.append(("lookup gen2 nchw smooth x6", gen(2, lambda: blk_c(smooth))))
        runs.append(("lookup gen0 nchw smooth x6", gen(0, lambda: blk_c(smooth))))
        runs.append(("probe (loads only) x6", gen(9, lambda: blk_c(coords))))
        runs.append(("probe (loads only) smooth x6", gen(9, lambda: blk_c(smooth))))
        runs.append(("probe (loads only) int x6", gen(9, lambda: blk_c(grid))))
        runs.append(("lookup+conv unfused x6", gen(2, lambda: torch.cudnn_convolution_relu(blk_l(coords), wconv, bconv, (1, 1), (0, 0), (1, 1), 1))))
        from liso_b200.slim.corr import PackedLookupConv

        packed = PackedLookupConv(wconv, bconv, 4, 3)
        def cgen(n, fn):
            def run():
                prev = lib.slimb200_lookup_conv_generation(n)
                try:
                    for _ in range(6):
                        state["out"] = fn()
                finally:
                    lib.slimb200_lookup_conv_generation(prev)
            return run

        for cg in (3, 4):
            runs.append(("fused gen%d x6" % cg, cgen(cg, lambda: blk_l.lookup_conv(coords, packed, relu=True))))
            runs.append(("fused gen%d int x6" % cg, cgen(cg, lambda: blk_l.lookup_conv(grid, packed, relu=True))))
            runs.append(("fused gen%d smooth x6" % cg, cgen(cg, lambda: blk_l.lookup_conv(smooth, packed, relu=True))))
    if not args.skip_deflate:
        # export writer: the maps of one batch (flow + dynamicness of both directions) as channel slices of two packed
        # 16-float decoder buffers, ~5 % occupied; empty cells: flow 0, dynamicness = the softmax denormal
        from liso_b200.slim.npz_stream import DeflateEncoder

        g = torch.Generator(device="cpu").manual_seed(1)
        bevs = []
        for _ in range(2):
            occ = torch.rand(B, H, Wd, generator=g) < 0.05
            bev = torch.zeros(B, H, Wd, 16)
            bev[..., 5] = torch.where(occ, torch.rand(B, H, Wd, generator=g), torch.full((), 3.8e-44))
            bev[..., 7:9] = torch.where(occ[..., None], torch.randn(B, H, Wd, 2, generator=g), torch.zeros(()))
            bevs.append(bev.to(dev))
        enc = DeflateEncoder(dev, slots=1)
        views = [b[..., 7:9] for b in bevs] + [b[..., 5] for b in bevs]
        packed = [v.contiguous() for v in views]
        runs.append(("deflate strided", lambda: enc.encode(views, 0)))
        runs.append(("deflate packed", lambda: enc.encode(packed, 0)))
    with torch.no_grad():
        for name, fn in runs:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            lib.slimb200_profile_begin()
            for _ in range(args.reps):
                fn()
            torch.cuda.synchronize()
            ms_k = (C.c_float * _lib.N_KERNELS)()
            n_k = (C.c_int64 * _lib.N_KERNELS)()
            _lib.check(lib.slimb200_profile_end(ms_k, n_k))
            for i in range(_lib.N_KERNELS):
                if n_k[i]:
                    print("%-16s %-24s %8.4f ms/launch  (%d launches)" % (name, lib.slimb200_kernel_name(i).decode(), ms_k[i] / n_k[i], n_k[i]))
            # the same calls back to back behind a spin kernel (queue full before the GPU starts), two events only:
            # stream time per call without per-launch event brackets
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(int(30e-3 * 1.9e9))
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print("%-16s %-24s %8.4f ms/call back-to-back (2 events, %d calls)" % (name, "(stream time)", e0.elapsed_time(e1) / args.reps, args.reps))
    sys.stdout.flush()


if __name__ == "__main__":
    main()
