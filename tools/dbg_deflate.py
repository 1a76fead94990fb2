import sys, ctypes as C
sys.path.insert(0, "/root/repo")
import torch
from liso_b200 import _lib
from liso_b200.slim.npz_stream import DeflateEncoder
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1)
B, H, W = 8, 640, 640
vs = []
for _ in range(2):
    occ = torch.rand(B, H, W, generator=g) < 0.05
    vs.append(torch.where(occ[..., None], torch.randn(B, H, W, 2, generator=g), torch.zeros(())).to(dev))
    vs.append(torch.where(occ, torch.rand(B, H, W, generator=g), torch.full((), 3.8e-44)).to(dev))
enc = DeflateEncoder(dev, slots=1)
for name, views in (("flow+dyn", vs), ("flow only", vs[0::2]), ("dyn only", vs[1::2])):
    for dbg in (0, 1, 2, 3, 4, 8, 15):
        raw.slimb200_deflate_debug(dbg)
        for _ in range(3): enc.encode(views, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): enc.encode(views, 0)
        e1.record(); torch.cuda.synchronize()
        print("%-10s dbg %2d (1 no crc, 2 no tokens, 4 no copy-out, 8 no loads): %.4f ms per encode" % (name, dbg, e0.elapsed_time(e1) / 20))
