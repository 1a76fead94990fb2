"""Diagnostics of the export loop on one GPU: graph captures, per-phase times (forward / encode / download / framing)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.slim import export
from liso_b200.slim.slim import SLIM
from liso_b200.synth import SyntheticExportDataset
from liso_b200.weights import synth_weights_like

wl = sys.argv[1] if len(sys.argv) > 1 else "K"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cfg = make_cfg(wl)
cfg.network["b200_canvas_memory_format"] = "channels_last"
model = SLIM(cfg, decode_iterations="last", static_aggregation=False).eval()
model.load_state_dict(synth_weights_like(model.state_dict(), 0))
model = model.to(dev).to(memory_format=torch.channels_last)
ds = SyntheticExportDataset(WORKLOADS[wl], n, frames=frames, pool=4, raw=True).prepare(4)
net = model.raft_network

# phase timing with synchronisation (not a throughput number)
pipe = export.ExportPipeline(model, dev, compress=True)
items = [ds[i] for i in range(8)]
batch = tuple(export._pin(d) for d in export.collate_pairs([tuple(it[1:]) for it in items]))
from liso_b200.datasets import preprocess_scans
from liso_b200.slim.npz_stream import DeflateEncoder

enc = DeflateEncoder(dev)
model.outputs_alias_static_buffers = True
for rep in range(4):
    t0 = time.perf_counter()
    up = tuple(pipe._upload(s, 0, "t%d" % t) for t, s in enumerate(batch))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    with torch.no_grad():
        st = tuple(preprocess_scans(s["pcl_full_w_ground_ta"], cfg) for s in up)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if frames == 3:
            by = model.forward_triple(*st)
            mods = [by[d][-1].modified_network_output for d in export.DIRECTIONS_TRIPLE]
        else:
            pf, pb = model(st[0], st[1], None)
            mods = [pf[-1].modified_network_output, pb[-1].modified_network_output]
    torch.cuda.synchronize(); t3 = time.perf_counter()
    enc.encode([m.static_flow for m in mods] + [m.dynamicness for m in mods], 0)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    enc.start_download(0); got = enc.fetch(0); t5 = time.perf_counter()
    print("rep %d: upload %.1f ms, preprocess %.1f, forward %.1f, encode %.2f, download %.1f (%.2f MB of %.1f MB raw); captures so far %d" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), got.total_bytes / 1e6,
        sum(m.static_flow.numel() * 4 + m.dynamicness.numel() * 4 for m in mods) / 1e6, getattr(net, "n_graph_captures", 0)))
sizes = {}
for v, key in enumerate(export.export_keys(len(mods) * 1 if False else len(mods))):
    sizes[key] = len(got.member(v, 0)[1])
print("bytes per member (sample 0):", sizes)
import zlib
shape, stream, r = got.member(0, 0)
raw = zlib.decompress(stream, -15)
print("zlib level 6 on the same flow map: %d bytes (ours %d)" % (len(zlib.compress(raw, 6)), len(stream)))
shape, stream, r = got.member(len(mods), 0)
raw = zlib.decompress(stream, -15)
print("zlib level 6 on the same dynamicness map: %d bytes (ours %d)" % (len(zlib.compress(raw, 6)), len(stream)))

c0 = getattr(net, "n_graph_captures", 0)
t = time.perf_counter()
res = export.run_flow_export(model, ds, "/dev/shm/slimb200_diag", cfg.data.bev_range_m, batch_size=8, device=dev, writer_workers=8,
                             compress_on_gpu=True, loader_workers=4, unlink_after_write=True)
print("run_flow_export: %d samples in %.2f s = %.1f samples/s; graph captures during the run: %d; %s" % (
    res["pairs"], res["elapsed_s_max"], res["pairs"] / res["elapsed_s_max"], getattr(net, "n_graph_captures", 0) - c0, res))
t = time.perf_counter()
k = 0
for chunk in export.iterate_batches(range(min(n, 32)), 8):
    items = [ds[i] for i in chunk]
    b = tuple(export._pin(d) for d in export.collate_pairs([tuple(it[1:]) for it in items]))
    k += 1
print("loader alone: %.1f ms per batch of 8" % (1e3 * (time.perf_counter() - t) / k))
