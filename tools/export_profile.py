"""cProfile of the launch thread of the export loop (where do the ~4 ms of host time per batch go?)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.slim import export
from liso_b200.slim.slim import SLIM
from liso_b200.synth import SyntheticExportDataset
from liso_b200.weights import synth_weights_like

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 480
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cfg = make_cfg("K")
cfg.network["b200_canvas_memory_format"] = "channels_last"
model = SLIM(cfg, decode_iterations="last", static_aggregation=False).eval()
model.load_state_dict(synth_weights_like(model.state_dict(), 0))
model = model.to(dev).to(memory_format=torch.channels_last)
ds = SyntheticExportDataset(WORKLOADS["K"], n, frames=frames, pool=4, raw=True, motion="shift").prepare(4)
kw = dict(batch_size=8, device=dev, writer_workers=4, compress_on_gpu=True, loader_workers=3, unlink_after_write=True)
warm = SyntheticExportDataset(WORKLOADS["K"], 16, frames=frames, pool=4, raw=True, motion="shift")
warm._cache = ds._cache
export.run_flow_export(model, warm, "/dev/shm/slimb200_prof_w", cfg.data.bev_range_m, **kw)
pr = cProfile.Profile()
pr.enable()
res = export.run_flow_export(model, ds, "/dev/shm/slimb200_prof", cfg.data.bev_range_m, **kw)
pr.disable()
print({k: round(v, 3) for k, v in res.items()})
print("samples/s", res["pairs"] / res["elapsed_s_max"])
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(30)
