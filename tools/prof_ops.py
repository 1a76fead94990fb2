#!/usr/bin/env python
"""torch.profiler attribution of one resident bench step to ATen ops / modules (measurement tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.slim.slim import SLIM
from liso_b200.synth import make_sample_dicts
from liso_b200.weights import synth_weights_like

fmt = sys.argv[1] if len(sys.argv) > 1 else "channels_last"
B = 8
dev = torch.device("cuda:0")
cfg = make_cfg("K")
cfg.network["b200_canvas_memory_format"] = fmt
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.benchmark = True
model = SLIM(cfg, decode_iterations=(sys.argv[2] if len(sys.argv) > 2 else "all"), static_aggregation=(len(sys.argv) <= 2 or sys.argv[2] == "all")).eval()
model.load_state_dict(synth_weights_like(model.state_dict(), 0))
model = model.to(dev)
if fmt == "channels_last":
    model = model.to(memory_format=torch.channels_last)
s0, s1 = make_sample_dicts(WORKLOADS["K"], [1000 + i for i in range(B)])
to = lambda s: {"pcl_full_no_ground_ta": [t.to(dev) for t in s["pcl_full_no_ground_ta"]], "pcl_ta": {k: v.to(dev) for k, v in s["pcl_ta"].items()}}
d0, d1 = to(s0), to(s1)
with torch.no_grad():
    for _ in range(3):
        model(d0, d1, None)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False, record_shapes=True) as prof:
        model(d0, d1, None)
        torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=40, max_shapes_column_width=70))
