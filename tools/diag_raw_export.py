import os, sys, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from liso_b200.config import WORKLOADS, make_cfg
from liso_b200.datasets import preprocess_scans
from liso_b200.slim import export
from liso_b200.slim.slim import SLIM
from liso_b200.synth import SyntheticExportDataset
from liso_b200.weights import synth_weights_like
cuda = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
cfg = make_cfg("T")
model = SLIM(cfg, decode_iterations="last", static_aggregation=False).eval()
model.load_state_dict(synth_weights_like(model.state_dict(), 0))
model = model.to(cuda)
ds = SyntheticExportDataset(WORKLOADS["T"], 7, frames=2, pool=2, raw=True)
for mode in ("compress+loader", "compress", "raw"):
    tmp = tempfile.mkdtemp()
    out = export.run_flow_export(model, ds, tmp, cfg.data.bev_range_m, batch_size=3, device=cuda, writer_workers=2,
                                 compress_on_gpu=mode != "raw", loader_workers=2 if "loader" in mode else 0)
    for chunk in ((0, 1, 2), (3, 4, 5), (6, 6, 6), (6,)):
        items = [ds[i] for i in chunk]
        with torch.no_grad():
            pf, pb = model(preprocess_scans([it[1]["pcl_full_w_ground_ta"].to(cuda) for it in items], cfg),
                           preprocess_scans([it[2]["pcl_full_w_ground_ta"].to(cuda) for it in items], cfg), None)
        for b, i in enumerate(chunk):
            z = np.load(os.path.join(tmp, "%06d.npz" % i))
            a = z["bev_raw_flow_t0_t1"]; r = pf[-1].modified_network_output.static_flow[b].cpu().numpy()
            print(mode, chunk, i, "equal" if np.array_equal(a, r) else "DIFF n=%d max=%.3e nz file=%d ref=%d" % ((a != r).sum(), np.abs(a - r).max(), (a != 0).sum(), (r != 0).sum()))
