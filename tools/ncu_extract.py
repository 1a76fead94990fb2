"""Extract the columns the profiles/ summaries keep from an `ncu --set full` report.

usage: python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_hot_kernels.csv [profiles/ncu_traffic.json]
The optional third argument rewrites the per-kernel DRAM traffic table bench.py reads (`roofline.traffic`): the mean of
dram__bytes_read.sum + dram__bytes_write.sum over the captured launches of each kernel.
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(r"^(Kernel Name|dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|dram__cycles_active.*|"
                  r"gpu__dram_throughput.*|gpu__time_duration\.sum|launch__.*|lts__throughput.*|LTS\.Triage.*lts__throughput.*|"
                  r"l1tex__m_l1tex2xbar_req_cycles_active.*|sm__pipe_tensor_cycles_active.*|sm__throughput.*|sm__warps_active.*|"
                  r"smsp__inst_executed\.sum|smsp__issue_active.*)$")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(head) if KEEP.match(h) and not h.startswith("launch__cluster") and "func_cache" not in h
            and "context" not in h and "device" not in h and "stream" not in h and "sub_launch" not in h and "thread_count" not in h
            and "waves" not in h and "sm_count" not in h and "uses_" not in h and "occupancy_per" not in h and "occupancy_cluster" not in h]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in body:
            w.writerow([re.sub(r"\(.*", "", r[i]) if head[i] == "Kernel Name" else r[i] for i in cols])
    if len(sys.argv) > 3:
        ki, ri, wi = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        acc = {}
        for r in body:
            name = re.sub(r"^.*::", "", re.sub(r"[<(].*", "", r[ki].replace("void ", "").replace("<unnamed>::", "")))
            t = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
            acc.setdefault(name, []).append(t)
        table = {k: sum(v) / len(v) for k, v in acc.items()}
        table["_source"] = out
        json.dump(table, open(sys.argv[3], "w"), indent=1)
        print(table)


if __name__ == "__main__":
    main()
