"""Turn an `ncu --set full` capture of the dominant conv kernel into profiles/r2_ncu_traffic.json, the file bench.py reads for
`roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch; nothing is hard-coded in bench.py).

    python tools/ncu_traffic.py gpurun_out/r2_ncu_halo.ncu-rep cfg2 [kernel-substring]

Picks the launch of the kernel with the largest duration (the local_fusion 192 -> 192 conv in the network step) unless the
capture holds one launch only.  Records the git commit the capture was taken at."""
import csv
import json
import os
import subprocess
import sys

rep, workload = sys.argv[1], sys.argv[2]
kname = sys.argv[3] if len(sys.argv) > 3 else "k_conv_tc"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def f(r, name, table):
    return float(r[idx[name]].replace(",", "")) * table.get(units[idx[name]], 1.0)


best = None
for r in rows[2:]:
    if len(r) < len(hdr) or kname not in r[idx["Kernel Name"]]:
        continue
    e = {"kernel": r[idx["Kernel Name"]].split("(")[0],
         "dram_bytes_read": f(r, "dram__bytes_read.sum", scale), "dram_bytes_write": f(r, "dram__bytes_write.sum", scale),
         "duration_us_under_ncu": f(r, "gpu__time_duration.sum", tscale),
         "tensor_pipe_active_pct": float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]].replace(",", "")),
         "grid": int(float(r[idx["launch__grid_size"]].replace(",", "")))}
    if best is None or e["duration_us_under_ncu"] > best["duration_us_under_ncu"]:
        best = e
assert best, f"no launch of {kname} in {rep}"
best["dram_bytes_per_launch"] = best["dram_bytes_read"] + best["dram_bytes_write"]
best["capture"] = os.path.basename(rep)
best["git"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=root).stdout.strip()
best["how"] = "ncu --set full --clock-control none, largest launch of the kernel in one network step (tools/ncu_traffic.py)"
path = os.path.join(root, "profiles", "r2_ncu_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[workload] = best
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(best, indent=1))
