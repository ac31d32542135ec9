"""Summarise an .ncu-rep (read here on the CPU box): per kernel name the launches, duration, DRAM bytes and throughput %,
tensor-pipe %, L2 throughput %, achieved occupancy.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = {
    "dur_us": "gpu__time_duration.sum",
    "dram_rd_MB": "dram__bytes_read.sum",
    "dram_wr_MB": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
}
units = rows[1]


def val(r, key):
    name = want[key]
    if name not in idx:
        return float("nan")
    s = r[idx[name]].replace(",", "")
    try:
        v = float(s)
    except ValueError:
        return float("nan")
    u = units[idx[name]]
    if key.startswith("dram_") and key.endswith("MB"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    if key == "dur_us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v


agg = collections.OrderedDict()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[idx["Kernel Name"]].split("(")[0]
    a = agg.setdefault(name, [])
    a.append({k: val(r, k) for k in want})
print(f"# ncu summary of `{rep}` (per kernel: mean over its captured launches; --set full, --clock-control none)\n")
print("| kernel | launches | mean µs | max µs | DRAM rd MB | DRAM wr MB | DRAM % | L2 % | tensor % | SM % | warps % | regs | grid (first) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, ls in agg.items():
    m = lambda k: sum(x[k] for x in ls) / len(ls)
    print(f"| `{name[:60]}` | {len(ls)} | {m('dur_us'):.1f} | {max(x['dur_us'] for x in ls):.1f} | {m('dram_rd_MB'):.1f} | {m('dram_wr_MB'):.1f} | "
          f"{m('dram_pct'):.1f} | {m('l2_pct'):.1f} | {m('tensor_pct'):.1f} | {m('sm_pct'):.1f} | {m('warps_pct'):.1f} | {int(ls[0]['regs'])} | {int(ls[0]['grid'])} |")
