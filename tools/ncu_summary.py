"""Summarise ncu outputs into profiles/ (tracked).  usage:
   ncu_summary.py full <prof.ncu-rep> <out.md>      per-launch DRAM bytes / throughput of a --set full capture
   ncu_summary.py launches <launches.csv> <out.md>  kernel shares of a gpu__time_duration launch list"""
import csv
import subprocess
import sys
from collections import defaultdict

mode, src, out = sys.argv[1:4]
if mode == "full":
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{src}` (per launch)\n\n")
        f.write("| kernel | " + " | ".join(c.replace("__", " ").replace(".sum", "") for c in cols) + " | sectors/request |\n")
        f.write("|---|" + "---|" * (len(cols) + 1) + "\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0]
            vals = [f"{r[idx[c]]} {units[idx[c]]}" if c in idx else "n/a" for c in cols]
            try:
                spr = float(r[idx[cols[6]]]) / float(r[idx[cols[7]]])
            except Exception:
                spr = float("nan")
            f.write(f"| `{name}` | " + " | ".join(vals) + f" | {spr:.2f} |\n")
else:
    tot = defaultdict(float)
    cnt = defaultdict(int)
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary of `{src}` (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k in sorted(tot, key=tot.get, reverse=True):
            f.write(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {tot[k]/cnt[k]:.2f} | {100*tot[k]/total:.1f}% |\n")
print(open(out).read())
