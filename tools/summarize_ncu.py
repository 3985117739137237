"""Turn gpurun_out/*.ncu-rep / launch CSVs into the text summaries committed under profiles/.
    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/out.txt "header line"
    python tools/summarize_ncu.py raw gpurun_out/x.ncu-rep profiles/out.txt "header line"
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__cycles_active.avg"]


def launches(src, dst, header):
    rows = list(csv.reader(open(src, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[kn].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [header, "(cold-cache, serialised per-launch times: compare SHARES, not absolutes; unit ns)",
             f"launches={sum(a[0] for a in agg.values())} total_ns={tot:.0f}"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:92]:92s} launches={c:5d} total_ns={t:12.0f} avg_ns={t / c:10.1f} share={t / tot:6.3f}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:8]))


def raw(src, dst, header):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [header]
    for r in rows[2:]:
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k} = {r[i]} {units[i]}")
        lines.append("---")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
