"""Summaries of Nsight Compute output for profiles/ (run where ncu is installed, no GPU needed).

    python profiles/summarize_ncu.py launches gpurun_out/launches_r1_n200.csv
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv; python profiles/summarize_ncu.py full raw.csv
"""
import collections
import csv
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H = rows[hdr]
    d = collections.defaultdict(list)
    dram = collections.defaultdict(float)  # bytes read + written, if the pass collected them
    for r in rows[hdr + 1:]:
        rec = dict(zip(H, r))
        name = rec["Kernel Name"].split("(")[0]
        val = float(rec["Metric Value"].replace(",", "") or 0)
        if rec.get("Metric Name") == "gpu__time_duration.sum":
            d[name].append(val / 1e3)
        elif rec.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(rec.get("Metric Unit", "byte"), 1.0)
            dram[name] += val * scale
    tot = sum(sum(v) for v in d.values())
    print(f"{'kernel':24s} {'launches':>8s} {'avg us':>10s} {'share':>7s} {'DRAM MB/launch':>15s}   (ncu-serialised, cold cache: compare shares)")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        mb = f"{dram[k] / len(v) / 1e6:15.1f}" if k in dram else f"{'-':>15s}"
        print(f"{k:24s} {len(v):8d} {sum(v) / len(v):10.1f} {sum(v) / tot:7.3f} {mb}")


def full(path):
    rows = list(csv.reader(open(path)))
    H = rows[0]
    cols = [("gpu__time_duration.sum", "t"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
            ("launch__registers_per_thread", "regs"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
            ("lts__t_sector_hit_rate.pct", "L2hit%"), ("smsp__inst_executed.sum", "warp-inst")]
    units = rows[1]
    stall = [(i, h) for i, h in enumerate(H) if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = r[H.index("Kernel Name")].split("(")[0]
        out = []
        for c, label in cols:
            if c in H:
                i = H.index(c)
                out.append(f"{label}={r[i]}{units[i] if label in ('rd', 'wr', 't') else ''}")
        top = sorted(((float(r[i].replace(',', '') or 0), h.split('stalled_')[1].replace('_per_issue_active.ratio', '')) for i, h in stall), reverse=True)[:3]
        print(f"{name:22s} " + " ".join(out))
        print(f"{'':22s} top stalls (warps per issue): " + ", ".join(f"{n} {v:.1f}" for v, n in top))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
