"""Turns gpurun_out/*.ncu-rep + launches csv into the tracked summaries under profiles/.
Usage: python profiles/summarize_ncu.py <round-tag> <launches.csv> <rep> [<rep> ...]"""
import csv
import json
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
SHORT = {"render_fwd_kernel": "render_fwd", "render_bwd_kernel": "render_bwd", "preprocess_fwd_kernel": "preprocess_fwd",
         "preprocess_bwd_kernel": "preprocess_bwd", "radix_onesweep_kernel": "sort_onesweep_pass",
         "radix_histogram_kernel": "sort_histogram", "duplicate_with_keys_kernel": "duplicate_with_keys",
         "identify_tile_ranges_kernel": "identify_tile_ranges", "radix_scan_hist_kernel": "sort_scan_hist",
         "tile_count_kernel": "tile_count", "tile_scan_kernel": "tile_scan", "tile_scatter_kernel": "tile_scatter",
         "tile_sort_warp_kernel": "tile_sort", "tile_sort_small_kernel": "tile_sort_mid",
         "tile_sort_large_kernel": "tile_sort_large"}


def short(name):
    for k, v in SHORT.items():
        if k in name:
            return v
    return None


def main():
    tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    here = os.path.dirname(os.path.abspath(__file__))
    out = [f"# ncu summary, round {tag} (headline workload: 1 M surfels, 1920x1080, fwd+bwd)\n"]
    traffic = {}
    stats = {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr = rows[0]
        out.append(f"\n## {os.path.basename(rep)}  (`ncu --set full --clock-control none --import-source on`)\n")
        out.append("| kernel | " + " | ".join(w.split("__")[-1].replace(".pct_of_peak_sustained_", " %") for w in WANT) + " |")
        out.append("|---|" + "---|" * len(WANT))
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            vals = [r[hdr.index(w)] if w in hdr else "" for w in WANT]
            out.append(f"| {name.split('(')[0][-40:]} | " + " | ".join(vals) + " |")
            s = short(name)
            if s:
                def val(w):
                    return float(r[hdr.index(w)]) if w in hdr and r[hdr.index(w)] not in ("", "n/a") else None
                stats[s] = {"issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                            "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
                            "warp_instructions": val("smsp__inst_executed.sum"),
                            "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                            "registers_per_thread": val("launch__registers_per_thread"),
                            "source": os.path.basename(rep)}
            if s and "dram__bytes_read.sum" in hdr:
                unit = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
                rd = float(r[hdr.index("dram__bytes_read.sum")]) * unit[rows[1][hdr.index("dram__bytes_read.sum")]]
                wr = float(r[hdr.index("dram__bytes_write.sum")]) * unit[rows[1][hdr.index("dram__bytes_write.sum")]]
                traffic[s] = rd + wr
    # launch list: share of a step per kernel
    per = {}
    with open(launches) as f:
        rows = [r for r in csv.reader(f) if len(r) > 14 and r[0].isdigit()]
    for r in rows:
        s = short(r[4])
        if s:
            per.setdefault(s, []).append(float(r[14]) / 1e3)
    out.append("\n## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES)\n")
    out.append("| kernel | launches captured | mean us / launch | share of our kernels |")
    out.append("|---|---|---|---|")
    tot = sum(sum(v) for v in per.values())
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / tot * 100:.1f} % |")
    open(os.path.join(here, f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
    tpath = os.path.join(here, "ncu_traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    old.update(traffic)
    json.dump(old, open(tpath, "w"), indent=1, sort_keys=True)
    spath = os.path.join(here, "ncu_kernel_stats.json")
    olds = json.load(open(spath)) if os.path.exists(spath) else {}
    olds.update(stats)
    json.dump(olds, open(spath, "w"), indent=1, sort_keys=True)
    print("\n".join(out[-12:]))


if __name__ == "__main__":
    main()
