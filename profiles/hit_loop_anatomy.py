"""Hit-loop anatomy of the two render kernels from an ncu report captured with --import-source on:
groups the SASS instructions of each kernel by how often a warp executes them (the execution count identifies
the loop an instruction lives in) and prints, per group, the instruction count, its share of the kernel, the
average number of active lanes and the share of the stall samples.
Usage: python profiles/hit_loop_anatomy.py <report.ncu-rep> [warps_per_kernel=65280] > profiles/r2_hit_loop.md"""
import csv
import subprocess
import sys
from collections import defaultdict


def page(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            return r, rows[i + 1:]
    return None, []


def main():
    rep = sys.argv[1]
    warps = float(sys.argv[2]) if len(sys.argv) > 2 else 65280.0
    print(f"# Hit-loop anatomy of the render kernels (ncu source page of {rep})\n")
    print(f"Counts are warp-level instructions executed per warp ({int(warps)} warps = tiles x 8), headline workload.\n")
    for kernel in ("render_fwd_kernel", "render_bwd_kernel"):
        hdr, data = page(rep, kernel)
        if not hdr:
            print(f"## {kernel}: not in the report\n")
            continue
        ix = {h: i for i, h in enumerate(hdr)}
        groups = defaultdict(lambda: [0, 0.0, 0.0, 0, 0.0])     # n_instr, instr_per_warp, lanes*instr, samples, thread_instr
        total_i, total_s = 0.0, 0
        for r in data:
            try:
                ie = float(r[ix["Instructions Executed"]]); te = float(r[ix["Thread Instructions Executed"]])
                sm = int(r[ix["# Samples"]])
            except (ValueError, IndexError):
                continue
            per_warp = ie / warps
            key = round(per_warp, 1)
            g = groups[key]
            g[0] += 1; g[1] += per_warp; g[3] += sm; g[4] += te
            g[2] += ie
            total_i += ie; total_s += sm
        print(f"## {kernel.replace('_kernel', '')}: {total_i / warps:.0f} instructions per warp, {total_i / 1e6:.0f} M in total\n")
        print("| executed per warp | SASS instructions at that count | instructions per warp | share | avg active lanes | stall samples |")
        print("|---|---|---|---|---|---|")
        top = sorted(groups.items(), key=lambda kv: -kv[1][1])[:10]
        for key, g in top:
            lanes = g[4] / g[2] if g[2] else 0.0
            print(f"| {key} | {g[0]} | {g[1]:.0f} | {100 * g[1] * warps / total_i:.1f} % | {lanes:.1f} | {100 * g[3] / max(total_s, 1):.1f} % |")
        print()
        # where the warps wait: stall reasons summed over the kernel's instructions (all samples)
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: 0 for h in reasons}
        for r in data:
            for h in reasons:
                try:
                    tot[h] += int(r[ix[h]])
                except (ValueError, IndexError):
                    pass
        allr = sum(tot.values()) or 1
        top = sorted(tot.items(), key=lambda kv: -kv[1])[:8]
        print("Stall samples by reason: " + ", ".join(f"{h[6:]} {100 * v / allr:.1f} %" for h, v in top) + "\n")


if __name__ == "__main__":
    main()
