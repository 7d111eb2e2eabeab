"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/r1_X_launches.md
    python scripts/summarize_ncu.py full     gpurun_out/prof_X.ncu-rep|prof_X_raw.csv  profiles/r1_X_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ns = v * (1e3 if unit.startswith("us") else 1e6 if unit.startswith("ms") else 1.0)
        name = row["Kernel Name"]
        key = re.sub(r"\(.*", "", name)
        key = re.sub(r"^void ", "", key)[:90]
        agg[key][0] += 1
        agg[key][1] += ns
        tot += ns
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off` over ONE timed step "
                "(cold-cache, serialised: compare SHARES, not absolutes).\n\n")
        f.write(f"total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches\n\n")
        f.write("| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |\n")


WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %")]


def full(src, dst):
    # src: a .ncu-rep, or the `ncu -i rep --page raw --csv` dump of one made on the GPU box (reps over ~20 MB do not
    # fit gpurun_out/'s 64 MiB cap, so scripts/gpu_round.sh converts there)
    if src.endswith(".csv"):
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), lab) for m, lab in WANT if m in hdr]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n`--clock-control none --import-source on`; one row per captured launch.\n\n")
        f.write("| kernel | " + " | ".join(f"{lab} ({units[i]})" if units[i] else lab for i, lab in cols) + " |\n")
        f.write("|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[ki].split("(")[0]).replace("stinet::", "")[:60]
            vals = []
            for i, _ in cols:
                try:
                    vals.append(f"{float(r[i].replace(',', '')):.1f}")
                except ValueError:
                    vals.append(r[i])
            f.write(f"| `{name}` | " + " | ".join(vals) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
