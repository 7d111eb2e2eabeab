"""profiles/traffic.json from the round's own `ncu --set full` exports (scripts/gpu_ncu_r2.sh): for every kernel family that
bench.py reports, the launch with the longest duration, its measured DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum), duration and DRAM utilisation.  bench.py attaches `dram_bytes_per_launch` of the dominant family to
`roofline.traffic`.

    python scripts/make_traffic.py r2_p
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILY = [  # (substring of the kernel name, bench.py family)
    ("gemm_tc_kernel", "linear"), ("edge_message_fwd_planes", "edge_message_fwd_planes"),
    ("edge_message_bwd_source_planes", "edge_message_bwd_planes"), ("edge_message_bwd_target_planes", "edge_message_bwd_planes/target"),
    ("segnorm_slice_apply_kernel<0>", "segnorm_fwd"), ("segnorm_slice_apply_kernel<1>", "segnorm_bwd"),
    ("segnorm_fused_fwd", "segnorm_fwd/cluster"), ("segnorm_fused_bwd", "segnorm_bwd/cluster"),
    ("split_colsum_kernel", "f16_split_colsum"), ("pool_max_fwd", "pool_max_fwd"), ("pool_max_bwd", "pool_max_bwd"),
    ("row_gather_vec", "unpool_fwd"), ("cluster_sum_vec", "unpool_bwd"), ("wplanes_split", "weight_planes_refresh"),
]


def main(tag):
    best = {}
    for part in ("fwd", "bwd"):
        path = os.path.join(ROOT, "gpurun_out", f"prof_{part}_{tag}_raw.csv")
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {name: hdr.index(name) for name in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                                  "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                                  "launch__grid_size")}

        def to_bytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

        def to_us(v, u):
            v = float(v.replace(",", ""))
            return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)

        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            fam = next((f for sub, f in FAMILY if sub in name), None)
            if fam is None:
                continue
            us = to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            if fam not in best or us > best[fam]["us"]:
                best[fam] = {"us": us, "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                             "dram_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                             "kernel": name.split("(")[0], "grid": r[col["launch__grid_size"]],
                             "note": f"ncu --set full, {tag} ({part} capture): the longest launch of this kernel in one cfg2 step; "
                                     f"{rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written in {us:.1f} us"}
    out = os.path.join(ROOT, "profiles", "traffic.json")
    for ent in best.values():
        ent["source"] = f"gpurun_out/prof_*_{tag}_raw.csv via scripts/make_traffic.py"
    json.dump({"cfg2": best}, open(out, "w"), indent=1)
    print(f"wrote {out}: {len(best)} families")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2_p")
