"""Turn what a GPU trip left in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md
    python tools/ncu_summary.py full gpurun_out/prof_r1_msda.ncu-rep profiles/r1_ncu_msda.csv

`launches`: per-kernel totals / shares of one train step from the `--metrics gpu__time_duration.sum` pass.
`full`    : one row per profiled launch of a `--set full` capture with the metrics the roofline uses.
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__cycles_active.avg", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum"]


FAMILIES = {"conv_tc": ("tc_fwd_persist", "tc_fwd_ts", "tc_pair_kernel", "tc_pair16_kernel", "tc_fwd_kernel"),
            "conv_wgrad_tc": ("tc_wgrad_kernel",), "msda_fwd": ("msda_fwd_kernel",), "msda_bwd": ("msda_bwd_kernel",),
            "bn": ("bn_apply_kernel", "bn_bwd_reduce_kernel", "bn_bwd_apply_kernel", "bn_stats_kernel", "bn_finalize_kernel"),
            "attention": ("attn_mma::",)}


def traffic(src, dst):
    """DRAM bytes per kernel family of one eager step from an ncu pass with
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum  ->  the json bench.py reads."""
    import json
    lines = open(src).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    per = collections.defaultdict(dict)
    for row in csv.DictReader(lines[start:]):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[row["ID"]][row["Metric Name"]] = v * scale
        per[row["ID"]]["name"] = row["Kernel Name"]
    out = {}
    for fam, keys in FAMILIES.items():
        rows = [r for r in per.values() if any(k in r["name"] for k in keys)]
        if not rows:
            continue
        rd = sum(r.get("dram__bytes_read.sum", 0.0) for r in rows)
        wr = sum(r.get("dram__bytes_write.sum", 0.0) for r in rows)
        out[fam] = {"launches": len(rows), "us": sum(r.get("gpu__time_duration.sum", 0.0) for r in rows),
                    "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": (rd + wr) / len(rows)}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: {"launches": v["launches"], "MB_per_launch": round(v["dram_bytes_per_launch"] / 1e6, 2)} for k, v in out.items()}))


def launches(src, dst):
    lines = open(src).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines[start:]):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        name = name.replace("native::", "")[:72]
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    ours = sum(t for k, (n, t) in agg.items() if not k.startswith(("at::", "vectorized_", "elementwise_", "unrolled_",
                                                                    "reduce_kernel", "index_", "sbtopk", "CatArray",
                                                                    "multi_tensor", "indexing_", "cub::", "bitonic")))
    with open(dst, "w") as f:
        f.write(f"# one eager D-FINE-m train step (batch 16, 640x640) under `ncu --metrics gpu__time_duration.sum "
                f"--clock-control none`\n\n{sum(n for n, _ in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time "
                f"(serialised, cold cache: compare SHARES); library (libdfine_sm100) kernels: {100 * ours / tot:.1f} %\n\n"
                "| kernel | launches | total us | share % | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
            f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.2f} | {t / n:.2f} |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    idx = [(k, head.index(k)) for k in KEYS if k in head]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{k} [{units[i]}]" if units[i] else k for k, i in idx])
        for r in rows[2:]:
            w.writerow([re.sub(r"\(CUtensorMap.*|\(const float.*|\(float.*", "", r[i]) if k == "Kernel Name" else r[i]
                        for k, i in idx])


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
