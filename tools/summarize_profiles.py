"""Turn the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep>[,<more.ncu-rep>] <workload-name> <points-per-launch> [kernels,to,drop]
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"void |at::native::|<unnamed>::|\(anonymous namespace\)::", "", name)[:72]


def launches(path, out_md):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["Metric Unit"], 1e-6)
        a = agg[short(r["Kernel Name"])]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["| ms | share | launches | kernel |", "|---:|---:|---:|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.0002:
            continue
        lines.append(f"| {v[1]:.3f} | {100 * v[1] / tot:.2f}% | {v[0]} | `{k}` |")
    lines.append(f"\ntotal {tot:.2f} ms over {len(rows)} launches (cold-cache, serialised by ncu: compare shares, not absolutes)")
    out_md.write("\n".join(lines) + "\n")
    return agg, tot


def full(reps, out_md, drop=()):
    """reps: comma-separated .ncu-rep files; kernels of later files replace same-named ones of earlier files."""
    res = {}
    for rep in reps.split(","):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ik = hdr.index("Kernel Name")
        seen = set()
        for r in rows[2:]:
            name = short(r[ik])
            if name in seen:
                continue
            seen.add(name)
            d = {}
            for k in KEYS:
                if k in hdr:
                    d[k] = (r[hdr.index(k)], units[hdr.index(k)])
            res[name] = d
    for name in drop:
        res.pop(name, None)
    for name, d in res.items():
        out_md.write(f"\n### `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k, (v, u) in d.items():
            if "nan" in v:
                continue                      # ncu could not collect this metric for this kernel (replay mismatch)
            out_md.write(f"| {k} | {v} | {u} |\n")
    return res


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def main():
    tag, lcsv, rep, workload, ppl = sys.argv[1:6]
    drop = sys.argv[6].split(",") if len(sys.argv) > 6 else ()
    with open(f"profiles/{tag}_launches.md", "w") as f:
        f.write(f"# {tag}: launch list of one timed bench.py step ({workload})\n\n"
                "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu`\n\n")
        launches(lcsv, f)
    with open(f"profiles/{tag}_ncu_full.md", "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the main kernels ({workload})\n")
        res = full(rep, f, drop)
    tr = {}
    for name, d in res.items():
        if "dram__bytes_read.sum" in d and "nan" not in d["dram__bytes_read.sum"][0]:
            tr[name.split("<")[0]] = {"workload": workload, "points_per_launch": int(ppl), "kernel": name,
                                      "dram_bytes_per_launch": to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])}
    json.dump(tr, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
    print("wrote profiles for", list(res))


if __name__ == "__main__":
    main()
