"""Summarise ncu output brought back from the GPU box into small tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv  > profiles/rN_launches_X.md
    python tools/ncu_summary.py full     gpurun_out/full_X.ncu-rep  > profiles/rN_full_X.md

`launches` takes the CSV of `ncu --metrics gpu__time_duration.sum --csv` (one row per launch) and prints
the per-kernel launch count, total time and share of the captured region.  `full` reads a `--set full`
report through `ncu -i ... --page raw --csv` (works without a GPU) and prints the metrics the roofline
quotes: DRAM bytes, tensor-pipe activity, occupancy, registers, duration.
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"source: {path}  ({sum(a[0] for a in agg.values())} launches, {tot / 1e3:.1f} us total device time; per-launch times are cold-cache and serialised -- compare shares)\n")
    print("| kernel | launches | total us | avg us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:90]}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / 1e3 / a[0]:.2f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    print(f"source: {path} (ncu --set full --clock-control none; cold-cache replay numbers)\n")
    for r in rows[2:]:
        d = dict(zip(h, zip(u, r)))
        print(f"### {d['Kernel Name'][1][:140]}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in KEEP:
            if k in d:
                print(f"| {k} | {d[k][1]} | {d[k][0]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
