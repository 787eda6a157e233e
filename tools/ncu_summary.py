"""Turns gpurun_out ncu artefacts into the text summaries committed under profiles/.
Usage: python tools/ncu_summary.py <launches.csv> <prof_a.ncu-rep> [<prof_b.ncu-rep> ...]"""
import collections
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    # The resident worker cannot be profiled launch by launch: ncu serialises kernels, so the stream-ordered
    # doorbell cannot arrive while an instance is being measured and every instance runs to its lifetime cap.
    artefact = {k: v for k, v in agg.items() if k.startswith("worker_resident_kernel")}
    for k in artefact:
        agg.pop(k)
    tot = sum(sum(v) for v in agg.values())
    print(f"# launch list: {path}  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)")
    for k, v in artefact.items():
        print(f"# excluded from the shares: {k} x{len(v)}, {sum(v) / len(v) / 1e3:.0f} us each = its lifetime cap (profiling artefact: "
              "under ncu the doorbell cannot arrive while the instance is being measured)")
    for k, v in agg.items():
        print(f"{k:34s} launches={len(v):4d} total_us={sum(v) / 1e3:10.1f} avg_us={sum(v) / len(v) / 1e3:8.1f} "
              f"share={100 * sum(v) / tot:5.1f}% min_us={min(v) / 1e3:.1f} max_us={max(v) / 1e3:.1f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"\n# full capture: {path}  (ncu --set full --clock-control none --import-source on)")
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:80s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    launches(sys.argv[1])
    for p in sys.argv[2:]:
        full(p)
