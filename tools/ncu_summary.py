#!/usr/bin/env python3
"""Condense an `ncu --page raw --csv` export into the handful of numbers the profiles/ summaries quote.
Usage: python tools/ncu_summary.py raw.csv [more.csv ...]"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        units = rows[1] if len(rows) > 1 else []
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print("==", path, "::", d.get("Kernel Name", "?")[:80])
            for k in KEYS:
                if k in d:
                    print("  %-75s %s %s" % (k, d[k], u.get(k, "")))
            # every pipe-utilisation percentage of the capture (integer work sits on fma / fmaheavy / alu)
            for k in sorted(d):
                if k not in KEYS and (k.startswith("sm__inst_executed_pipe_") or k.startswith("sm__pipe_")) and "pct_of_peak_sustained_active" in k and d[k] not in ("", "n/a", "0"):
                    print("  %-75s %s %s" % (k, d[k], u.get(k, "")))
            stalls = [(float(v.replace(",", "")), k) for k, v in d.items() if "average_warp" in k and "issue_stalled" in k and k.endswith("_per_warp_active.pct") is False and v not in ("", "n/a")]
            st = []
            for k, v in d.items():
                if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                    try:
                        st.append((float(v.replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            print("  stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in st[:8]))


if __name__ == "__main__":
    main()
