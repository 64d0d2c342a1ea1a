#!/usr/bin/env python
"""Summarise ncu CSV exports into small text tables (committed under profiles/).
  launches: ncu_summary.py launches <launches.csv> [skip_launches]   -> per-kernel count / total / mean / share
  raw:      ncu_summary.py raw <prof_raw.csv[.gz]>                    -> per-kernel key metrics of a --set full capture
"""
import csv, gzip, sys, collections, io


def opn(p):
    return io.TextIOWrapper(gzip.open(p), newline="") if p.endswith(".gz") else open(p, newline="")


def rows_of(p):
    with opn(p) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def launches(p, skip=0):
    rs = rows_of(p)
    rs = [r for r in rs if r.get("Metric Name") == "gpu__time_duration.sum"]
    rs = rs[skip:]
    agg = collections.OrderedDict()
    for r in rs:
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if unit in ("nsecond", "ns"): v /= 1e3
        elif unit in ("msecond", "ms"): v *= 1e3
        elif unit in ("second", "s"): v *= 1e6
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print("%-44s %7s %12s %10s %10s %7s" % ("kernel", "count", "total_us", "mean_us", "max_us", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %7d %12.1f %10.2f %10.2f %6.1f%%" % (k[:44], a[0], a[1], a[1] / a[0], a[2], 100 * a[1] / tot))
    print("%-44s %7d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def raw(p):
    rs = rows_of(p)
    if not rs:
        print("empty"); return
    cols = rs[0].keys()
    units = rs[0]
    data = rs[1:] if rs[0].get("ID", "") == "" else rs
    seen = collections.OrderedDict()
    for r in data:
        name = r.get("Kernel Name", "?").split("(")[0]
        seen.setdefault(name, []).append(r)
    for name, lst in seen.items():
        print("== %s  (%d captured launches; first / last shown)" % (name, len(lst)))
        for k in KEYS:
            if k in cols:
                vals = [x[k] for x in (lst[0], lst[-1])]
                print("   %-78s %-12s %s" % (k, units.get(k, ""), " | ".join(vals)))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        raw(sys.argv[2])
