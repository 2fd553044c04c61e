#!/usr/bin/env python
"""Turns the ncu outputs of tools/profile_round.sh (gpurun_out/*_<tag>.*) into the text summaries under profiles/.
usage: python tools/summarise_profiles.py r1d"""
import collections
import csv
import io
import subprocess
import sys

tag = sys.argv[1]
G = "gpurun_out"


def launches():
    rows = [r for r in csv.reader(open(f"{G}/launches_{tag}.csv")) if len(r) > 10 and r[0] != "ID"]
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r[0], dict(name=r[4], grid=r[8]))[r[12]] = float(r[14].replace(",", ""))
    agg = collections.OrderedDict()
    for k in per.values():
        nm = k["name"].split("(")[0].replace("admmb::", "")
        a = agg.setdefault(nm, dict(n=0, us=0.0, mb=0.0))
        a["n"] += 1
        a["us"] += k["gpu__time_duration.sum"] / 1e3
        a["mb"] += (k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]) / 1e6
    tot = sum(a["us"] for a in agg.values())
    out = [f"profiles/{tag}_launches.txt -- ncu launch list of ONE steady-state frame (10 ADMM iterations) of the 998,250-tet cube, after 20 conditioning + 5 warm-up frames",
           "command: tools/profile_round.sh (ADMMB_NO_GRAPH=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size "
           "--clock-control none --csv python bench.py --cube 55 --steps 1 --warmup 5 --no-cpu-baseline --no-pairs --profile-region)",
           f"(cold-cache, serialised per-launch times: compare SHARES with bench.py's phases_ms_per_iteration, not absolutes; raw CSV: {tag}_launches_ncu.csv)",
           "%-48s %8s %10s %7s %10s %8s" % ("kernel", "launches", "time us", "share", "DRAM MB", "GB/s")]
    for nm, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out.append("%-48s %8d %10.1f %6.1f%% %10.1f %8.1f" % (nm[:48], a["n"], a["us"], 100 * a["us"] / tot, a["mb"], a["mb"] / a["us"] * 1e3 if a["us"] else 0))
    open(f"profiles/{tag}_launches.txt", "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copy(f"{G}/launches_{tag}.csv", f"profiles/{tag}_launches_ncu.csv")
    print("\n".join(out))


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    return hdr, rows[2:]


LOCAL_METRICS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
                 "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                 "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                 "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                 "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
                 "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
                 "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def local():
    hdr, rows = raw(f"{G}/prof_local_{tag}.ncu-rep")
    units = None
    r = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    out = [f"profiles/{tag}_local.txt -- ncu --set full --clock-control none --import-source on, one launch of the hyperelastic local-step kernel, 998,250-tet cube,",
           "frame 26 (20 conditioning + 5 warm-up frames: the regime bench.py times), 5th ADMM iteration of the frame; tools/profile_round.sh + tools/summarise_profiles.py", ""]
    for m in LOCAL_METRICS + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]:
        if m in ix:
            out.append("%-95s %s" % (m, r[ix[m]]))
    open(f"profiles/{tag}_local.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:25]))


def solve():
    hdr, rows = raw(f"{G}/prof_solve_{tag}.ncu-rep")
    ix = {h: i for i, h in enumerate(hdr)}
    f = lambda r, m: float(r[ix[m]].replace(",", "")) if m in ix and r[ix[m]] not in ("", "n/a") else float("nan")
    out = [f"profiles/{tag}_solve.txt -- ncu --set full --clock-control none: the {len(rows)} per-level launches of ONE direct solve ({len(rows) // 2} levels forward, "
           f"{len(rows) // 2} backward), 998,250-tet cube ({rows[0][ix['Kernel Name']].split('(')[0]}).",
           "Under ncu every launch is serialised and cold; in the production path the launches are nodes of one CUDA graph (bench.py phases_ms_per_iteration.solve).", "",
           "%10s %10s %10s %10s %10s %8s %6s %10s" % ("grid", "us", "MB r+w", "dram %", "warps %", "waves", "regs", "L2 hit %")]
    tus = tmb = 0.0
    for r in rows:
        us = f(r, "gpu__time_duration.sum")
        unit_scale = 1.0
        mb = (f(r, "dram__bytes_read.sum") + f(r, "dram__bytes_write.sum"))
        out.append("%10d %10.3f %10.3f %10.2f %10.2f %8.2f %6d %10.2f" % (
            f(r, "launch__grid_size"), us, mb, f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), f(r, "launch__waves_per_multiprocessor"),
            f(r, "launch__registers_per_thread"), f(r, "lts__t_sector_hit_rate.pct")))
        tus += us
        tmb += mb
    out += ["", f"sum over the {len(rows)} launches: {tus:.1f} us (serialised, cold), DRAM traffic {tmb:.1f} MB read + written"]
    open(f"profiles/{tag}_solve.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    import os
    launches()
    local()
    if os.path.exists(f"{G}/prof_solve_{tag}.ncu-rep"):
        solve()
