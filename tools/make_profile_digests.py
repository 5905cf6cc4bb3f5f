#!/usr/bin/env python
"""Turn what tools/profile_round.sh left in gpurun_out/ into the tracked summaries under profiles/.

  tools/make_profile_digests.py TAG [OUT_PREFIX]      e.g.  tools/make_profile_digests.py r1f r01
For every gpurun_out/prof_<TAG>_<kernel>.metrics.csv: profiles/<OUT_PREFIX>_ncu_<kernel>.txt (selected metrics of the one
captured launch + hottest source lines).  Also copies the launch list, the bench lines and the pytest tail."""
import csv
import glob
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pre = sys.argv[2] if len(sys.argv) > 2 else "r01"
src, dst = os.path.join(REPO, "gpurun_out"), os.path.join(REPO, "profiles")
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_shared_loads", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed")
for f in sorted(glob.glob(os.path.join(src, "prof_%s_*.metrics.csv" % tag))):
    k = os.path.basename(f)[len("prof_%s_" % tag):-len(".metrics.csv")]
    rows = list(csv.reader(open(f)))
    h, u, v = rows[0], rows[1], rows[2]
    with open(os.path.join(dst, "%s_ncu_%s.txt" % (pre, k)), "w") as o:
        o.write("# ncu --set full --clock-control none --import-source on -k %s -c 1 python bench.py --steps 1 --warmup 1 --no-cpu   (capture %s; tools/profile_kernels.sh)\n" % (k, tag))
        o.write("# selected metrics of the one captured launch\n")
        for name in KEEP:
            if name in h:
                i = h.index(name)
                o.write("%s\t%s\t%s\n" % (name, v[i], u[i]))
        o.write("\n# hottest source lines (share of stall samples / of executed warp instructions), tools/ncu_lines.py\n")
        o.write(open(f[:-len(".metrics.csv")] + ".lines.txt").read())
for a, b in (("%s_launches.csv" % tag, "%s_launches_final.csv" % pre), ("%s_pytest_gpu.txt" % tag, "%s_pytest_gpu.txt" % pre),
             ("%s_bench_1gpu.json" % tag, "%s_bench_1gpu_final.json" % pre), ("%s_bench_reference.json" % tag, "%s_bench_reference_arm.json" % pre)):
    if os.path.exists(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(dst, b))
