"""Summarises an .ncu-rep (read here, without a GPU) into the text files committed under profiles/.

    python tools/ncu_summary.py gpurun_out/all_r1b.ncu-rep profiles/r01b
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__sass_branch_targets_threads_divergent.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(prefix + "_metrics.txt", "w") as f:
        f.write("# ncu --set full --clock-control none, read with `ncu -i %s --page raw --csv`\n" % rep)
        f.write("# per-launch values are cold-cache and serialised: compare shares, not absolutes\n")
        for r in rows[2:]:
            f.write("\n== launch %s: %s\n" % (r[idx["ID"]], r[idx["Kernel Name"]]))
            for m in METRICS:
                if m in idx:
                    f.write("%-86s %-14s %s\n" % (m, units[idx[m]], r[idx[m]]))
    # opcode mix + hottest instructions of the first launch of each distinct kernel
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        short = re.sub(r"[^a-z_]", "", name.split("airice_")[-1].split("(")[0].split("<")[0])
        key = name
        if key in seen:
            continue
        seen.add(key)
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", r[idx["ID"]], "--launch-count", "1"],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        h2 = None
        for i, sr in enumerate(srows):
            if "Source" in sr and "Instructions Executed" in sr:
                h2 = sr
                body = srows[i + 1:]
                break
        if h2 is None:
            continue
        iS, iSamp, iInst = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
        body = [b for b in body if len(b) > iInst and b[iInst].isdigit()]
        body = body[:len(body) // 2] if len(body) > 1 and body[0][iS] == body[len(body) // 2][iS] else body
        tot_i = sum(int(b[iInst]) for b in body) or 1
        tot_s = sum(int(b[iSamp]) for b in body) or 1
        ops, samp = collections.Counter(), collections.Counter()
        for b in body:
            op = re.sub(r"^@!?U?P\d+\s+", "", b[iS].strip()).split()[0].split(".")[0]
            ops[op] += int(b[iInst]); samp[op] += int(b[iSamp])
        tag = short + ("_" + re.sub(r"[^0-9]", "", name.split("<")[1].split(">")[0]) if "<" in name.split("airice_")[-1] else "")
        with open(prefix + "_%s_sass_mix.txt" % tag, "w") as f:
            f.write("# %s (launch %s)\n# SASS instructions: %d, warp-level instructions executed: %d, stall samples: %d\n" % (
                name, r[idx["ID"]], len(body), tot_i, tot_s))
            f.write("# opcode            share of executed instr   share of stall samples\n")
            for op, c in ops.most_common(24):
                f.write("%-12s %8.2f%% %8.2f%%\n" % (op, 100.0 * c / tot_i, 100.0 * samp[op] / tot_s))
            f.write("\n# 40 instructions with the most stall samples (index, executed/1e6, samples, SASS)\n")
            order = sorted(range(len(body)), key=lambda i: -int(body[i][iSamp]))[:40]
            for i in sorted(order):
                f.write("%5d %10.3f %7d  %s\n" % (i, int(body[i][iInst]) / 1e6, int(body[i][iSamp]), body[i][iS].strip()))


if __name__ == "__main__":
    main()
