"""Writes profiles/kernel_facts.json from this round's ncu summaries (profiles/r02_*_metrics.txt, themselves made by
tools/ncu_summary.py from the `ncu --set full` reports): per kernel, DRAM bytes and executed FP64 flop per unit of work,
FP64-pipe and issue-slot activity.  bench.py quotes these next to its live timings (it cannot run under a profiler)."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def launches(path):
    out, cur = [], None
    for line in open(path):
        m = re.match(r"== launch (\d+): (.*)", line)
        if m:
            cur = {"name": m.group(2)}
            out.append(cur)
            continue
        p = line.split()
        if cur is not None and len(p) >= 2 and not line.startswith("#"):
            try:
                val = float(p[-1])
            except ValueError:
                continue
            unit = p[-2] if len(p) >= 3 else ""
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
            cur[p[0]] = val * scale
    return out


def dram(l):
    return l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)


def flops(l, mix_file):
    """executed FP64 flop of a launch: warp instructions x mean active lanes x (2 DFMA + DMUL + DADD shares) from the
    opcode mix tools/ncu_summary.py wrote (the per-opcode thread-instruction counters are not in this ncu's raw page)"""
    share = {}
    warp = 0
    for line in open(os.path.join(P, mix_file)):
        m = re.match(r"# SASS instructions: \d+, warp-level instructions executed: (\d+)", line)
        if m:
            warp = int(m.group(1))
        m = re.match(r"(DFMA|DMUL|DADD)\s+([\d.]+)%", line)
        if m:
            share[m.group(1)] = float(m.group(2)) / 100
    lanes = l.get("smsp__thread_inst_executed_per_inst_executed.ratio", 32.0)
    return warp * lanes * (2 * share.get("DFMA", 0) + share.get("DMUL", 0) + share.get("DADD", 0))


def main():
    facts = {}
    s = launches(os.path.join(P, "r02f_solve1e7_metrics.txt"))
    k1 = [l for l in s if "solve_kernel<1" in l["name"]][0]
    k2 = [l for l in s if "solve_kernel<2" in l["name"]][0]
    n = 1e7
    facts["airice_solve_kernel"] = {
        "dram_bytes_per_unit": (dram(k1) + dram(k2)) / n, "sass_fp64_flop_per_unit": (flops(k1, "r02f_solve1e7_solve_kernel_11_sass_mix.txt") + flops(k2, "r02f_solve1e7_solve_kernel_20_sass_mix.txt")) / n,
        "fp64_pipe_active_pct": k1["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"],
        "issue_active_pct": k1["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "lanes_per_instruction": k1.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "ncu_ms": [k1["gpu__time_duration.sum"], k2["gpu__time_duration.sum"]], "unit": "pair",
        "source": "profiles/r02f_solve1e7_metrics.txt: ncu --set full --clock-control none of tools/ncu_target.py solve 1e7 "
                  "(airice_solve_kernel<1> + <2>, one launch each); algorithmic 89 B/pair"}
    lk = launches(os.path.join(P, "r02g_lookup1e7_metrics.txt"))[0]
    facts["airice_lookup_kernel"] = {
        "dram_bytes_per_unit": dram(lk) / n, "dram_read_bytes_per_unit": lk["dram__bytes_read.sum"] / n,
        "dram_throughput_pct": lk["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
        "l2_hit_pct": lk["lts__t_sector_hit_rate.pct"], "issue_active_pct": lk["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "long_scoreboard_per_issue": lk["smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"],
        "ncu_ms": lk["gpu__time_duration.sum"], "unit": "lookup",
        "source": "profiles/r02g_lookup1e7_metrics.txt: ncu --set full of 1e7 lookups on the reference-grid table; algorithmic 265 B/lookup"}
    al = launches(os.path.join(P, "r02f_all_metrics.txt"))
    tb = [l for l in al if "airice_table_kernel<1, 0>" in l["name"]]
    big = max(tb, key=lambda l: l["gpu__time_duration.sum"])
    cells = 9701 * 900
    facts["airice_table_kernel"] = {
        "dram_bytes_per_unit": dram(big) / cells, "sass_fp64_flop_per_unit": flops(min(tb, key=lambda l: l["gpu__time_duration.sum"]), "r02f_all_table_kernel_10_sass_mix.txt") / (4851 * 177),
        "sass_note": "opcode mix of the README-grid launch (4851 x 177 cells, first launch of that kernel in the report)",
        "fp64_pipe_active_pct": big["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"],
        "issue_active_pct": big["smsp__issue_active.avg.pct_of_peak_sustained_active"], "ncu_ms": big["gpu__time_duration.sum"],
        "unit": "cell", "source": "profiles/r02f_all_metrics.txt: airice_table_kernel<f64> on the reference grid (9701 x 900 cells, 17 f64 "
                                  "columns = 136 B/cell written)"}
    m64 = launches(os.path.join(P, "r02f_multi64_metrics.txt"))
    tm = [l for l in m64 if "table_multi_kernel" in l["name"]][0]
    prep = [l for l in m64 if "row_" in l["name"]]
    facts["airice_table_multi_kernel"] = {
        "dram_bytes_per_unit": dram(tm) / (cells * 64), "dram_throughput_pct": tm["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
        "fp64_pipe_active_pct": tm["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"],
        "issue_active_pct": tm["smsp__issue_active.avg.pct_of_peak_sustained_active"], "ncu_ms": tm["gpu__time_duration.sum"],
        "row_prep_ncu_ms_per_32_tables": sum(l["gpu__time_duration.sum"] for l in prep[:3]),
        "unit": "cell x antenna", "source": "profiles/r02f_multi64_metrics.txt: 64 antennas' reference-grid tables in one pass, lookup layout only "
                                            "(52 B per cell and antenna); row_range + row_lut + row_block kernels of the first 32 tables"}
    for key, pat in (("airice_inice_ladder_kernel", "inice_ladder"), ("airice_inice_dr_kernel", "inice_dr"), ("airice_path_fill_kernel", "path_fill")):
        l = [x for x in al if pat in x["name"]]
        if l:
            l = l[0]
            facts[key] = {"ncu_ms": l["gpu__time_duration.sum"], "fp64_pipe_active_pct": l["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"],
                          "issue_active_pct": l["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                          "lanes_per_instruction": l.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                          "dram_throughput_pct": l["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
                          "source": "profiles/r02f_all_metrics.txt (1e6 pairs / 1024 rays)"}
    json.dump(facts, open(os.path.join(P, "kernel_facts.json"), "w"), indent=1)
    print("wrote profiles/kernel_facts.json:", {k: {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "source"} for k, v in facts.items()})


if __name__ == "__main__":
    main()
