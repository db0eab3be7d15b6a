#!/usr/bin/env python3
"""tools/ncu_summaries.py -- CPU side of tools/gpu_r02_profile.sh: reads gpurun_out/r02_prof_*.ncu-rep with `ncu -i ... --page raw
--csv` and writes the committed summaries under profiles/ (one text file per capture, selected metrics per kernel), the
launch summary of the bench run, and profiles/matvec_traffic_256x256x256_np1.json (what bench.py prints as roofline.traffic)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def summarize(name, title):
    rep = os.path.join(G, name + ".ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep); return None
    ks, units = raw(rep)
    lines = [title, "ncu --set full --clock-control none (cold caches, serialised launches); one block per captured launch", ""]
    for d in ks:
        lines.append("--- %s" % d["Kernel Name"][:110])
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                lines.append("  %-88s %s %s" % (k, d[k], units.get(k, "")))
    open(os.path.join(P, name.replace("r02_prof_", "r02_ncu_") + "_summary.txt"), "w").write("\n".join(lines) + "\n")
    return ks, units


def fnum(x):
    return float(x.replace(",", ""))


def main():
    r = summarize("r02_prof_matvec", "k_matvec<16,1,2,MV_ALL>, 256^3 on one B200 (tools/run_cg.py --nx 256 --iters 6, 4th SpMV launch)")
    if r:
        ks, units = r
        d = ks[0]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = fnum(d["dram__bytes_read.sum"]) * scale[units["dram__bytes_read.sum"]]
        wr = fnum(d["dram__bytes_write.sum"]) * scale[units["dram__bytes_write.sum"]]
        n, nnz = 257 ** 3, (3 * 257 - 2) ** 3
        json.dump({"kernel": "k_matvec<16,1,2,MV_ALL>", "config": "256x256x256 np=1", "source": "profiles/r02_ncu_matvec_summary.txt (ncu --set full --clock-control none, tools/gpu_r02_profile.sh)",
                   "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
                   "algorithmic_bytes_per_launch": 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n}, open(os.path.join(P, "matvec_traffic_256x256x256_np1.json"), "w"), indent=1)
    summarize("r02_prof_vec", "dot / vector kernels of one CG iteration, 256^3 on one B200 (tools/run_cg.py --nx 256 --iters 8)")
    summarize("r02_prof_setup", "setup kernels, 256^3 on one B200 (tools/asm_probe.py --nx 256 --reps 1, assembly_mode 0 then 1)")
    # launch list of the bench run
    fn = os.path.join(G, "r02_launches_bench_256.csv")
    if os.path.exists(fn):
        rows = list(csv.reader(open(fn, errors="ignore")))
        hdr, agg, order = None, collections.defaultdict(list), []
        for r in rows:
            if "Kernel Name" in r:
                hdr = r; continue
            if hdr and len(r) == len(hdr):
                d = dict(zip(hdr, r))
                if d.get("Metric Name") == "gpu__time_duration.sum":
                    v = fnum(d["Metric Value"]); u = d["Metric Unit"]
                    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
                    agg[d["Kernel Name"][:70]].append(v)
        tot = sum(sum(v) for v in agg.values())
        lines = ["launch list of `python bench.py --steps 1 --warmup 1` (256^3, one B200) under ncu --metrics gpu__time_duration.sum --clock-control none -c 330",
                 "(cold caches, serialised launches: compare SHARES)", "", "%-72s %5s %10s %10s %10s %7s" % ("kernel", "n", "avg us", "min us", "max us", "share")]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append("%-72s %5d %10.2f %10.2f %10.2f %6.1f%%" % (k, len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))
        open(os.path.join(P, "r02_launch_summary.txt"), "w").write("\n".join(lines) + "\n")
        print("\n".join(lines))


if __name__ == "__main__":
    main()
