"""Turns ncu reports (gpurun_out/*.ncu-rep) into the markdown summaries kept under profiles/.
usage: python tools/ncu_summary.py out.md title rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def main():
    out, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    md = [f"# {title}\n", "Extracted with `ncu -i <report> --page raw --csv` by tools/ncu_summary.py from `ncu --set full --clock-control none "
          "--import-source on` captures (commands in profiles/README.md). Per launch; caches cold and launches serialised by the "
          "profiler, so compare shares and ratios, not absolute times, with bench.py.\n"]
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        md.append(f"\n## {rep.split('/')[-1]}\n")
        import os
        limit = int(os.environ.get("NCU_SUMMARY_LIMIT", "1000"))
        for r in rows[2:2 + limit]:
            name = r[hdr.index("Kernel Name")]
            md.append(f"\n### {name.split('(')[0]}  (launch id {r[hdr.index('ID')]})\n\n| metric | value | unit |\n|---|---|---|")
            for w in WANT:
                if w in hdr:
                    md.append(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")
            top = sorted([(float(r[hdr.index(h)] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                          for h in stall], reverse=True)[:6]
            md.append("| top warp stalls (warps per issue-active cycle) | " + ", ".join(f"{n} {v:.2f}" for v, n in top) + " | |")
    open(out, "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
