#!/usr/bin/env python3
"""dev tool: condense an .ncu-rep (read here with `ncu -i`) into the small text summary committed under profiles/.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "sm__cycles_elapsed.max", "smsp__average_warp_latency_per_inst_issued.ratio"]
STALL = "smsp__average_warps_issue_stalled_"
def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# condensed from {rep} (ncu --set full --clock-control none); per launch"]
    for r in rows[2:]:
        lines.append(f"== {r[4]}  block {r[7]} grid {r[8]}")
        stalls = []
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                lines.append(f"  {h} [{u}] = {v}")
            elif h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(v), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines.append("  stall cycles per issued instruction: " + ", ".join(f"{n}={x:.2f}" for x, n in stalls[:7]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
if __name__ == "__main__":
    main()
