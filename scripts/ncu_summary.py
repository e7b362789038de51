#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel per report row) into a small text table for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
    "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        print("kernel:", rec.get("Kernel Name"), "| grid", rec.get("Grid Size"), "| block", rec.get("Block Size"))
        for h, u in zip(hdr, units):
            short = h.split(".TriageCompute.")[-1]
            if short in KEYS:
                print(f"  {short:95s} {rec[h]:>18s} {u}")
        stalls = sorted(((float(rec[h] or 0), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h in hdr
                         if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")), reverse=True)
        print("  warp stall reasons (warps stalled per issue-active cycle):")
        for v, nm in stalls[:9]:
            print(f"    {nm:30s} {v:8.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
