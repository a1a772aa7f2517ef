"""Prints the handful of ncu raw metrics we track from a .ncu-rep (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-substring]"""
import csv, subprocess, sys, io

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"

def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if filt and filt not in d.get("Kernel Name", ""):
            continue
        print("### kernel:", d.get("Kernel Name"), " id", d.get("ID"))
        u = dict(zip(hdr, units))
        for k in KEYS:
            if k in d:
                print("  %-72s %s %s" % (k, d[k], u[k]))
        st = sorted(((float(d[k] or 0), k) for k in hdr if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")), reverse=True)
        print("  stalls (warps per issue):", ", ".join("%s=%.2f" % (k[len(STALL):-len("_per_issue_active.ratio")], v) for v, k in st[:8]))

main()
