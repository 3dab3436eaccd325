"""Turn an .ncu-rep (gpurun_out/) into the short tab-separated summary kept under profiles/:
    python profiles/_ncu_summary.py gpurun_out/X.ncu-rep profiles/X_ncu_raw.txt "# header line" ...
Keeps the launch geometry, DRAM / L2 / L1 traffic, pipe utilisation, issue and stall metrics of every captured kernel."""
import csv
import subprocess
import sys

KEEP = ("dram__bytes", "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "sm__throughput.avg.pct", "dram__throughput.avg.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_tma_ld.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled", "smsp__thread_inst_executed_per_inst_executed.ratio")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode().splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for line in sys.argv[3:]:
            f.write(line + "\n")
        for r in rows[2:]:
            f.write("Kernel Name\t%s\t\n" % r[hdr.index("Kernel Name")])
            for i, k in enumerate(hdr):
                if k.startswith(KEEP) and "per_issue_active" in k or (k.startswith(KEEP) and "issue_stalled" not in k):
                    if ".max." in k or ".min." in k or k.endswith(".per_second") and "dram" not in k:
                        continue
                    f.write("%s\t%s\t%s\n" % (k, r[i], units[i]))


if __name__ == "__main__":
    main()
