"""Print the handful of ncu metrics the roofline discussion needs from a .ncu-rep (read on the CPU box)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread ", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "smsp__inst_executed.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum ", "lts__t_sector_hit_rate",
        "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct", "sm__cycles_elapsed.avg ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum "]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== kernel:", name[:100])
        for h, u, v in zip(hdr, units, vals):
            hh = h + " "
            if any(w in hh for w in WANT):
                try:
                    if float(v) == 0.0 and "stalled" in h:
                        continue
                except ValueError:
                    pass
                print("  %-95s %-10s %s" % (h, u, v))


if __name__ == "__main__":
    main(sys.argv[1])
