"""Condense an `ncu --set full` report into the handful of numbers DESIGN.md / bench.py quote.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/prof_rNN.txt

Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and prints, per captured launch, duration, DRAM
bytes read/written, throughput percentages, occupancy limits, instruction counts and pipe utilisation."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("kernel:", name)
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (r[i], units[i])
                print(f"  {k:90s} {r[i]:>16s} {units[i]}")
        try:
            rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            tot = float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]]
            t = float(vals["gpu__time_duration.sum"][0]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[vals["gpu__time_duration.sum"][1]]
            print(f"  {'dram traffic (read+write) per launch':90s} {tot/1e6:16.1f} MB   -> {tot/t/1e9:.0f} GB/s under ncu")
        except Exception as e:  # noqa: BLE001
            print("  (traffic summary unavailable:", e, ")")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
