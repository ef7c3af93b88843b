"""Print the key metrics of an .ncu-rep (read with `ncu -i ... --page raw --csv`) as a compact summary."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum ", "lts__t_sectors_srcunit_tex_op_read.sum ", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ", "launch__registers_per_thread ",
    "launch__grid_size", "launch__block_size", "smsp__average_warps_issue_stalled", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed_pipe_", "sm__inst_executed_pipe_", "smsp__pcsamp_warps_issue_stalled",
    "local_", "Kernel Name",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        for h, u, v in zip(hdr, units, r):
            hh = h + " "
            if any(k in hh for k in KEYS + extra):
                try:
                    if float(v.replace(",", "")) == 0.0:
                        continue
                except ValueError:
                    pass
                print(f"{h:100s} {u:14s} {v}")


if __name__ == "__main__":
    main()
