"""Summarise an .ncu-rep (read here, without a GPU): selected raw metrics per captured launch."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_subpipe", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_bytes.sum ", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum ",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg ", "sm__cycles_active.avg "]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print("kernels:", [r[ki][:60] for r in data])
    for i, h in enumerate(hdr):
        if any((h + " ").startswith(w) for w in WANT):
            print(f"{h} [{units[i]}]: {[r[i] for r in data]}")


if __name__ == "__main__":
    main(sys.argv[1])
