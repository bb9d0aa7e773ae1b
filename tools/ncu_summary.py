"""Condense an ncu report (.ncu-rep, `--set full`) into the per-launch table committed under profiles/:
duration, DRAM bytes / %, SM %, tensor-pipe %, warps active, registers, IPC, grid.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv ["first-column title"]"""
import csv
import subprocess
import sys

METRICS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "sm__inst_executed.avg.per_cycle_elapsed", "launch__grid_size", "launch__block_size"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else "launch"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(m) for m in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([title] + [hdr[i] for i in idx])
        w.writerow([""] + [units[i] for i in idx])
        for k, r in enumerate(rows[2:], 1):
            w.writerow([k] + [r[i] for i in idx])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
