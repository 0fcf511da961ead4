"""Summarise .ncu-rep captures (ncu --set full) into markdown: python scripts/ncu_summarize.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        head, units = rows[0], rows[1]
        print(f"## {path.split('/')[-1]}\n")
        for row in rows[2:]:
            print(f"- Kernel Name: {row[head.index('Kernel Name')][:110]}")
            for m in METRICS:
                if m in head:
                    i = head.index(m)
                    print(f"  - {m}: {row[i]} {units[i]}")
            print()


if __name__ == "__main__":
    main()
