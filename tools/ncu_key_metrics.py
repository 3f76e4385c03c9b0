"""Reduce an `ncu -i X.ncu-rep --page raw --csv` export to the columns the profiles/ summaries quote.
Usage: python tools/ncu_key_metrics.py raw.csv > key_metrics.csv"""
import csv
import sys

KEEP = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_active.avg"]


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    head = rows[start]
    idx = [head.index(k) for k in KEEP if k in head]
    w = csv.writer(sys.stdout)
    for r in rows[start:]:
        if len(r) >= len(head):
            w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()
