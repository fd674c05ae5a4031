"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the handful of metrics the profile summaries quote.
Usage: python scripts/ncu_summary.py file.raw.csv [...]   (build container; reads only the exported CSV)"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"== {path}: {d.get('Kernel Name', '?')[:60]} grid {d.get('Grid Size')} block {d.get('Block Size')}")
            for h, u in zip(hdr, units):
                if any(k == h or (k in h) for k in KEYS) or "tensor" in h or "stall" in h.lower() and "pct" in h:
                    v = d[h]
                    if v not in ("", "0", "n/a"):
                        print(f"   {h} [{u}] = {v}")


if __name__ == "__main__":
    main()
