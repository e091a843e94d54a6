"""Summarise an `ncu --page raw --csv` export (one kernel) into the handful of metrics the roofline
discussion needs.  usage: python tools/ncu_summary.py gpurun_out/<tag>/prof_<wl>_raw.csv [...]"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr = None
        for i, r in enumerate(rows):
            if r and r[0] == "ID":
                hdr = i
                break
        if hdr is None:
            print(path, "no header")
            continue
        names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
        d = {n: (v, u) for n, u, v in zip(names, units, vals)}
        print("==", path, "|", d.get("Kernel Name", ("?",))[0][:90])
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k][0]:>18s} {d[k][1]}")
        pat = sys.argv  # noqa
        for n in names:
            if ("issue_stalled" in n or "warp_issue_stalled" in n) and n.endswith("_per_warp_active.pct"):
                v = d[n][0]
                try:
                    if float(v.replace(",", "")) >= 3.0:
                        print(f"  {n:75s} {v:>18s}")
                except ValueError:
                    pass


if __name__ == "__main__":
    main()
