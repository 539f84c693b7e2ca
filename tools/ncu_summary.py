"""Key metrics of an .ncu-rep (ncu --set full) as text: duration, DRAM bytes, tensor/SM/DRAM %, stalls."""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("kernel:", name[:150])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k} = {r[i]} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 5:
                    stalls.append((v, h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")))
        print("  stalls (>5% of warp-active):", ", ".join(f"{n} {v:.0f}%" for v, n in sorted(stalls, reverse=True)))


if __name__ == "__main__":
    main(sys.argv[1])
