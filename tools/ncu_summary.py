"""Selected metrics of every kernel in an .ncu-rep as text (`metric  value  unit`), for profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep "header comment" > profiles/rNN_x_ncu_full.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "nvlrx__bytes.sum", "nvltx__bytes.sum", "nvlrx__bytes_data_user.sum", "nvltx__bytes_data_user.sum",
        "pcie__read_bytes.sum", "pcie__write_bytes.sum"]


def main():
    rep = sys.argv[1]
    print("# " + (sys.argv[2] if len(sys.argv) > 2 else rep))
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("\n== %s  grid %s block %s" % (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-86s %s %s" % (w, r[i].replace(",", ""), units[i]))
        nv = [h for h in hdr if h.startswith("nvl") and h not in WANT]
        for h in nv:
            i = hdr.index(h)
            if r[i] not in ("", "0", "n/a"):
                print("%-86s %s %s" % (h, r[i].replace(",", ""), units[i]))


if __name__ == "__main__":
    main()
