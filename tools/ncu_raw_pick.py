"""Print selected metrics of `ncu --page raw --csv` exports.  usage: python tools/ncu_raw_pick.py file.csv [more.csv ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct"]


def main(paths):
    for p in paths:
        rows = list(csv.reader(open(p)))
        if len(rows) < 3:
            print(p, ": no data")
            continue
        hdr, units, val = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
        print("==", p, d.get("Kernel Name", ("", ""))[0][:70])
        for k in KEYS:
            if k in d:
                print("   %-90s %s %s" % (k, d[k][0], d[k][1]))


if __name__ == "__main__":
    main(sys.argv[1:])
