"""Key metrics of an `ncu --set full` report as JSON (read here, on the GPU-less box):  python tools/ncu_summary.py <report.ncu-rep> <out.json>"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[head.index("Kernel Name")]}
        for w in WANT:
            if w in head:
                i = head.index(w)
                try:
                    d[w] = float(r[i].replace(",", ""))
                except ValueError:
                    d[w] = r[i]
                d[w + " [unit]"] = units[i]
        if "dram__bytes_read.sum" in d:
            mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            d["dram_bytes_total"] = d["dram__bytes_read.sum"] * mult.get(d["dram__bytes_read.sum [unit]"], 1.0) + \
                d["dram__bytes_write.sum"] * mult.get(d["dram__bytes_write.sum [unit]"], 1.0)
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(out, [(x["kernel"][:50], x.get("gpu__time_duration.sum")) for x in res])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
