"""Summarise `ncu --page raw --csv` exports: one JSON object per profiled launch with the counters the roofline claims rest on.
    python scripts/ncu_summary.py gpurun_out/r2_pair_full_raw.csv [more.csv ...] > profiles/r2_ncu_summary.json"""
import csv, json, sys

WANT = {
    "gpu__time_duration.sum": "gpu_time",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_cycles_pct_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__cluster_size": "cluster",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_hz",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_pct",
    "smsp__average_warp_latency_issue_stalled_barrier_per_warp_active.pct": "stall_barrier_pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar_ratio",
}

out = []
for path in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    if len(rows) < 3:
        continue
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    for r in rows[2:]:
        if len(r) != len(head):
            continue
        d = {"file": path.split("/")[-1], "kernel": r[col["Kernel Name"]][:120]}
        for k, name in WANT.items():
            if k in col and r[col[k]] != "":
                try:
                    v = float(r[col[k]].replace(",", ""))
                except ValueError:
                    continue
                u = units[col[k]]
                if name == "gpu_time":
                    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)      # -> ms
                    name2 = "gpu_time_ms"
                elif u in ("Gbyte",):
                    v, name2 = v * 1e9, name
                elif u in ("Mbyte",):
                    v, name2 = v * 1e6, name
                elif u in ("Kbyte",):
                    v, name2 = v * 1e3, name
                else:
                    name2 = name
                d[name2] = v
        out.append(d)
print(json.dumps(out, indent=1))
