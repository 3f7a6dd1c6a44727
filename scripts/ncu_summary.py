"""Extract a compact per-kernel summary (CSV) from an .ncu-rep:
  python scripts/ncu_summary.py rep out.csv [traffic.json [B,C,H,W]]"""
import csv, subprocess, sys
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    keys = [k for k in KEYS if k in idx]
    w.writerow(keys)
    w.writerow([units[idx[k]] for k in keys])
    for r in rows[2:]:
        w.writerow([r[idx[k]][:100] for k in keys])
print("wrote", sys.argv[2])

# optional third argument: JSON with per-kernel DRAM traffic in bytes per launch (read by bench.py -> roofline.traffic)
if len(sys.argv) > 3:
    import json, re
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    acc = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).strip()
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_write.sum"]]]
        ms = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * tscale[units[idx["gpu__time_duration.sum"]]]
        a = acc.setdefault(name, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "ncu_ms": 0.0})
        a["launches"] += 1
        a["dram_read_bytes"] += rd
        a["dram_write_bytes"] += wr
        a["ncu_ms"] += ms
    for a in acc.values():
        n = a.pop("launches")
        for k in list(a):
            a[k] = a[k] / n
        a["launches_averaged"] = n
    wl = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else None  # B,C,H,W of the captured run
    json.dump({"workload": wl, "kernels": acc}, open(sys.argv[3], "w"), indent=1, sort_keys=True)
    print("wrote", sys.argv[3])
