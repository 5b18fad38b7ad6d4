"""Summarise an ncu report of the sweep kernel into profiles/<tag>.json (+ the per-function instruction table).
usage: summarize_ncu.py <report.ncu-rep> <tag> <n_windows> <window_len> [nvdisasm -g listing]"""
import csv, json, subprocess, sys, io, os
rep, tag, nwin, wlen = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h, u, v = r[0], r[1], r[2]
def g(k):
    return float(v[h.index(k)])
bases = nwin * wlen
keys = {
    "duration_ms": "gpu__time_duration.sum", "warp_instructions": "smsp__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "cycles_per_issued_instruction": "smsp__average_warp_latency_per_inst_issued.ratio",
    "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_branch_resolving": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "dram_read_GB": "dram__bytes_read.sum", "dram_write_GB": "dram__bytes_write.sum",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "registers_per_thread": "launch__registers_per_thread", "active_lanes_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "warps_active_per_smsp": "smsp__warps_active.avg.per_cycle_active",
}
out = {"report": os.path.basename(rep), "kernel": "k_sweep", "windows": nwin, "window_len": wlen}
for k, m in keys.items():
    try: out[k] = g(m)
    except ValueError: out[k] = None
ur, uw = u[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_write.sum")]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
dram = out["dram_read_GB"] * scale[ur] + out["dram_write_GB"] * scale[uw]
out["dram_read_GB"] = out["dram_read_GB"] * scale[ur] / 1e9; out["dram_write_GB"] = out["dram_write_GB"] * scale[uw] / 1e9
out["dram_bytes_per_base"] = dram / bases
out["warp_instructions_per_base"] = out["warp_instructions"] / bases
out["Mbp_per_s_under_ncu_clock"] = bases / 1e6 / (out["duration_ms"] / 1e3) if out["duration_ms"] else None
json.dump(out, open("profiles/%s.json" % tag, "w"), indent=1)
print(json.dumps(out, indent=1))
