#!/usr/bin/env python
"""Summarise an ncu report (gpurun_out/*.ncu-rep) into profiles/<name>.json + .txt:
per-launch duration, DRAM bytes, throughput, occupancy, issue rate, warp-stall reasons."""
import csv, io, json, subprocess, sys

rep, name = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
launches = []
for r in data:
    d = {"kernel": r[col["Kernel Name"]]}
    for w in want:
        if w in col:
            d[w] = f"{r[col[w]]} {units[col[w]]}".strip()
    d["stalls_per_issue"] = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: float(r[col[h]])
                             for h in stalls if r[col[h]] not in ("", "n/a")}
    d["stalls_per_issue"] = dict(sorted(d["stalls_per_issue"].items(), key=lambda kv: -kv[1])[:8])
    launches.append(d)
json.dump({"report": rep, "launches": launches}, open(f"profiles/{name}.json", "w"), indent=1)
with open(f"profiles/{name}.txt", "w") as f:
    for i, d in enumerate(launches):
        f.write(f"== launch {i}: {d['kernel']}\n")
        for k, v in d.items():
            if k not in ("kernel", "stalls_per_issue"):
                f.write(f"  {k:70s} {v}\n")
        f.write("  top warp-stall reasons (warps stalled per issue-active cycle):\n")
        for k, v in d["stalls_per_issue"].items():
            f.write(f"    {k:30s} {v:.3f}\n")
print(open(f"profiles/{name}.txt").read())
