#!/usr/bin/env python
"""BASELINE.json configs[4]: env-count sweep 1K-4M on one GPU, heterogeneous puzzle sizes padded to
13x13 (the config-3 pool), traceback, terminal validation.  Writes gpurun_out/r1_env_sweep.json (copied to profiles/)."""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
for logn in range(10, 23):
    n = 1 << logn
    steps = 400 if n <= (1 << 16) else 100 if n <= (1 << 20) else 40
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "config3", "--envs-per-gpu", str(n),
                        "--steps", str(steps), "--warmup", "10", "--no-cpu-baseline", "--no-e2e"], capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        row = {"envs": n, "env_steps_per_s": j["value"], "ms_per_step": j["ms_per_step"],
               "whole_step_frac": j["roofline"].get("whole_step", {}).get("frac", j["roofline"]["frac"]),
               "kernel": j["roofline"]["kernel"].split(" ")[0], "launches_per_step": j["gpu_launches"] // j["steps"]}
    except Exception:
        row = {"envs": n, "error": r.stderr[-300:]}
    out.append(row)
    print(row, flush=True)
json.dump({"workload": "config3 pool (1,024 puzzles, 3x3-6x6 full rules, traceback=True), 1 B200", "rows": out},
          open(os.path.join(root, "gpurun_out", "r1_env_sweep.json"), "w"), indent=1)
