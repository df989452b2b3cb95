#!/usr/bin/env python
"""BASELINE.json configs[4]: env-count sweep 1K-4M envs PER GPU at G GPUs of one box (heterogeneous 3x3-6x6 full-rule
puzzles padded to 13x13 -- the config-3 pool --, traceback, terminal validation), beside the reference CPU env on the
box's host cores.  Every point is one `bench.py --workload config5 --envs-per-gpu N` run (torchrun for G > 1, one rank
per GPU, max-over-ranks timing).  Writes gpurun_out/r2_env_sweep_{G}gpu.json (copied to profiles/).

    python tools/env_sweep.py --gpus 1 [--logn 10 12 14 16 17 18 20 22] [--cpu]      (under gpurun [--gpus G])
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run_point(gpus, n, steps, port):
    args = ["--workload", "config5", "--envs-per-gpu", str(n), "--gpus", str(gpus), "--steps", str(steps), "--warmup", "10",
            "--no-cpu-baseline", "--no-extra"]
    if gpus == 1:
        cmd = [sys.executable, BENCH, *args]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), BENCH, *args]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    if not lines:
        return {"envs_per_gpu": n, "error": (r.stderr or r.stdout)[-400:]}
    j = json.loads(lines[-1])
    return {"envs_per_gpu": n, "global_envs": n * gpus, "env_steps_per_s": j["value"], "ms_per_step": j["ms_per_step"],
            "e2e_env_steps_per_s": j["e2e"]["value"] if j.get("e2e") else None,
            "whole_step_frac": j["roofline"]["whole_step"]["frac"], "kernel": j["roofline"]["kernel"].split(" ")[0],
            "launches_per_step": j["gpu_launches"] // j["steps"], "ms_per_step_by_rank": j["run"]["ms_per_step_by_rank"],
            "sm_mhz": (j.get("clocks") or {}).get("sm_mhz"), "throttle_reasons": (j.get("clocks") or {}).get("reasons")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--logn", type=int, nargs="*", default=list(range(10, 23)))
    ap.add_argument("--cpu", action="store_true", help="also time the reference CPU env on the host cores (bench.py --impl reference)")
    ap.add_argument("--single", default=None, help="1-GPU sweep file to compute scaling efficiency against")
    a = ap.parse_args()
    rows = []
    for k, logn in enumerate(a.logn):
        n = 1 << logn
        steps = 400 if n <= (1 << 16) else 100 if n <= (1 << 20) else 40
        row = run_point(a.gpus, n, steps, 29500 + k)
        rows.append(row)
        print(row, flush=True)
    out = {"workload": "BASELINE configs[4]: config-3 pool (1,024 puzzles, 3x3-6x6 full rules padded to 13x13, traceback=True, "
                       "terminal validation, full rewrite of the reference-dtype observation), weak scaling: envs PER GPU fixed",
           "gpus": a.gpus, "rows": rows}
    single = a.single or os.path.join(ROOT, "profiles", "r2_env_sweep_1gpu.json")
    if a.gpus > 1 and os.path.exists(single):
        one = {r["envs_per_gpu"]: r for r in json.load(open(single))["rows"] if "error" not in r}
        for r in rows:
            if "error" not in r and r["envs_per_gpu"] in one:
                r["efficiency_vs_1gpu"] = round(r["env_steps_per_s"] / (a.gpus * one[r["envs_per_gpu"]]["env_steps_per_s"]), 4)
    if a.cpu:
        r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "config5", "--steps", "3", "--warmup", "1",
                            "--cpu-seconds", "6"], capture_output=True, text=True, cwd=ROOT)
        try:
            out["cpu_baseline"] = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1])["cpu_baseline"]
            for row in rows:
                if "error" not in row:
                    row["vs_cpu"] = round(row["env_steps_per_s"] / out["cpu_baseline"]["value"], 1)
        except Exception:
            out["cpu_baseline"] = {"error": r.stderr[-300:]}
    dst = os.path.join(ROOT, "gpurun_out", f"r2_env_sweep_{a.gpus}gpu.json")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
