"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm (`--impl reference`: the UNMODIFIED
reference env from baseline/_ref or /root/reference on the host cores, the oracle port only when neither exists)
prints ONE JSON line with the contract's keys on rank 0 and
nothing on the other ranks of a torchrun launch; the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, **env):
    e = dict(os.environ, **{k: str(v) for k, v in env.items()})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, env=e, timeout=300, cwd=ROOT)


def test_reference_arm_line_and_ranks():
    res = run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--cpu-seconds", "1"],
              RANK=0, WORLD_SIZE=2, LOCAL_RANK=0)
    assert res.returncode == 0, res.stderr[-500:]
    lines = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "env-steps/sec (whole box)" and j["unit"] == "env-steps/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 2 and j["value"] > 0
    assert j["steps"] == 2 and j["warmup"] == 1                  # echoes the driver's --steps / --warmup
    cfg = j["config"]
    assert cfg["workload"] == "config4" and cfg["envs_per_gpu"] == 131072 and cfg["global_envs"] == 2 * 131072
    assert cfg["lattice"] == [13, 13] and len(cfg["planes"]) == 10 and "l2" in cfg
    cb = j["cpu_baseline"]
    import refharness
    want_kind = "reference" if refharness.reference_available() else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    if want_kind == "reference":
        assert cb["port_value"] > 0                              # the port's figure stays beside the reference's
    assert j["e2e"] == {"value": j["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the other ranks of the same launch exit 0 without work and without output
    res = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--cpu-seconds", "1"], RANK=1, WORLD_SIZE=2, LOCAL_RANK=1)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        return                      # (on a GPU box the arm runs; tests/test_gpu_* and the driver cover it)
    res = run(["--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-extra"])
    assert res.returncode != 0 and "CUDA" in (res.stderr + res.stdout)
