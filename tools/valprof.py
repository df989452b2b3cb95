"""Where one queued terminal state spends k_val's time: cycle counters of validate_env (prologue / flood fill /
exact fit / everything else) per snapshot, mean and maximum, in the delta step where k_val is the critical path.

    bash tools/build_variants.sh "valprof -DVAL_PROFILE"
    SPARC_B200_LIB=$PWD/build/variants/lib_valprof.so python tools/valprof.py [--obs-update delta|rewrite] [--obs-format ...]
    (run under gpurun; clock64 counts SM cycles, 1.965 GHz on the pool's B200s)
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sparc_gym_b200.dataset import config_pool  # noqa: E402
from sparc_gym_b200.vec_env import SparcVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--obs-update", default="delta")
ap.add_argument("--obs-format", default="reference")
ap.add_argument("--envs", type=int, default=131072)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--pool", type=int, default=4096)
a = ap.parse_args()
rows = config_pool(4, n_puzzles=a.pool)
env = SparcVecEnv(rows, a.envs, traceback=False, max_steps=1000, autoreset_mode="same_step", validate="terminal",
                  obs_format=a.obs_format, obs_update=a.obs_update)
env.reset()
acts = torch.randint(0, 4, (200 + a.steps, a.envs), dtype=torch.uint8, device="cuda")
for t in range(200):
    env.step(acts[t])
out = (C.c_uint64 * 8)()
env.lib.sparc_debug_valprof.argtypes = [C.c_void_p, C.c_int]
assert env.lib.sparc_debug_valprof(out, 1) == 0, "build the library with -DVAL_PROFILE and select it with SPARC_B200_LIB"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(200, 200 + a.steps):
    env.step(acts[t])
e1.record()
torch.cuda.synchronize()
env.lib.sparc_debug_valprof(out, 1)
tot, pro, flood, fit, n, mx_tot, mx_flood, mx_fit = (int(v) for v in out)
ghz = 1.965
print(f"{a.obs_update}/{a.obs_format}: {e0.elapsed_time(e1) / a.steps * 1e3:.1f} us per step (launched from Python), "
      f"{n / a.steps:.0f} snapshots per step")
for name, v in (("total", tot), ("prologue", pro), ("flood fill", flood), ("exact fit", fit), ("rules + rest", tot - pro - flood - fit)):
    print(f"  mean {name:12s} {v / max(n, 1):9.0f} cycles = {v / max(n, 1) / ghz / 1e3:6.2f} us")
for name, v in (("total", mx_tot), ("flood fill", mx_flood), ("exact fit", mx_fit)):
    print(f"  max  {name:12s} {v:9d} cycles = {v / ghz / 1e3:6.2f} us   (over all snapshots of the {a.steps} steps)")
