"""Where one CTA of the fused small-batch kernel k_env spends its time (config 2: 4,096 envs, tile 4).

    bash tools/build_variants.sh "envprof -DENV_PROFILE"
    SPARC_B200_LIB=$PWD/build/variants/lib_envprof.so python tools/envprof.py [--workload config2] [--envs N] [--validate terminal]
    (run under gpurun; clock64 counts SM cycles, 1.965 GHz on the pool's B200s; globaltimer gives the launch's own span)
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import WORKLOADS  # noqa: E402
from sparc_gym_b200.dataset import config_pool  # noqa: E402
from sparc_gym_b200.vec_env import SparcVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config2")
ap.add_argument("--envs", type=int, default=None)
ap.add_argument("--validate", default="terminal")
ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
cfg, n_env, pool_n, traceback, max_steps, _ = WORKLOADS[a.workload]
n_env = a.envs or n_env
env = SparcVecEnv(config_pool(cfg, n_puzzles=pool_n), n_env, traceback=traceback, max_steps=max_steps, autoreset_mode="same_step",
                  validate=a.validate)
env.reset()
acts = torch.randint(0, 4, (100 + a.steps, n_env), dtype=torch.uint8, device="cuda")
for t in range(100):
    env.step(acts[t])
out = (C.c_uint64 * 16)()
env.lib.sparc_debug_envprof.argtypes = [C.c_void_p, C.c_int]
assert env.lib.sparc_debug_envprof(out, 1) == 0, "build the library with -DENV_PROFILE and select it with SPARC_B200_LIB"
ghz = 1.965
names = ("phase A (thread 0)", "barrier (slowest env)", "stage 1 + barrier", "(unused)", "stage 2, CTA done")
# steady state: the K steps as ONE CUDA graph (as bench.py times them); sums and maxima over all CTAs of all launches
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for t in range(100, 100 + a.steps):
            env.step(acts[t])
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph.replay()
env.lib.sparc_debug_envprof(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
graph.replay()
e1.record()
torch.cuda.synchronize()
env.lib.sparc_debug_envprof(out, 1)
v = [int(x) for x in out]
n = v[0]
print(f"{a.workload}: {n_env} envs, {n // a.steps} CTAs per launch, validate={a.validate}; graph replay {e0.elapsed_time(e1) / a.steps * 1e3:.2f} us per step "
      f"(with the profiling code)")
for k, name in enumerate(names):
    print(f"  {name:24s} mean {v[1 + k] / n:8.0f} cycles = {v[1 + k] / n / ghz / 1e3:6.2f} us   max {v[6 + k]:8d} = {v[6 + k] / ghz / 1e3:6.2f} us")
print("one launch at a time (idle GPU before every launch):")
spans, skews = [], []
acc = [0] * 16
for t in range(100, 100 + a.steps):                 # one launch at a time: the globaltimer span belongs to ONE launch
    env.step(acts[t])
    env.lib.sparc_debug_envprof(out, 1)
    v = [int(x) for x in out]
    for k in range(11):
        acc[k] = max(acc[k], v[k]) if k >= 6 else acc[k] + v[k]
    spans.append(v[12] - v[11])
    skews.append(v[13] / max(v[0], 1) - (v[11] & 0xffffffff))
n = acc[0]
print(f"{a.workload}: {n_env} envs, {n // a.steps} CTAs per launch, validate={a.validate}")
for k, name in enumerate(names):
    print(f"  {name:24s} mean {acc[1 + k] / n:8.0f} cycles = {acc[1 + k] / n / ghz / 1e3:6.2f} us   max {acc[6 + k]:8d} = {acc[6 + k] / ghz / 1e3:6.2f} us")
spans.sort(); skews.sort()
print(f"  first CTA begin -> last CTA end (globaltimer): median {spans[len(spans) // 2] / 1e3:.2f} us, min {spans[0] / 1e3:.2f}, max {spans[-1] / 1e3:.2f}")
print(f"  mean CTA begin after the first CTA's: median {skews[len(skews) // 2] / 1e3:.2f} us")
