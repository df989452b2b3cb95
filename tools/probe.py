"""Many A/B arms in ONE process (a gpurun call costs minutes, a bench.py start ~20 s; an arm here costs ~2 s).

Each positional argument is one arm: space-separated tokens, UPPERCASE=value are environment knobs of the library
(read by sparc_env_create: SPARC_TILE, SPARC_SPLIT, SPARC_ENC_TILE, SPARC_VAL_GRID, ...), lowercase=value are
workload=config2|config3|config4  envs=N  validate=terminal|never  format=reference|compact  update=rewrite|delta
steps=K  obs=0 (no observation: dynamics only)  graph=0 (the K steps launched from Python).  The K steps are one CUDA graph (as in bench.py), replayed `reps` times, best and median reported.

    gpurun -- 'python tools/probe.py "workload=config2" "workload=config2 SPARC_TILE=8" "workload=config2 validate=never"'

The library under test is SPARC_B200_LIB (or the in-tree build): run the script once per build variant.
"""
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from sparc_gym_b200.dataset import config_pool  # noqa: E402
from sparc_gym_b200.vec_env import DeviceStore, SparcVecEnv, _as_store  # noqa: E402

_stores = {}


def store_for(cfg, pool_n):
    key = (cfg, pool_n)
    if key not in _stores:
        _stores[key] = DeviceStore(_as_store(config_pool(cfg, n_puzzles=pool_n)), "cuda:0")
    return _stores[key]


def run_arm(arm, reps=5):
    opts = {"workload": "config4", "validate": "terminal", "format": "reference", "update": "rewrite", "steps": "200", "obs": "1", "graph": "1"}
    envs = {}
    for tok in arm.split():
        k, v = tok.split("=", 1)
        (envs if k.isupper() else opts)[k] = v
    cfg, n_env, pool_n, traceback, max_steps, _ = WORKLOADS[opts["workload"]]
    n_env = int(opts.get("envs", n_env))
    K, W = int(opts["steps"]), 20
    old = {k: os.environ.get(k) for k in envs}
    os.environ.update(envs)
    try:
        env = SparcVecEnv(store_for(cfg, pool_n), n_env, traceback=traceback, max_steps=max_steps, autoreset_mode="same_step",
                          validate=opts["validate"], obs_format=opts["format"], obs_update=opts["update"], write_obs=opts["obs"] != "0")
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    actions = torch.randint(0, 4, (K + W, n_env), dtype=torch.uint8, device="cuda", generator=gen)
    env.reset()
    for t in range(W):
        env.step(actions[t])
    torch.cuda.synchronize()
    graph = None
    if opts["graph"] != "0":
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for t in range(W, W + K):
                    env.step(actions[t])
        torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        ev0.record()
        if graph is not None:
            graph.replay()
        else:
            for t in range(W, W + K):
                env.step(actions[t])
        ev1.record()
        torch.cuda.synchronize()
        ms.append(ev0.elapsed_time(ev1) / K)
    stats = env.stats()
    env.close()
    best, med = min(ms), statistics.median(ms)
    print(f"{arm:70s} | best {best * 1e3:8.2f} us  median {med * 1e3:8.2f} us  {n_env / best * 1e3:.3e} env-steps/s  "
          f"episodes {int(stats[0])}", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    print("lib:", os.environ.get("SPARC_B200_LIB", "in-tree"), flush=True)
    for a in sys.argv[1:]:
        try:
            run_arm(a)
        except Exception as ex:  # an arm with a bad knob must not cost the others their turn
            print(f"{a:70s} | FAILED: {ex}", flush=True)
    print("total %.1f s" % (time.time() - t0))
