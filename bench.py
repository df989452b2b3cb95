#!/usr/bin/env python
"""bench.py -- env-steps/s of the B200 SPaRC environment (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank/GPU)
    python bench.py --impl reference ...                   (the reference's CPU algorithm, host cores)

A "step" is one lockstep pass of the hot path over all envs of the workload: actions ->
dynamics -> terminal rule validation -> autoreset -> full rewrite of the 'new' observation
(int32 base planes, int32 colour, int64 additional_info) into preallocated torch tensors.

Default workload = BASELINE.json configs[3] ("1M envs, 6x6 full rule set with 8 colors, sharded
across 8xB200"), i.e. the configuration the north-star target (>=1e9 steps/s on 8 GPUs at >=50%
of HBM roofline, 6x6 full-rule puzzles) is quoted on: 131,072 envs PER GPU (weak scaling), pool of
4,096 synthetic 6x6 puzzles.  `--workload config2|config3` select the other GPU configs.
Prints ONE JSON line (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config id, envs per GPU, pool size, traceback, max_steps, description)
    # BASELINE configs[0]: the reference's own CPU-runnable case; through the N=1 Gymnasium drop-in (--workload config1
    # times sparc_gym_b200.make(...).step(), one env, host in / host out: see dropin_bench)
    "config1": (1, 1, 1, False, 1000, "single 4x4 gaps+dots puzzle, random-action rollout via the gym.make surface (one env)"),
    "config2": (2, 4096, 64, False, 1000, "4,096 lockstep envs, 4x4 gaps+dots puzzles (9x9 lattice, 5 planes)"),
    "config3": (3, 65536, 1024, True, 1000, "65,536 envs, mixed 3x3-6x6 full-rule puzzles padded to 13x13, traceback, terminal validation"),
    "config4": (4, 131072, 4096, False, 1000, "131,072 envs per GPU (1M on 8 GPUs), 6x6 full rule set, 8 colours (13x13 lattice, 10 planes)"),
    # BASELINE configs[4]: the env-count sweep; one point per run (--envs-per-gpu N), tools/env_sweep.py drives 2^10..2^22
    "config5": (5, 65536, 1024, True, 1000, "env-count sweep point: --envs-per-gpu N envs per GPU, heterogeneous 3x3-6x6 full-rule puzzles padded to 13x13, traceback"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, n_env):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_env launch from the committed `ncu --set full` capture of
    this workload (profiles/r1_traffic.json), or None when the capture does not match the run."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            break
    else:
        return None
    with open(p) as f:
        t = json.load(f).get(workload)
    if not t or t.get("envs_per_launch") != n_env:
        return None
    return t["dram_bytes_per_launch"]


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU while it is under load: NVML in-process (pynvml), or the
    nvidia-smi query of B200_PROFILING.md when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period, self.rows, self.stop_flag = index, period, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK counts visible devices: honour CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            bits = [0x8, 0x40, 0x20, 0x4]     # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap
            return [str(mhz), str(self.max_mhz)] + ["Active" if r & b else "Not Active" for b in bits]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return [c.strip() for c in out.split(",")] if out else None

    def run(self):
        while not self.stop_flag.is_set():
            try:
                row = self._sample()
                if row:
                    self.rows.append(row)
            except Exception:
                pass
            self.stop_flag.wait(self.period if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (neither NVML nor nvidia-smi)"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = [n for k, n in enumerate(self.NAMES) if any(r[2 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference env (baseline/_ref or /root/reference, through oracle/refharness.py) on the host
# cores, one env per worker process -- AsyncVectorEnv's architecture.  Falls back to the oracle port
# (oracle/sparc_oracle.py, reference_cost=True) only when no reference tree is installed, and says so in `kind`.
# ------------------------------------------------------------------------------------------------
_W = {}            # per-worker state (the env lives across samples: the reference's ctor re-parses every row)


def _cpu_worker_init(rows_by_worker, traceback, max_steps, kind, counter):
    with counter.get_lock():
        w = counter.value
        counter.value += 1
    rows = rows_by_worker[w]
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if kind == "reference":
        import refharness as rh
        env = rh.make_reference_env(rows, observation="new", traceback=traceback, max_steps=max_steps)
        _W["zero_leak"] = rh.zero_leak
    else:
        import sparc_oracle as so
        env = so.OracleEnv(rows, observation="new", traceback=traceback, max_steps=max_steps, reference_cost=True)
        _W["zero_leak"] = None
    _W.update(env=env, rows=rows, k=0, acts=np.random.default_rng(w).integers(0, 4, 1 << 16), n=0, fresh=True)


def _cpu_reset():
    env, k = _W["env"], _W["k"]
    if _W["zero_leak"]:
        _W["zero_leak"](env, k)                     # the reference leaks visited / agent planes across episodes
    env.reset(options={"puzzle_id": _W["rows"][k]["id"]})


def _cpu_sample(budget_s):
    """`budget_s` seconds of random-action steps on this worker's env, resetting (next puzzle) when an episode ends."""
    env, acts = _W["env"], _W["acts"]
    if _W["fresh"]:
        _cpu_reset()
        _W["fresh"] = False
    n0, n = _W["n"], _W["n"]
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        _, _, te, tr, _ = env.step(int(acts[n & 0xFFFF]))
        n += 1
        if te or tr:
            _W["k"] = (_W["k"] + 1) % len(_W["rows"])
            _cpu_reset()
    _W["n"] = n
    return n - n0, time.perf_counter() - t0


class CpuArm:
    """Worker pool of CPU envs over a bounded sample of the workload's puzzles (8 per worker)."""

    def __init__(self, workload, kind=None, cores=None):
        import multiprocessing as mp
        from sparc_gym_b200.dataset import config_pool
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refharness as rh
        cfg, _, _, traceback, max_steps, _ = WORKLOADS[workload]
        self.workload = workload
        self.cores = 1 if cfg == 1 else (cores or os.cpu_count() or 1)      # config 1 is ONE env (gym.make, no vectorisation)
        self.kind = kind or ("reference" if rh.reference_available() else "port")
        self.ref_root = rh.REFERENCE_ROOT if self.kind == "reference" else None
        rows = config_pool(cfg, n_puzzles=8 * self.cores)
        by_worker = [rows] if cfg == 1 else [rows[8 * w:8 * w + 8] for w in range(self.cores)]
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_worker_init,
                             initargs=(by_worker, traceback, max_steps, self.kind, ctx.Value("i", 0)))

    def sample(self, budget_s):
        """One bounded sample: every worker steps for `budget_s` seconds.  Returns (steps, env-steps/s summed over workers)."""
        res = self.pool.map(_cpu_sample, [budget_s] * self.cores, chunksize=1)
        return sum(r[0] for r in res), sum(r[0] / r[1] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self, steps, budget_s, n_samples):
        what = ("the UNMODIFIED reference env (SPaRC_Gym.SPaRC_Gym from %s, stub gymnasium/pygame modules, rows injected "
                "through datasets.load_dataset, Q1 plane leak zeroed before each reset)" % os.path.relpath(self.ref_root, ROOT)
                if self.kind == "reference" else "the oracle port (oracle/sparc_oracle.py, reference_cost=True)")
        return (f"{steps} random-action steps of the {self.workload} puzzles ({8 * self.cores} of the pool, 8 per worker) on {what}, "
                f"one env per worker process, {n_samples} sample(s) of {budget_s:.1f}s per worker, rate = sum of per-worker "
                "step rates averaged over the samples; AsyncVectorEnv stand-in (gymnasium absent)")


def cpu_baseline(workload, budget_s=10.0, samples=1, warmup=0, kind=None, cores=None, with_port=False):
    arm = CpuArm(workload, kind=kind, cores=cores)
    for _ in range(warmup):
        arm.sample(budget_s)
    tot, rates = 0, []
    for _ in range(samples):
        n, r = arm.sample(budget_s)
        tot += n
        rates.append(r)
    arm.close()
    out = {"value": round(sum(rates) / len(rates), 1), "unit": "env-steps/s", "cores": arm.cores, "kind": arm.kind,
           "sample": arm.describe(tot, budget_s, samples), "per_sample": [round(r, 1) for r in rates]}
    if with_port and arm.kind == "reference":
        port = CpuArm(workload, kind="port", cores=cores)
        _, r = port.sample(min(budget_s, 5.0))
        port.close()
        out["port_value"] = round(r, 1)
        out["port_note"] = ("the oracle port (reference_cost=True) on the same puzzles and cores, one sample: "
                            "%.2fx the unmodified reference" % (r / max(out["value"], 1e-9)))
    return out


def dropin_bench(n_steps=10000, warmup=500, lazy_info=False, device="cuda:0"):
    """BASELINE configs[0] through the reference's own call surface: gym.make('SPaRc-Gym', observation='new',
    traceback=False, max_steps=1000) -> reset / step(action) with a Python int in and (obs dict of numpy planes, reward,
    terminated, truncated, info) out, one env, every step a full round trip to the GPU.  Default: `info` is the
    reference's plain dict, rule_status built every step; lazy_info=True is the drop-in's opt-in (rule_status filled in
    on first lookup, never looked up here)."""
    import sparc_gym_b200
    from sparc_gym_b200.dataset import config_pool
    row = config_pool(1)[0]
    env = sparc_gym_b200.make("SPaRc-Gym", puzzles=[row], observation="new", traceback=False, max_steps=1000, device=device,
                              lazy_info=lazy_info)
    acts = np.random.default_rng(0).integers(0, 4, n_steps + warmup).tolist()
    env.reset(options={"puzzle_id": row["id"]})
    episodes, t0 = 0, 0.0
    for t, a in enumerate(acts):
        if t == warmup:
            t0 = time.perf_counter()
        obs, r, te, tr, info = env.step(a)
        if te or tr:
            episodes += 1
            env.reset(options={"puzzle_id": row["id"]})
    dt = time.perf_counter() - t0
    env.close()
    return {"value": round(n_steps / dt, 1), "unit": "env-steps/s", "us_per_step": round(1e6 * dt / n_steps, 2), "steps": n_steps,
            "episodes": episodes, "lazy_info": lazy_info}


def workload_config(args, world, store=None):
    """The `config` object of the JSON line: the workload only, identical in both arms (run details go to `run`)."""
    cfg, n_env, pool_n, traceback, max_steps, desc = WORKLOADS[args.workload]
    n_env = args.envs_per_gpu or n_env
    pool_n = args.pool or pool_n
    if store is None:
        from sparc_gym_b200.dataset import config_pool
        from sparc_gym_b200.ingest import ingest_rows
        from sparc_gym_b200.store import PuzzleStore
        store = PuzzleStore(ingest_rows(config_pool(cfg, n_puzzles=pool_n)))
    return {"workload": args.workload, "description": desc, "envs_per_gpu": n_env, "global_envs": n_env * world,
            "pool_puzzles": pool_n, "lattice": [store.slab_x, store.slab_y], "planes": list(store.plane_keys),
            "traceback": traceback, "max_steps": max_steps, "autoreset": args.autoreset, "validate": args.validate,
            "obs_format": args.obs_format, "obs_update": args.obs_update,
            "l2": _l2_note(store.algorithmic_bytes_per_step(args.obs_format) * n_env, args.obs_update),
            "parallelism": f"env-sharded x{world}, no data-path collective"}


def _l2_note(step_bytes, update):
    if update == "delta":
        return "delta update: a step touches a few cells per env (L2-resident); latency-bound, not an HBM measurement"
    if step_bytes > 126e6:
        return "no flush needed: every step rewrites %.0f MB of observation, more than the 126 MB L2" % (step_bytes / 1e6)
    return ("observation per step (%.1f MB) fits in the 126 MB L2: writes may be absorbed by L2, the number is "
            "launch/latency-bound and not an HBM measurement" % (step_bytes / 1e6))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config4", choices=sorted(WORKLOADS))
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--pool", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--validate", default="terminal")
    ap.add_argument("--autoreset", default="same_step")
    ap.add_argument("--no-obs", action="store_true", help="diagnostic: skip the observation rewrite (dynamics only)")
    ap.add_argument("--no-extra", action="store_true", help="skip the short config-2 side measurement")
    ap.add_argument("--obs-format", default="reference", choices=["reference", "compact"],
                    help="reference = int32/int32/int64 planes (the reference's dtypes, the headline); compact = uint8/uint8/int32")
    ap.add_argument("--obs-update", default="rewrite", choices=["rewrite", "delta"],
                    help="rewrite = every step writes every observation in full (the headline); delta = the persistent "
                         "observation is touched in place, full rewrite only for envs that begin a new episode")
    ap.add_argument("--stats-every", type=int, default=0,
                    help="also time the steps (launched from Python, no graph) with the episode statistics reduced over all ranks "
                         "every K steps on a side stream (ShardedRunner.stats_async) inside the timed region; reported under "
                         "`stats_collective` beside the same loop without it")
    ap.add_argument("--obs-to-host", action="store_true",
                    help="also time steps whose WHOLE observation is copied to pinned host memory (what a host-side consumer "
                         "of the reference's API receives); reported under `obs_to_host`, never the headline")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of one CUDA graph of K steps")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    cfg, n_env, pool_n, traceback, max_steps, desc = WORKLOADS[args.workload]
    n_env = args.envs_per_gpu or n_env
    pool_n = args.pool or pool_n

    if args.impl == "reference":
        if rank != 0:
            return
        # One "step" of this arm = one bounded sample: every host core runs the reference's step() on its own env for
        # `budget` seconds.  K timed samples after W warm-up samples, sized so that the whole run ends within ~2.5 min.
        K, W = max(1, args.steps), max(0, args.warmup)
        budget = max(0.5, min(args.cpu_seconds, 150.0 / (K + W)))
        cb = cpu_baseline(args.workload, budget_s=budget, samples=K, warmup=W, with_port=not args.no_extra)
        print(json.dumps({"impl": "reference", "metric": "env-steps/sec (whole box)", "value": cb["value"],
                          "unit": "env-steps/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                          "ms_per_step": round(1e3 * budget, 1),                    # one step = one bounded CPU sample
                          "host_only": True, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int32", "data": "synthetic",
                          "config": workload_config(args, world),
                          "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from sparc_gym_b200.dataset import config_pool
    from sparc_gym_b200.runner import ShardedRunner, reduce_stats

    if args.workload == "config1":                                  # the N=1 drop-in, wall-clock (one env cannot fill a GPU)
        if rank != 0:
            return
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        fast, lazy = dropin_bench(lazy_info=False), dropin_bench(lazy_info=True)
        line = {"metric": "env-steps/sec (whole box)", "value": fast["value"], "unit": "env-steps/s", "n_gpus": 1,
                "steps": fast["steps"], "warmup": 500, "ms_per_step": round(fast["us_per_step"] / 1e3, 5), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": "config1", "description": WORKLOADS["config1"][5], "traceback": False, "max_steps": 1000,
                           "api": "sparc_gym_b200.make('SPaRc-Gym', ...).step(a): one env, host in / host out, wall clock"},
                "dropin": fast, "dropin_lazy_info": lazy}
        if not args.no_cpu_baseline:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", "config1",
                                  "--steps", "2", "--warmup", "1", "--cpu-seconds", "4", "--no-extra"], capture_output=True, text=True)
            try:
                line["cpu_baseline"] = json.loads(res.stdout.strip().splitlines()[-1])["cpu_baseline"]
                line["vs_reference_env_1core"] = round(fast["value"] / line["cpu_baseline"]["value"], 2)
            except Exception:
                line["cpu_baseline"] = {"error": res.stderr[-300:]}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t_build = time.perf_counter()
    rows = config_pool(cfg, n_puzzles=pool_n)
    # env-sharded: rank r owns global envs [r*n_env, (r+1)*n_env); global env i starts on puzzle i mod pool
    runner = ShardedRunner(rows, n_env * world, rank=rank, world=world, local_rank=local_rank, traceback=traceback,
                           max_steps=max_steps, autoreset_mode=args.autoreset, puzzle_stride=1, validate=args.validate,
                           write_obs=not args.no_obs, obs_format=args.obs_format, obs_update=args.obs_update)
    env = runner.env
    st = env.store
    t_build = time.perf_counter() - t_build
    K, W = max(1, args.steps), max(3, args.warmup)        # timing rule: at least 3 untimed warm-up steps
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = torch.randint(0, 4, (K + W, n_env), dtype=torch.uint8, device=dev, generator=gen)
    sampler = ClockSampler(local_rank) if rank == 0 else None     # samples through warm-up, timed steps and e2e
    if sampler:
        sampler.start()
    runner.reset()
    for t in range(W):
        env.step(actions[t])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, CUDA events on the launching stream
    launches0 = env.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph = None
    if not args.no_graph:
        # the K timed steps as ONE CUDA graph (K kernel nodes): removes the per-launch host cost
        # of the Python/ctypes call, which dominates small batches
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):      # capture records, it does not execute
                for t in range(W, W + K):
                    env.step(actions[t])
        torch.cuda.current_stream().wait_stream(side)
        # one untimed replay first (K more warm-up steps): the first launch of a graph also uploads it to the device,
        # which is not part of a step (config 2: 10.1 vs 8.4 us per step; config 4: 0.180 vs 0.177 ms)
        graph.replay()
    barrier()
    ev0.record()
    if graph is not None:
        graph.replay()
    else:
        for t in range(W, W + K):
            env.step(actions[t])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.kernel_launches - launches0          # the ABI counts at capture time too: kernel nodes of the graph
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank_ms = [ms]
    if world > 1:
        gathered = [torch.zeros_like(t_ms) for _ in range(world)]
        dist.all_gather(gathered, t_ms)
        per_rank_ms = [float(g.item()) for g in gathered]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())

    # ---- the dominant kernel on its own: K launches of k_encode (state -> observation), CUDA events
    enc_ms = None
    if not args.no_obs and args.obs_update == "rewrite":
        for _ in range(3):
            env.encode()
        barrier()
        ev0.record()
        for _ in range(K):
            env.encode()
        ev1.record()
        barrier()
        e_t = torch.tensor([ev0.elapsed_time(ev1) / K], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
        enc_ms = float(e_t.item())

    # ---- end to end: host action buffer -> H2D -> step -> D2H reward/terminated/truncated, every step
    e2e = None
    if not args.no_e2e:
        # fresh uniform actions every step, as in the device-resident loop (one vector repeated K times walks every env
        # into a wall within a few steps: no episode ends, nothing to validate or reset -- an easier workload)
        a_host = torch.randint(0, 4, (K + 3, n_env), dtype=torch.uint8, generator=torch.Generator().manual_seed(4321 + rank)).pin_memory()
        r_host = torch.empty(n_env, dtype=torch.float32).pin_memory()
        te_host = torch.empty(n_env, dtype=torch.uint8).pin_memory()
        tr_host = torch.empty(n_env, dtype=torch.uint8).pin_memory()
        for t in range(3):
            env.step_host(a_host[K + t], r_host, te_host, tr_host)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for t in range(K):
            env.step_host(a_host[t], r_host, te_host, tr_host)
        ev1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e_ms = torch.tensor([max(ev0.elapsed_time(ev1), 0.0)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        e2e = {"value": round(n_env * world * K / (float(e_ms.item()) / 1e3), 1), "unit": "env-steps/s",
               "h2d_bytes_per_step": n_env * world, "d2h_bytes_per_step": n_env * world * 6,
               "wall_ms_per_step": round(wall_ms / K, 4),
               "api": "SparcVecEnv.step_host -> sparc_step_host (pinned host actions in; reward f32 + terminated u8 + "
                      "truncated u8 out; observation tensors stay on the device for the policy; the call returns when the "
                      "results are in the host buffers, the observation rewrite of that step is stream-ordered behind it "
                      "and overlaps the caller's turn-around; the timed region ends with a full synchronise)"}

    # ---- the statistics collective inside the loop: K steps from Python with and without stats_async
    stats_coll = None
    if args.stats_every > 0:
        def loop():
            barrier()
            ev0.record()
            for t in range(W, W + K):
                runner.step(actions[t])
            ev1.record()
            barrier()
            t_ = torch.tensor([ev0.elapsed_time(ev1) / K], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return float(t_.item())
        for t in range(3):
            runner.step(actions[t])
        a = runner.stats_async(every=args.stats_every)
        for t in range(args.stats_every):
            runner.step(actions[t])                  # first reduction (communicator set-up) outside the timed region
        a.latest(wait=True)
        base_ms = with_ms = 1e9
        for _ in range(4):                           # alternating, best of 4 each: Python-launched loops jitter by a few %
            runner._async = None
            base_ms = min(base_ms, loop())
            runner._async = a
            with_ms = min(with_ms, loop())
        got = a.latest(wait=True)
        runner._async = None
        stats_coll = {"every": args.stats_every, "ms_per_step_without": round(base_ms, 5), "ms_per_step_with": round(with_ms, 5),
                      "overhead_frac": round(with_ms / base_ms - 1.0, 4), "reductions": a.reductions,
                      "latest_tick": got[0] if got else None, "latest_episodes": got[1]["episodes"] if got else None,
                      "how": "K steps launched from Python (no CUDA graph), max over ranks, alternating, best of 4; every `every` steps k_read_stats + "
                             "all_reduce(int64[8]) on a side stream through a process group of its own (NCCL over NVLink)"}

    # ---- the whole observation to the host, every step (a host consumer; PCIe-bound)
    obs_host = None
    if (args.obs_to_host or (world == 1 and args.workload == "config4" and not args.no_extra)) and not args.no_obs:
        host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in env.obs.items()}
        nbytes = sum(v.numel() * v.element_size() for v in host.values())
        K2 = max(3, min(K, 8))
        for t in range(2):
            env.step(actions[t])
            for k in host:
                host[k].copy_(env.obs[k], non_blocking=True)
        barrier()
        ev0.record()
        for t in range(K2):
            env.step(actions[W + t % K])
            for k in host:
                host[k].copy_(env.obs[k], non_blocking=True)     # same stream: the next step may not overwrite what is being copied
        ev1.record()
        barrier()
        oh_ms = ev0.elapsed_time(ev1) / K2
        obs_host = {"value": round(n_env * world / (oh_ms / 1e3), 1), "unit": "env-steps/s", "ms_per_step": round(oh_ms, 4),
                    "d2h_bytes_per_step": nbytes * world, "d2h_GBs_per_gpu": round(nbytes / (oh_ms / 1e3) / 1e9, 2), "steps": K2,
                    "obs_format": args.obs_format,
                    "note": "step + cudaMemcpyAsync of base / color / additional_info into pinned host tensors, one stream; bound by "
                            "the host link, not by the kernels"}
        del host

    if sampler and not args.no_e2e:
        # keep the GPU under the same load for ~1.5 s so that nvidia-smi (100 ms period) sees it
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            if graph is not None:
                graph.replay()
            else:
                env.step(actions[W])
            torch.cuda.synchronize()
    clocks = sampler.summary() if sampler else None

    # ---- episode statistics: the only cross-GPU traffic (NCCL all-reduce of int64[8])
    stats = reduce_stats(env.stats())

    if rank == 0:
        peak, peak_src = measured_peaks()
        b_alg = st.algorithmic_bytes_per_step(args.obs_format)
        if args.obs_update == "delta":
            # state + io + four touched plane elements + the full rewrites of the envs that began a new episode,
            # amortised with the measured episode rate of this run
            elem = 4 if args.obs_format == "reference" else 1
            rate = stats["episodes"] / max(stats["steps"], 1)
            b_alg = int(round(208 + 4 * elem + rate * (b_alg - 208)))
        value = n_env * world * K / (ms_max / 1e3)
        step_ms = ms_max / K
        fused = launches == K                      # small batches: the single fused kernel k_env is the step
        kernel_ms = step_ms if (fused or enc_ms is None) else enc_ms
        # the dominant kernel's OWN algorithmic bytes: k_encode writes the observation and reads 16 + 4*nv + 4 bytes of
        # state per env; it neither writes state nor touches actions / rewards (those belong to k_dyn and are in b_alg)
        S = st.slab_x * st.slab_y
        k_bytes = b_alg if (fused or enc_ms is None) else (b_alg - 208) + 16 + 4 * ((S + 31) // 32) + 4
        achieved = k_bytes * n_env / (kernel_ms / 1e3) / 1e9
        step_achieved = b_alg * n_env / (step_ms / 1e3) / 1e9
        write_peak = 7250.0          # tools/write_peak.py: 1.15 GB memset on a B200 of this pool (a write-only stream beats the copy figure)
        line = {
            "metric": "env-steps/sec (whole box)", "value": round(value, 1), "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(ms_max / K, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": workload_config(args, world, st),
            "run": {"launch": ("one CUDA graph of the K steps, replayed once untimed (K more warm-up steps) and once timed" if graph is not None else "K steps launched from Python")
                              + f" ({launches // K} kernel launches per step)",
                    "obs_mode": ("full rewrite" if args.obs_update == "rewrite" else
                                 "DELTA: persistent observation touched in place, envs that begin a new episode rewritten in "
                                 "full (own algorithmic bytes; latency-bound, not HBM-bound)") +
                                (", reference dtypes" if args.obs_format == "reference" else
                                 ", COMPACT dtypes (uint8 planes, uint8 colour, int32 additional_info): own algorithmic bytes"),
                    "setup_s": round(t_build, 1),
                    "ms_per_step_by_rank": [round(m / K, 5) for m in per_rank_ms]},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": ncu_traffic(args.workload, n_env) if (args.obs_format == "reference" and args.obs_update == "rewrite") else None, "peak_source": peak_src,
                         "kernel": "k_env (fused step)" if fused else
                                   "the whole step = k_dyn (dynamics + in-place observation update), then k_post (terminal validation + rewrite of the envs that began a new episode)" if args.obs_update == "delta" else
                                   "k_encode (state -> observation; the step = k_dyn, then k_val || k_encode)",
                         "algorithmic_bytes_per_env_step": k_bytes, "envs_per_launch": n_env,
                         "bytes_note": ("the kernel's own bytes: observation written + state read" if k_bytes != b_alg else
                                        "whole-step bytes (SURVEY 8d): observation + state read and written + io"),
                         "kernel_ms": round(kernel_ms, 5), "share_of_step": round(kernel_ms / step_ms, 4),
                         "frac_of_write_only_ceiling": round(achieved / write_peak, 4),
                         "whole_step": {"achieved": round(step_achieved, 1), "frac": round(step_achieved / peak, 4),
                                        "ms": round(step_ms, 5), "algorithmic_bytes_per_env_step": b_alg,
                                        "frac_of_write_only_ceiling": round(step_achieved / write_peak, 4)},
                         "write_only_ceiling_GBs": write_peak,
                         "note": "a write-only stream exceeds the measured COPY bandwidth (peak): frac can pass 1; "
                                 "frac_of_write_only_ceiling is against the measured memset rate"},
            "episode_stats": stats,
        }
        if obs_host is not None:
            line["obs_to_host"] = obs_host
        if stats_coll is not None:
            line["stats_collective"] = stats_coll
        if world == 1 and args.workload == "config4" and not args.no_extra and not args.no_obs:
            # BASELINE configs[1] (4,096 envs, 4x4 gaps+dots: a launch/latency-bound batch) measured in the same
            # run, for reference; the headline stays the configuration the north-star target is quoted on
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "config2", "--steps", "500",
                                  "--warmup", "20", "--no-cpu-baseline", "--no-extra"], capture_output=True, text=True)
            try:
                j2 = json.loads(res.stdout.strip().splitlines()[-1])
                line["other_workloads"] = {"config2": {
                    "description": j2["config"]["description"], "value": j2["value"], "unit": j2["unit"],
                    "ms_per_step": j2["ms_per_step"], "e2e": j2["e2e"]["value"] if j2.get("e2e") else None,
                    "roofline_frac": j2["roofline"]["frac"], "kernel": j2["roofline"]["kernel"],
                    "note": "11.5 MB of observation per step: L2-resident and launch/latency bound, not HBM bound"}}
            except Exception:
                line["other_workloads"] = {"config2": {"error": res.stderr[-200:]}}
            # BASELINE configs[0] through the gym.make surface (one env, host in / host out), beside the reference env on one core
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "config1"], capture_output=True, text=True)
            try:
                j1 = json.loads(res.stdout.strip().splitlines()[-1])
                line["other_workloads"]["config1_dropin"] = {
                    "description": j1["config"]["description"], "api": j1["config"]["api"], "value": j1["value"], "unit": j1["unit"],
                    "us_per_step": j1["dropin"]["us_per_step"], "with_lazy_info": j1["dropin_lazy_info"]["value"],
                    "reference_env_1core": j1.get("cpu_baseline", {}).get("value"), "vs_reference_env_1core": j1.get("vs_reference_env_1core")}
            except Exception:
                line["other_workloads"]["config1_dropin"] = {"error": (res.stderr or res.stdout)[-200:]}
            # the same workload in the compact observation format (uint8 / uint8 / int32), under its OWN algorithmic bytes
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--steps", "300",
                                  "--warmup", "20", "--no-cpu-baseline", "--no-extra", "--no-e2e", "--obs-format", "compact", "--obs-to-host"],
                                 capture_output=True, text=True)
            try:
                j3 = json.loads(res.stdout.strip().splitlines()[-1])
                line["other_workloads"][args.workload + "_compact_obs"] = {
                    "description": "same envs and puzzles, observation in uint8 planes / uint8 colour / int32 additional_info",
                    "value": j3["value"], "unit": j3["unit"], "ms_per_step": j3["ms_per_step"],
                    "algorithmic_bytes_per_env_step": j3["roofline"]["whole_step"]["algorithmic_bytes_per_env_step"],
                    "roofline_frac_whole_step": j3["roofline"]["whole_step"]["frac"],
                    "roofline_frac_encoder": j3["roofline"]["frac"], "obs_to_host": j3.get("obs_to_host")}
            except Exception:
                line["other_workloads"][args.workload + "_compact_obs"] = {"error": res.stderr[-200:]}
            # the same workload with the observation updated in place (reference dtypes), under its OWN algorithmic bytes
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--steps", "300",
                                  "--warmup", "20", "--no-cpu-baseline", "--no-extra", "--no-e2e", "--obs-update", "delta"],
                                 capture_output=True, text=True)
            try:
                j4 = json.loads(res.stdout.strip().splitlines()[-1])
                line["other_workloads"][args.workload + "_delta_obs"] = {
                    "description": "same envs and puzzles, the persistent observation touched in place (visited / agent_location "
                                   "cells), envs that begin a new episode rewritten in full; same tensor contents after every step",
                    "value": j4["value"], "unit": j4["unit"], "ms_per_step": j4["ms_per_step"],
                    "algorithmic_bytes_per_env_step": j4["roofline"]["whole_step"]["algorithmic_bytes_per_env_step"],
                    "roofline_frac_whole_step": j4["roofline"]["whole_step"]["frac"],
                    "note": "latency-bound (k_dyn, then k_post: terminal validation + new-episode rewrites), not HBM-bound"}
            except Exception:
                line["other_workloads"][args.workload + "_delta_obs"] = {"error": res.stderr[-200:]}
        if not args.no_cpu_baseline and world == 1:
            # fresh interpreter: the worker pool must not be forked from a process holding a CUDA context
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload",
                                  args.workload, "--steps", "2", "--warmup", "0", "--cpu-seconds", str(min(args.cpu_seconds, 8.0))],
                                 capture_output=True, text=True)
            try:
                line["cpu_baseline"] = json.loads(res.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception:
                line["cpu_baseline"] = {"value": None, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: " + res.stderr[-300:]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
