"""Step-time tail of the terminal validation under a PERFECT agent: every env replays the listed solution of its
puzzle, so every episode ends with all rules satisfied and the exact-fit search runs on every poly region (the
most validation-heavy workload there is).  Pool = the config-4 puzzles with `--dense` of them replaced by dense
polyomino boards (dataset.gen_dense_poly_puzzle: 1e3..1e4+ search nodes each).  Prints median / p99 / max step
time without and with the dense boards.

    python tools/hard_tail.py [--dense 41] [--steps 400]         (run under gpurun)
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def run(n_dense, steps, agent="perfect"):
    import numpy as np
    import torch
    from fuzzrows import solution_actions
    from sparc_gym_b200.dataset import config_pool, gen_dense_poly_puzzle
    from sparc_gym_b200.vec_env import SparcVecEnv
    N, n_pz = 131072, 4096
    rows = config_pool(4, n_puzzles=n_pz)
    rng = np.random.default_rng(0)
    for k, slot in enumerate(rng.choice(n_pz, n_dense, replace=False)):
        rows[int(slot)] = gen_dense_poly_puzzle(6, 6, k)
    sols = [solution_actions(r, 0) for r in rows]
    table = torch.zeros((n_pz, max(map(len, sols)) + 1), dtype=torch.uint8)
    for i, s in enumerate(sols):
        table[i, :len(s)] = torch.tensor(s, dtype=torch.uint8)
    table = table.cuda()
    env = SparcVecEnv(rows, N, max_steps=1000, autoreset_mode="same_step", validate="terminal", device="cuda:0")
    env.reset()
    pz = (torch.arange(N, device="cuda") % n_pz)
    t_idx = torch.zeros(N, dtype=torch.long, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    warm = 100                                       # solutions differ in length: the envs desynchronise on their own
    for t in range(warm + steps):
        a = table[pz, t_idx] if agent == "perfect" else torch.randint(0, 4, (N,), dtype=torch.uint8, device="cuda")
        if t >= warm:
            ev[t - warm].record()
        _, _, te, tr, _ = env.step(a)
        done = (te | tr).bool()
        t_idx = torch.where(done, torch.zeros_like(t_idx), t_idx + 1)
        pz = torch.where(done, (pz + 1) % n_pz, pz)
    ev[steps].record()
    torch.cuda.synchronize()
    ms = np.array([ev[t].elapsed_time(ev[t + 1]) for t in range(steps)])
    st = env.stats().cpu().numpy()
    print(f"{agent} agent, SPARC_LINGER_BUDGET={os.environ.get('SPARC_LINGER_BUDGET', 'default')}, dense boards {n_dense}/{n_pz}: "
          f"step ms median {np.median(ms):.3f} p99 {np.percentile(ms, 99):.3f} max {ms.max():.3f} "
          f"(incl. the torch policy ops); episodes {st[0]}, solved {st[1]}, rules_pass {st[7]}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dense", type=int, default=41)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--agent", default="perfect", choices=["perfect", "random"],
                    help="perfect: every env replays its listed solution (every episode ends in a full validation); random: uniform actions")
    a = ap.parse_args()
    if a.run:
        run(a.dense, a.steps, a.agent)
    else:
        for dense in (0, a.dense):
            subprocess.run([sys.executable, os.path.abspath(__file__), "--run", "--dense", str(dense), "--steps", str(a.steps), "--agent", a.agent])
