"""A short run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck): the fused step, the split step
(k_dyn -> k_val || k_encode, k_val_hard on a pool with dense boards), the delta step (k_dyn -> k_post) in both
observation formats, masked resets, sparc_encode / sparc_validate / checkpoint, and the single-env drop-in.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py        (under gpurun)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sparc_gym_b200  # noqa: E402
from sparc_gym_b200.dataset import RULES_FULL, gen_dense_poly_puzzle, gen_pool  # noqa: E402
from sparc_gym_b200.vec_env import SparcVecEnv  # noqa: E402

rows = gen_pool(24, [(2, 2), (3, 3), (4, 4), (6, 6)], RULES_FULL, seed0=5) + [gen_dense_poly_puzzle(5, 5, s) for s in range(4)]
gen = torch.Generator(device="cuda").manual_seed(0)
n_steps = int(os.environ.get("SANITIZE_STEPS", 12))
for n_env in (200, 8192 + 64):
    for fmt in ("reference", "compact"):
        for upd in ("rewrite", "delta"):
            env = SparcVecEnv(rows, n_env, traceback=True, max_steps=10, autoreset_mode="same_step", validate="terminal",
                              obs_format=fmt, obs_update=upd)
            env.reset()
            for t in range(n_steps):
                env.step(torch.randint(0, 4, (n_env,), dtype=torch.uint8, device="cuda", generator=gen))
                if t == 5:
                    env.reset(torch.randint(0, len(rows), (n_env,), dtype=torch.int32, device="cuda", generator=gen),
                              torch.randint(0, 2, (n_env,), dtype=torch.uint8, device="cuda", generator=gen))
            sd = env.state_dict()
            env.load_state_dict(sd)
            env.encode()
            env.validate(region_map=True, region_fit=True)
            st = env.stats().cpu().tolist()
            torch.cuda.synchronize()
            print(n_env, fmt, upd, "episodes", st[0], "rules_pass", st[7], flush=True)
            env.close()
one = sparc_gym_b200.make("SPaRc-Gym", puzzles=rows[:6], traceback=True, max_steps=30)
one.reset()
for a in np.random.default_rng(0).integers(0, 4, 40).tolist():
    one.step(a)[4]["rule_status"]
one.close()
print("sanitize run done")
