"""GPU, BASELINE.json sizes: size-independent properties the domain offers, at 131,072 envs
(config 4's per-GPU shard) where no CPU oracle can follow step by step.

* known answer: replaying a puzzle's listed solution gives +0.01 on every move and
  +1 / terminated at the end, and the planted solution satisfies all eight rules (checked against the
  unmodified reference for these generator settings in tests/golden);
* one-hot invariants of the observation after every kind of step;
* conservation laws of the episode statistics.
"""
import numpy as np
import pytest
import torch

from fuzzrows import solution_actions
from sparc_gym_b200.dataset import config_pool

pytestmark = pytest.mark.gpu
N = 131072


@pytest.fixture(scope="module")
def rows():
    return config_pool(4, n_puzzles=512)


def test_every_listed_solution_wins_and_satisfies_all_rules(rows):
    from sparc_gym_b200.vec_env import SparcVecEnv
    env = SparcVecEnv(rows, N, traceback=False, max_steps=1000, autoreset_mode="disabled", validate="terminal")
    sols = [solution_actions(r, 0) for r in rows]
    T = max(map(len, sols))
    table = np.zeros((len(rows), T), np.uint8)
    length = np.array([len(s) for s in sols])
    for i, s in enumerate(sols):
        table[i, :len(s)] = s
    pz = np.arange(N) % len(rows)
    acts = torch.from_numpy(np.ascontiguousarray(table[pz].T)).cuda()          # [T, N]
    env.reset()
    ret = torch.zeros(N, dtype=torch.int32, device="cuda")
    won = torch.zeros(N, dtype=torch.bool, device="cuda")
    flags_at_end = torch.zeros(N, dtype=torch.int32, device="cuda")
    ln = torch.from_numpy(length[pz]).cuda()
    for t in range(T):
        obs, r, te, tr, info = env.step(acts[t])
        live = ln > t
        code = info["reward_code"].to(torch.int32)
        assert bool(((code == 1) | (code == 3))[live].all()), f"step {t}: a solution move was not rewarded"
        ending = ln == t + 1
        assert bool((te.bool() == ending)[live].all()) and not bool(tr.bool()[live].any())
        won |= ending & (code == 3)
        flags_at_end = torch.where(ending, info["rule_flags"].to(torch.int32) & 0xFFFF, flags_at_end)
        ret += torch.where(live, (code == 1).to(torch.int32) + 100 * (code == 3).to(torch.int32), 0)
    assert bool(won.all())
    assert bool((flags_at_end == (0x1FF | 0x8000)).all()), "a planted solution violates a rule"
    st = env.stats().cpu().tolist()
    assert st[0] == N and st[1] == N and st[2] == 0 and st[3] == 0 and st[7] == N      # episodes = solved = rules_pass
    assert st[5] == int(ret.sum()) and st[6] == int(ln.sum())                          # return (cents), length
    # final observation: agent on the target, visited == solution path length
    base = obs["base"]
    assert bool((base[:, 2] * base[:, 3]).sum(dim=(1, 2)).eq(1).all())
    assert bool(base[:, 0].sum(dim=(1, 2)).eq(ln + 1).all())


def test_one_hot_invariants_and_statistics_under_random_play(rows):
    from sparc_gym_b200.vec_env import SparcVecEnv
    env = SparcVecEnv(rows, N, traceback=True, max_steps=50, autoreset_mode="same_step", validate="terminal")
    obs, info = env.reset()
    static0 = obs["base"][:, [1, 3, 4, 5, 6, 7, 8, 9]].clone()
    color0, info0 = obs["color"].clone(), obs["additional_info"].clone()
    g = torch.Generator(device="cuda").manual_seed(7)
    done_total = 0
    for t in range(60):
        a = torch.randint(0, 4, (N,), dtype=torch.uint8, device="cuda", generator=g)
        obs, r, te, tr, info = env.step(a)
        base = obs["base"]
        assert int(base.min()) == 0 and int(base.max()) == 1
        assert bool(base[:, 2].sum(dim=(1, 2)).eq(1).all())                        # exactly one agent cell
        assert bool((base[:, 2] * base[:, 0]).sum(dim=(1, 2)).eq(1).all())         # ... and it is visited
        assert not bool((base[:, 0] * base[:, 1]).any())                           # nobody stands on a gap
        assert not bool((te.bool() & tr.bool()).any())
        done = te.bool() | tr.bool()
        done_total += int(done.sum())
        code = info["reward_code"]
        assert bool(((code >= 3) == done).all())                                   # +-1 exactly at episode ends
        assert bool(((info["rule_flags"].to(torch.int32) & 0x8000) != 0).eq(done).all())
    # puzzles cycle with stride 1 through a pool of 512: static planes must equal those of puzzle (i + k) mod pool
    st = env.stats().cpu().tolist()
    assert st[0] == done_total and st[1] + st[2] == st[0] and st[4] == 60 * N
    assert tuple(static0.shape[1:]) == (8, 13, 13) and color0.dtype == torch.int32 and info0.dtype == torch.int64
