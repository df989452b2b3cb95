"""GPU, BASELINE.json sizes, against the ORACLE: config-3 (65,536 envs, mixed sizes, traceback) and config-4 (131,072
envs, 6x6 full rules) batches run through the default launch strategy (the split step k_dyn -> k_val || k_encode, no
environment overrides); 256 env indices are drawn at random and exactly those envs' action streams are replayed in
oracle/sparc_oracle.py, comparing the observation, reward, flags, legal mask, step counter and agent location of the
sampled envs after every one of >= 100 steps, autoresets included (SG.py:1111-1238)."""
import numpy as np
import pytest
import torch

from parity import OracleLockstep, compare
from sparc_gym_b200.dataset import config_pool

pytestmark = pytest.mark.gpu
WINDOW = 64           # consecutive puzzles an oracle group can follow its envs through (cyclic cursor, stride 1)


def _sample(n_env, pool, n_groups, per_group, rng):
    """n_groups random base puzzles; per group `per_group` random envs that start on that puzzle (env i starts on
    puzzle i mod pool)."""
    bases = rng.choice(pool, n_groups, replace=False)
    groups = []
    for b in bases:
        ks = rng.choice(n_env // pool, per_group, replace=False)
        groups.append((int(b), np.sort(b + ks * pool)))
    return groups


def _run(cfg, n_env, pool_n, traceback, max_steps, T, seed):
    from sparc_gym_b200.vec_env import SparcVecEnv
    rows = config_pool(cfg, n_puzzles=pool_n)
    env = SparcVecEnv(rows, n_env, traceback=traceback, max_steps=max_steps, autoreset_mode="next_step",
                      validate="terminal", device="cuda:0")
    rng = np.random.default_rng(seed)
    groups = _sample(n_env, pool_n, 16, 16, rng)
    idx = np.concatenate([g[1] for g in groups])
    assert len(idx) == 256 and len(set(idx.tolist())) == 256
    oracles = []
    for base, envs in groups:
        window = [rows[(base + k) % pool_n] for k in range(WINDOW)]
        orc = OracleLockstep(window, len(envs), env.store.plane_keys, (env.store.slab_x, env.store.slab_y), traceback,
                             max_steps, autoreset_mode="next_step", puzzle_stride=1)
        oracles.append(orc)
    t_idx = torch.from_numpy(idx).cuda()

    def gathered(res):
        obs, reward, term, trunc, info = res
        return ({k: v[t_idx] for k, v in obs.items()}, reward[t_idx], term[t_idx], trunc[t_idx],
                {k: v[t_idx] for k, v in info.items()})

    out = env.reset()
    exp = [o.reset(puzzle_idx=np.zeros(o.N, np.int64)) for o in oracles]
    cat = lambda parts: {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    compare(gathered((out[0], env.reward, env.terminated, env.truncated, out[1])), cat(exp), "reset")
    gen = torch.Generator(device="cuda").manual_seed(seed)
    episodes = np.zeros(len(idx), np.int64)
    n_valid = 0
    for t in range(T):
        a = torch.randint(0, 4, (n_env,), dtype=torch.uint8, device="cuda", generator=gen)
        a_s = a[t_idx].cpu().numpy()
        valid, exps, off = [], [], 0
        for o in oracles:
            was_done = list(o.done)
            v = o.step(a_s[off:off + o.N])
            episodes[off:off + o.N] += np.array(was_done)          # a finished env loads its next puzzle on this step
            valid += v
            exps.append(o.expected())
            off += o.N
        assert episodes.max() < WINDOW - 1, "an env outran its oracle window"
        exp = cat(exps)
        cf = [v and bool(exp["terminated"][i] or exp["truncated"][i]) for i, v in enumerate(valid)]
        n_valid += sum(cf)
        compare(gathered(env.step(a)), exp, f"step {t}", check_flags=cf)
    assert env.kernel_launches >= 3 * T            # the split step (k_dyn, k_val, k_encode), not the fused kernel
    return n_valid, int(episodes.sum())


def test_config3_batch_sampled_envs_match_the_oracle():
    """(max_steps = 40 instead of the config's 1,000: with traceback a random agent hardly ever gets stuck, and the
    test wants every sampled env to end episodes -- truncation, terminal validation, autoreset -- inside 110 steps)"""
    n_valid, episodes = _run(3, 65536, 1024, True, 40, 110, seed=31)
    assert n_valid > 400 and episodes > 400


def test_config3_batch_long_episodes():
    n_valid, episodes = _run(3, 65536, 1024, True, 1000, 100, seed=33)
    assert n_valid > 10


def test_config4_batch_sampled_envs_match_the_oracle():
    n_valid, episodes = _run(4, 131072, 4096, False, 1000, 100, seed=32)
    assert n_valid > 300 and episodes > 300
