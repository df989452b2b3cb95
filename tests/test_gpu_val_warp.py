"""GPU: the warp-cooperative terminal validator (k_val -> validate_warp, sparc_val_warp.cuh) against the scalar
validator (k_validate -> validate_env, the code the CPU build of the tests pins to the oracle) on the SAME terminal
states, at split-step batch sizes: fuzzed rows (stacked / off-region ylops, colliding triangle counts, multi-instance
shapes, colourless stars), generated full-rule boards of every size, and dense polyomino boards whose exact-fit
search needs 1e3..1e4 nodes (the 32-way split search).  Envs freeze when they finish (autoreset disabled), so the
state k_val judged is still there for sparc_validate to judge again."""
import numpy as np
import pytest
import torch

from fuzzrows import fuzz_row, solution_actions
from sparc_gym_b200.dataset import RULES_FULL, gen_dense_poly_puzzle, gen_pool

pytestmark = pytest.mark.gpu


def _policy_table(rows, T, rng):
    """Per puzzle an action row: its listed solution (when it has one), padded with random actions."""
    table = rng.integers(0, 4, (len(rows), T)).astype(np.uint8)
    for i, r in enumerate(rows):
        if r["solutions"]:
            s = solution_actions(r, int(rng.integers(len(r["solutions"]))))[:T]
            table[i, :len(s)] = s
    return table


def _run(rows, N, T, traceback, max_steps, seed, follow, obs_update="rewrite"):
    from sparc_gym_b200.vec_env import SparcVecEnv
    rng = np.random.default_rng(seed)
    env = SparcVecEnv(rows, N, traceback=traceback, max_steps=max_steps, autoreset_mode="disabled", validate="terminal",
                      obs_update=obs_update)
    assert N >= 8192                                         # the split step: k_dyn -> k_val || k_encode
    table = _policy_table(rows, T, rng)
    pz = np.arange(N) % len(rows)
    # an env follows its puzzle's row up to a random deviation step, then plays randomly
    dev = np.where(rng.random(N) < follow, T, rng.integers(0, T, N))
    env.reset()
    judged = wins = passes = 0
    for t in range(T):
        a = np.where(t < dev, table[pz, t], rng.integers(0, 5, N)).astype(np.uint8)     # 4 = an invalid action id
        _, _, te, tr, info = env.step(torch.from_numpy(a).cuda())
        flags = info["rule_flags"].to(torch.int32) & 0xFFFF
        fresh = (flags & 0x8000) != 0                        # validated this step = finished this step
        done = (te | tr).bool()
        assert bool((fresh <= done).all())
        if not bool(fresh.any()):
            continue
        scalar = env.validate(mask=fresh).to(torch.int32) & 0xFFFF
        bad = fresh & (scalar != flags)
        if bool(bad.any()):
            i = int(torch.nonzero(bad)[0])
            raise AssertionError(f"step {t}, env {i}, puzzle {rows[pz[i]]['id']}: k_val {int(flags[i]):#06x} != k_validate "
                                 f"{int(scalar[i]):#06x}; {int(bad.sum())} of {int(fresh.sum())} differ")
        judged += int(fresh.sum())
        wins += int((fresh & (info["reward_code"] == 3)).sum())
        passes += int((fresh & ((flags & 0x100) != 0)).sum())
    st = env.stats().cpu().tolist()
    assert st[0] == judged and st[7] == passes               # RULES_PASS is counted by k_val
    return judged, wins, passes


@pytest.mark.parametrize("obs_update", ["rewrite", "delta"])
def test_fuzzed_rows(obs_update):
    rows = [fuzz_row(s) for s in range(700, 820)]
    judged, wins, passes = _run(rows, 8192 + 77, 48, True, 40, seed=1, follow=0.5, obs_update=obs_update)
    assert judged > 6000 and passes > 50


@pytest.mark.parametrize("traceback,obs_update", [(False, "rewrite"), (True, "rewrite"), (True, "delta")])
def test_generated_full_rule_boards(traceback, obs_update):
    rows = gen_pool(96, [(1, 1), (2, 3), (3, 3), (4, 4), (5, 5), (6, 6), (7, 7), (6, 3)], RULES_FULL, seed0=4000, extra_solutions=1)
    judged, wins, passes = _run(rows, 16384, 150, traceback, 140, seed=2, follow=0.6, obs_update=obs_update)
    assert judged > 12000 and wins > 5000 and passes > 5000


@pytest.mark.parametrize("obs_update", ["rewrite", "delta"])
def test_dense_polyomino_boards(obs_update):
    """(a pool with dense boards switches the node budgets and k_val_hard on: the thread-per-snapshot k_val of the
    full-rewrite step and the warp-per-snapshot k_post of the delta step both hand their over-budget snapshots over)"""
    rows = [gen_dense_poly_puzzle(6, 6, s) for s in range(40)] + [gen_dense_poly_puzzle(5, 4, s) for s in range(8)] + \
        gen_pool(16, [(6, 6)], RULES_FULL, seed0=9000)
    judged, wins, passes = _run(rows, 8192, 30, False, 100, seed=3, follow=0.9, obs_update=obs_update)
    assert judged > 5000 and wins > 5000 and 0 < passes < judged      # some dense boards tile, most do not


def test_more_than_six_ylops():
    """Puzzles with seven ylop instances (6-bit demand counters, header flag bit 2) through the split step's k_val, the
    delta step's k_post and sparc_validate: listed solutions give 0x81ff / 0x807f (areas balanced / not)."""
    from sparc_gym_b200.dataset import gen_many_ylops_puzzle
    from sparc_gym_b200.vec_env import SparcVecEnv
    rows = [gen_many_ylops_puzzle(s, solvable=(s % 2 == 0)) for s in range(16)] + gen_pool(16, [(3, 3), (4, 4)], RULES_FULL, seed0=77)
    N = 8192 + 32
    sols = [solution_actions(r, 0) for r in rows]
    T = max(map(len, sols))
    table = np.full((len(rows), T), 4, np.uint8)
    for i, s in enumerate(sols):
        table[i, :len(s)] = s
    pz = np.arange(N) % len(rows)
    for upd in ("rewrite", "delta"):
        env = SparcVecEnv(rows, N, max_steps=200, autoreset_mode="disabled", validate="terminal", obs_update=upd)
        env.reset()
        flags = torch.zeros(N, dtype=torch.int32, device="cuda")
        for t in range(T):
            _, _, te, tr, info = env.step(torch.from_numpy(np.ascontiguousarray(table[pz, t])).cuda())
            f = info["rule_flags"].to(torch.int32) & 0xFFFF
            flags = torch.where(f != 0, f, flags)
        want = np.where(pz < 16, np.where(pz % 2 == 0, 0x81FF, 0x807F), 0x81FF)
        assert np.array_equal(flags.cpu().numpy(), want), upd
        assert np.array_equal((env.validate().to(torch.int32) & 0xFFFF).cpu().numpy(), want)
