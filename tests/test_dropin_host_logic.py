"""CPU: the host logic of the single-env Gymnasium drop-in (sparc_gym_b200/gym_env.py: containers, key order,
reward types, `info['rule_status']` flags / details / region summaries, reset cursor, 'SPaRC' text overlays)
replayed against the golden traces of the UNMODIFIED reference.  There is no GPU here, so the CUDA vec env is
swapped (monkeypatch, test only) for the host-compiled build of the same device code (tests/emu); the GPU
version of this test (test_gpu_gym_dropin.py) runs the real library."""
import gzip
import json
import os

import numpy as np
import pytest
import torch

from emu.emu_env import EmuVecEnv
from tracing import diff_traces, run_episode

G = os.path.join(os.path.dirname(__file__), "golden")


class EmuTorchVecEnv(EmuVecEnv):
    """EmuVecEnv with SparcVecEnv's tensor-typed surface (torch views of the same numpy buffers)."""

    device = torch.device("cpu")

    def __init__(self, puzzles, num_envs, device=None, write_obs=True, **kw):
        super().__init__(puzzles, num_envs, **kw)
        self._np = {k: getattr(self, k) for k in ("reward", "reward_code", "terminated", "truncated", "legal_mask",
                                                   "step_count", "agent_xy")}
        for k, a in self._np.items():
            setattr(self, k, torch.from_numpy(a))
        self.obs = {k: torch.from_numpy(v) for k, v in self.obs.items()}

    def reset(self, puzzle_idx=None, mask=None):
        return super().reset(None if puzzle_idx is None else np.asarray(puzzle_idx), mask)

    def step(self, actions):
        return super().step(np.asarray(actions))

    def validate(self, **kw):
        out = super().validate(**kw)
        res = tuple(torch.from_numpy(np.ascontiguousarray(o).view(np.int16) if o.dtype == np.uint16 else o)
                    for o in (out if isinstance(out, tuple) else (out,)))
        return res if len(res) > 1 else res[0]          # like SparcVecEnv.validate: the flags alone when nothing else is asked for

    def close(self):
        pass


@pytest.fixture()
def dropin(monkeypatch):
    import sparc_gym_b200.vec_env as ve
    import sparc_gym_b200.gym_env as ge
    monkeypatch.setattr(ve, "SparcVecEnv", EmuTorchVecEnv)
    return ge


def test_dropin_host_logic_replays_reference_traces(dropin):
    with gzip.open(os.path.join(G, "traces.json.gz"), "rb") as f:
        g = json.loads(f.read())
    envs, steps = {}, 0
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = dropin.make("SPaRc-Gym", puzzles=g["rows"], observation="new", traceback=key[0], max_steps=key[1])
        row = g["rows"][ep["row"]]
        got = run_episode(envs[key], row["id"], ep["actions"])
        d = diff_traces(ep["trace"], json.loads(json.dumps(got)))
        assert d is None, f"{row['id']} traceback={key[0]}: {d}"
        steps += len(got)
    assert steps > 3000


def test_dropin_host_logic_text_mode_and_cursor(dropin):
    with gzip.open(os.path.join(G, "sparc_text.json.gz"), "rb") as f:
        g = json.loads(f.read())
    envs = {}
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = dropin.make("SPaRC-Gym", puzzles=g["rows"], observation="SPaRC", traceback=key[0], max_steps=key[1])
        env = envs[key]
        obs, info = env.reset(options={"puzzle_id": g["rows"][ep["row"]]["id"]})
        assert [obs, None, None, None, info["legal_actions"]] == ep["trace"][0]
        for a, want in zip(ep["actions"], ep["trace"][1:]):
            obs, r, te, tr, info = env.step(a)
            assert [obs, [type(r).__name__, float(r)], te, tr, info["legal_actions"]] == want
    with gzip.open(os.path.join(G, "resets.json.gz"), "rb") as f:
        g = json.loads(f.read())
    env = dropin.SPaRC_Gym(puzzles=g["rows"], max_steps=10)
    for s, want in zip(g["script"], g["index"]):
        if s[0] == "none":
            env.reset()
        elif s[0] == "seed":
            env.reset(seed=s[1])
        elif s[0] == "id":
            env.reset(options={"puzzle_id": s[1]})
        else:
            env.reset(seed=s[1], options={"puzzle_id": s[2]})
        assert env.current_puzzle_index == want, s


def test_dropin_host_logic_config1_rollout(dropin):
    """BASELINE configs[0] through the drop-in (gym.make surface) over the host build of the device code: the first
    3,000 of the 10,000 recorded steps, every output incl. the CRC of the whole observation
    (tests/test_gpu_gym_dropin.py replays all of them on the GPU)."""
    from parity import replay_config1
    with gzip.open(os.path.join(G, "config1.json.gz"), "rb") as f:
        g = json.loads(f.read())
    env = dropin.make("SPaRc-Gym", puzzles=[g["row"]], observation="new", traceback=False, max_steps=g["config"]["max_steps"])
    assert replay_config1(env, g, n_steps=3000) > 30


def test_verify_paths_host_logic(monkeypatch):
    """sparc_gym_b200.verify.verify_paths (the LLM-evaluation caller's env work as one batched call) over the host
    build: listed solutions are solved and satisfy every rule, their prefixes are neither; return, step count and
    flags against the oracle replaying the same actions one env at a time (GPU twin: tests/test_gpu_verify.py)."""
    import sparc_oracle as so
    import sparc_gym_b200.verify as vf
    from sparc_gym_b200.dataset import RULES_FULL, gen_pool
    monkeypatch.setattr(vf, "SparcVecEnv", EmuTorchVecEnv)
    rows = gen_pool(10, [(3, 3), (4, 4), (5, 5)], RULES_FULL, seed0=700, extra_solutions=1)
    idx, paths = [], []
    for i, r in enumerate(rows):
        sol = [(p["x"], p["y"]) for p in r["solutions"][0]["path"]]
        idx += [i, i, i]
        paths += [sol, sol[:max(2, len(sol) // 2)], sol[:3] + [sol[1]] + sol[3:]]
    res = vf.verify_paths(rows, idx, paths, device="cpu")
    assert res["solved"][0::3].all() and res["all_rules_satisfied"][0::3].all() and res["terminated"][0::3].all()
    assert not res["solved"][1::3].any() and not res["reached_target"][1::3].any()
    env = so.OracleEnv(rows, max_steps=10 ** 6)
    for k, (i, p) in enumerate(zip(idx, paths)):
        env.reset(options={"puzzle_id": rows[i]["id"]})
        total, n, te, tr, info = 0, 0, False, False, env._info()
        for a in vf.path_to_actions(p):
            _, r, te, tr, info = env.step(a)
            total += round(r * 100)
            n += 1
            if te or tr:
                break
        assert res["return_cents"][k] == total and res["steps"][k] == n, k
        assert bool(res["terminated"][k]) == te and bool(res["truncated"][k]) == tr, k
        for name in so.RULE_NAMES:
            assert bool(res[name][k]) == info["rule_status"][name]["passed"], (k, name)
