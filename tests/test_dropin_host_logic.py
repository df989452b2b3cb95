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

    def __init__(self, puzzles, num_envs, device=None, **kw):
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
        return tuple(torch.from_numpy(np.ascontiguousarray(o).view(np.int16) if o.dtype == np.uint16 else o)
                     for o in (out if isinstance(out, tuple) else (out,)))

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
