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


def test_dropin_steps_past_the_end_like_the_reference(dropin):
    """SURVEY Q4 / ADVICE: step() does not guard a finished episode -- the reference's recorded outputs for 12 steps
    past done (autoreset mode 'continue' on the device)."""
    with gzip.open(os.path.join(G, "postdone.json.gz"), "rb") as f:
        g = json.loads(f.read())
    envs = {}
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = dropin.make("SPaRC-Gym", puzzles=g["rows"], observation="new", traceback=key[0], max_steps=key[1])
        row = g["rows"][ep["row"]]
        got = run_episode(envs[key], row["id"], ep["actions"], stop_at_done=False)
        d = diff_traces(ep["trace"], json.loads(json.dumps(got)))
        assert d is None, f"{row['id']} traceback={key[0]}: {d}"


@pytest.mark.parametrize("lazy", [False, True])
def test_dropin_lazy_info_and_text_render(dropin, capsys, lazy):
    """`info` is the reference's plain dict by default; with lazy_info=True it behaves like one (key order, equality,
    json) although rule_status is filled in on first use -- from a snapshot of ITS step, not of the env's current
    state; render_mode is accepted as in the reference and render() draws text."""
    from sparc_gym_b200.dataset import RULES_FULL, gen_puzzle
    from fuzzrows import solution_actions
    row = gen_puzzle(4, 4, 2, RULES_FULL)
    env = dropin.make("SPaRc-Gym", puzzles=[row], render_mode="llm", max_steps=50, lazy_info=lazy)
    obs, info0 = env.reset(options={"puzzle_id": row["id"]})
    assert (type(info0) is dict) == (not lazy)
    assert list(info0.keys()) == ["solution_count", "difficulty", "grid_x_size", "grid_y_size", "legal_actions", "current_step",
                                  "agent_location", "rule_status", "Rewards"]
    infos = [info0]
    for a in solution_actions(row, 0):
        obs, r, te, tr, info = env.step(a)
        infos.append(info)
    assert te and r == 1 and type(r) is int
    # looked at only now, after the episode moved on: every info still describes its own step
    assert infos[0]["rule_status"]["reached_target"]["passed"] is False
    assert infos[-1]["rule_status"]["all_rules_satisfied"]["passed"] is True
    vis = [i["rule_status"]["all_dots_collected"]["detail"]["collected"] for i in infos]
    assert vis == sorted(vis) and vis[0] <= vis[-1]
    assert infos[-1]["rule_status"]["reached_target"]["detail"]["agent_loc"] == list(map(int, env._target_location))
    assert env.rule_status is infos[-1]["rule_status"]
    assert [k for k, _ in infos[3].items()][7] == "rule_status" and infos[3]["rule_status"]["_terminated"] == {"passed": True, "detail": False}
    assert dict(infos[2]) == infos[2] and "rule_status" in infos[2] and len(infos[2]) == 9
    text = env.render()
    assert isinstance(text, str) and text.count("\n") == env.y_size - 1 and "L" in text and "V" in text
    human = dropin.make("SPaRc-Gym", puzzles=[row], render_mode="human")
    assert human.render() is None and "L" in capsys.readouterr().out
    assert dropin.make("SPaRc-Gym", puzzles=[row], render_mode="something else").render() is None


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


def _dataset_files(tmp_path):
    """The same rows as .json, .jsonl and .parquet, in the HF dataset's schema (SG.py:239-363)."""
    import pandas as pd
    from fuzzrows import fuzz_row
    from sparc_gym_b200.dataset import RULES_FULL, gen_pool
    rows = gen_pool(10, [(3, 3), (4, 4), (6, 6), (2, 5)], RULES_FULL, seed0=900, extra_solutions=1) + [fuzz_row(s) for s in range(850, 858)]
    from sparc_gym_b200.dataset import RULES_GAPS_DOTS, gen_puzzle
    bad = gen_puzzle(9, 2, 5, RULES_GAPS_DOTS)                        # 19 x 5 lattice: beyond the 15 x 15 boards
    bad["id"] = "too-wide"
    paths = {}
    for ext in ("json", "jsonl", "parquet"):
        p = tmp_path / f"sparc.{ext}"
        if ext == "json":
            p.write_text(json.dumps(rows + [bad]))
        elif ext == "jsonl":
            p.write_text("\n".join(json.dumps(r) for r in rows + [bad]) + "\n")
        else:
            pd.DataFrame([{k: v for k, v in r.items() if k != "puzzle_array"} for r in rows + [bad]]).to_parquet(p)
        paths[ext] = p
    return rows, paths


def test_dataset_files_and_all_rows_sweep(monkeypatch, tmp_path):
    """dataset.load_rows on JSON / JSONL / parquet files in the dataset's schema gives the same pool, and
    tools/verify_dataset.sweep (ingest every row, replay every listed solution in one batched call) reports the row the
    store refuses and +1 for every listed solution of the others (over the host build of the device code here; GPU twin:
    tests/test_gpu_verify.py)."""
    import importlib.util
    import sparc_gym_b200.verify as vf
    from sparc_gym_b200.dataset import load_rows
    from sparc_gym_b200.ingest import ingest_rows
    from sparc_gym_b200.store import PuzzleStore
    monkeypatch.setattr(vf, "SparcVecEnv", EmuTorchVecEnv)
    spec = importlib.util.spec_from_file_location("verify_dataset", os.path.join(os.path.dirname(G), "..", "tools", "verify_dataset.py"))
    vd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vd)
    rows, paths = _dataset_files(tmp_path)
    want = PuzzleStore(ingest_rows(rows))
    for ext, p in paths.items():
        back = load_rows(p)
        assert len(back) == len(rows) + 1
        got = PuzzleStore(ingest_rows(back[:-1]))
        for name in ("headers", "node_boards", "cell_boards", "static_bits", "color4", "info", "shapes", "trie"):
            assert np.array_equal(getattr(want, name), getattr(got, name)), (ext, name)
        rep = vd.sweep(back, device="cpu")
        assert rep["rows"] == len(rows) + 1 and rep["accepted"] == len(rows)
        assert [r["id"] for r in rep["rejected"]] == ["too-wide"] and "15" in rep["rejected"][0]["error"]
        n_sol = sum(int(r["solution_count"]) for r in rows)
        # every listed solution of the generated rows ends on the target and earns +1; the fuzzed rows also list walks
        # that stop short of it, which the sweep reports
        assert rep["listed_solutions"] == n_sol and rep["solutions_rewarded_plus_one"] >= 20
        assert all(f["id"].startswith("fuzz") for f in rep["reward_failures"])
        assert rep["solutions_rewarded_plus_one"] + len(rep["reward_failures"]) == n_sol
        assert rep["solutions_satisfying_all_rules"] >= 10           # the planted solutions of the generated rows
