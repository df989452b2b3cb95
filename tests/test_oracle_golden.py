"""Pin the CPU oracle (oracle/sparc_oracle.py) to golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only."""
import gzip
import json
import os

import numpy as np
import pytest

import sparc_oracle as so
from tracing import _pack, _pack_bits, diff_traces, run_episode

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with gzip.open(os.path.join(G, name), "rb") as f:
        return json.loads(f.read())


def test_traces_match_reference():
    g = load("traces.json.gz")
    envs = {}
    n_steps = 0
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = so.OracleEnv(g["rows"], observation="new", traceback=key[0], max_steps=key[1])
        row = g["rows"][ep["row"]]
        got = run_episode(envs[key], row["id"], ep["actions"])
        d = diff_traces(ep["trace"], json.loads(json.dumps(got)))
        assert d is None, f"{row['id']} traceback={key[0]}: {d}"
        n_steps += len(got)
    assert n_steps > 3000
    # known answer: replaying a listed solution ends with +1 (SG.py:1205-1209)
    wins = [ep for ep in g["episodes"] if ep["trace"][-1].get("reward") == ["int", 1.0]]
    assert len(wins) >= 8


def test_stepping_past_the_end_matches_reference():
    """no guard after done -- the reference's own outputs for up to 12 steps past the end of an episode."""
    g = load("postdone.json.gz")
    envs, post = {}, 0
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = so.OracleEnv(g["rows"], observation="new", traceback=key[0], max_steps=key[1])
        row = g["rows"][ep["row"]]
        got = run_episode(envs[key], row["id"], ep["actions"], stop_at_done=False)
        d = diff_traces(ep["trace"], json.loads(json.dumps(got)))
        assert d is None, f"{row['id']} traceback={key[0]}: {d}"
        done = [s["terminated"] or s["truncated"] for s in ep["trace"][1:]]
        post += sum(done[1:][i] is not None for i in range(len(done) - 1) if any(done[:i + 1]))
    assert post > 900


def test_polyfit_known_answers():
    cases = load("polyfit.json.gz")
    assert sum(c["expected"] for c in cases) > 50 and sum(not c["expected"] for c in cases) > 50
    for i, c in enumerate(cases):
        polys = [(x["name"], c["shapes"][x["name"]]) for x in c["instances"] if x["kind"] == "poly"]
        ylops = [(x["name"], c["shapes"][x["name"]]) for x in c["instances"] if x["kind"] == "ylop"]
        got = so.exact_fit(c["X"], c["Y"], [tuple(r) for r in c["region"]], polys, ylops)
        assert got == c["expected"], f"case {i}"


def test_reset_cursor():
    g = load("resets.json.gz")
    env = so.OracleEnv(g["rows"], max_steps=10)
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


def test_ingest_quirks():
    g = load("ingest.json.gz")
    assert g["gap_first_raises"]["raises"] == "UnboundLocalError"
    for name, case in g.items():
        if "raises" in case:
            with pytest.raises(Exception) as ei:
                so.process_rows(case["rows"])
            assert type(ei.value).__name__ == case["raises"], name
            continue
        got = so.process_rows(case["rows"])
        assert len(got) == len(case["puzzles"])
        for p, want in zip(got, case["puzzles"]):
            assert list(p["obs_array"].keys()) == want["keys"], name
            for k in want["keys"]:
                assert _pack_bits(p["obs_array"][k]) == want["planes"][k], (name, k)
            assert _pack(p["color_array"]) == want["color"], name
            assert _pack(p["additional_info"]) == want["additional_info"], name
            assert list(p["start_location"]) == want["start"] and list(p["target_location"]) == want["target"]


def test_state_leak_mode_reproduces_q1():
    """clean_reset=False aliases the cached planes like the reference (SG.py:149-151)."""
    g = load("resets.json.gz")
    env = so.OracleEnv(g["rows"][:1], clean_reset=False, max_steps=10)
    env.reset(options={"puzzle_id": g["rows"][0]["id"]})
    for a in (0, 1, 2, 3, 0, 1):
        env.step(a)
    moved = int(env.obs_array["visited"].sum())
    obs, _ = env.reset(options={"puzzle_id": g["rows"][0]["id"]})
    if moved > 1:
        assert int(obs["base"]["visited"].sum()) == moved   # stale cells survive the reset
    clean = so.OracleEnv(g["rows"][:1], max_steps=10)
    obs, _ = clean.reset(options={"puzzle_id": g["rows"][0]["id"]})
    assert int(obs["base"]["visited"].sum()) == 1 and int(obs["base"]["agent_location"].sum()) == 1


def test_sparc_text_observation_matches_reference():
    g = load("sparc_text.json.gz")
    envs = {}
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = so.OracleEnv(g["rows"], observation="SPaRC", traceback=key[0], max_steps=key[1])
        env = envs[key]
        obs, info = env.reset(options={"puzzle_id": g["rows"][ep["row"]]["id"]})
        assert [obs, None, None, None, info["legal_actions"]] == ep["trace"][0]
        for a, want in zip(ep["actions"], ep["trace"][1:]):
            obs, r, te, tr, info = env.step(a)
            assert [obs, [type(r).__name__, float(r)], te, tr, info["legal_actions"]] == want


def test_config1_rollout_matches_reference():
    """BASELINE configs[0]: 10,000 random steps on one 4x4 gaps+dots puzzle (observation='new', traceback=False,
    max_steps=1000), every step against the recording of the unmodified reference."""
    from parity import replay_config1
    g = load("config1.json.gz")
    env = so.OracleEnv([g["row"]], observation="new", traceback=False, max_steps=g["config"]["max_steps"])
    assert replay_config1(env, g) == g["golden"]["episodes"] > 100
