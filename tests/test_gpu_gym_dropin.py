"""GPU: the single-env Gymnasium drop-in reproduces the golden traces recorded from the UNMODIFIED
reference (tests/golden/traces.json.gz) -- dict key order, shapes, dtypes, reward value AND Python
type, legal actions, rule flags, region summaries -- step by step."""
import gzip
import json
import os

import pytest

from tracing import diff_traces, run_episode

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_dropin_replays_reference_traces():
    import sparc_gym_b200
    with gzip.open(os.path.join(G, "traces.json.gz"), "rb") as f:
        g = json.loads(f.read())
    envs, steps = {}, 0
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = sparc_gym_b200.make("SPaRc-Gym", puzzles=g["rows"], observation="new",
                                            traceback=key[0], max_steps=key[1])
        row = g["rows"][ep["row"]]
        got = run_episode(envs[key], row["id"], ep["actions"])
        d = diff_traces(ep["trace"], json.loads(json.dumps(got)))
        assert d is None, f"{row['id']} traceback={key[0]}: {d}"
        steps += len(got)
    assert steps > 3000


def test_reset_cursor_matches_reference():
    import sparc_gym_b200
    with gzip.open(os.path.join(G, "resets.json.gz"), "rb") as f:
        g = json.loads(f.read())
    env = sparc_gym_b200.SPaRC_Gym(puzzles=g["rows"], max_steps=10)
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


def test_dropin_sparc_text_observation_matches_reference():
    """observation='SPaRC': the JSON text grid with its L / V / + / . overlays, step by step."""
    import sparc_gym_b200
    with gzip.open(os.path.join(G, "sparc_text.json.gz"), "rb") as f:
        g = json.loads(f.read())
    envs = {}
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = sparc_gym_b200.make("SPaRC-Gym", puzzles=g["rows"], observation="SPaRC",
                                            traceback=key[0], max_steps=key[1])
        env = envs[key]
        obs, info = env.reset(options={"puzzle_id": g["rows"][ep["row"]]["id"]})
        assert [obs, None, None, None, info["legal_actions"]] == ep["trace"][0]
        for a, want in zip(ep["actions"], ep["trace"][1:]):
            obs, r, te, tr, info = env.step(a)
            assert [obs, [type(r).__name__, float(r)], te, tr, info["legal_actions"]] == want


def test_dropin_config1_rollout_matches_reference():
    """BASELINE configs[0] on the GPU: gym.make(..., observation='new', traceback=False, max_steps=1000) on one 4x4
    gaps+dots puzzle, the recorded 10,000 random steps of the unmodified reference, every output of every step
    (reward value and type, flags, legal actions, agent, step counter, all_rules_satisfied, CRC of the observation)."""
    import gzip, json, os
    import sparc_gym_b200
    from parity import replay_config1
    with gzip.open(os.path.join(os.path.dirname(__file__), "golden", "config1.json.gz"), "rb") as f:
        g = json.loads(f.read())
    env = sparc_gym_b200.make("SPaRc-Gym", puzzles=[g["row"]], observation="new", traceback=False,
                              max_steps=g["config"]["max_steps"])
    assert replay_config1(env, g) == g["golden"]["episodes"] > 100


def test_dropin_steps_past_the_end_like_the_reference():
    """the reference does not guard a finished episode; its recorded outputs for 12 steps past done
    (tests/golden/postdone.json.gz), incl. the sticky outcome reward after a win."""
    import sparc_gym_b200
    with gzip.open(os.path.join(G, "postdone.json.gz"), "rb") as f:
        g = json.loads(f.read())
    envs, steps = {}, 0
    for ep in g["episodes"]:
        key = (ep["traceback"], ep["max_steps"])
        if key not in envs:
            envs[key] = sparc_gym_b200.make("SPaRC-Gym", puzzles=g["rows"], observation="new", traceback=key[0], max_steps=key[1])
        row = g["rows"][ep["row"]]
        got = run_episode(envs[key], row["id"], ep["actions"], stop_at_done=False)
        d = diff_traces(ep["trace"], json.loads(json.dumps(got)))
        assert d is None, f"{row['id']} traceback={key[0]}: {d}"
        steps += len(got)
    assert steps > 1800


@pytest.mark.parametrize("lazy", [False, True])
def test_dropin_info_is_lazy_but_belongs_to_its_step_and_render_is_text(lazy):
    import sparc_gym_b200
    from fuzzrows import solution_actions
    from sparc_gym_b200.dataset import RULES_FULL, gen_puzzle
    row = gen_puzzle(5, 5, 3, RULES_FULL)
    env = sparc_gym_b200.make("SPaRc-Gym", puzzles=[row], render_mode="llm", max_steps=80, lazy_info=lazy)
    obs, info = env.reset(options={"puzzle_id": row["id"]})
    infos = [info]
    for a in solution_actions(row, 0):
        obs, r, te, tr, info = env.step(a)
        infos.append(info)
    assert te and r == 1 and type(r) is int
    assert infos[0]["rule_status"]["reached_target"]["passed"] is False            # filled in only now
    assert infos[-1]["rule_status"]["all_rules_satisfied"]["passed"] is True
    mid = infos[len(infos) // 2]
    assert mid["rule_status"]["reached_target"]["detail"]["agent_loc"] != infos[-1]["rule_status"]["reached_target"]["detail"]["agent_loc"]
    text = env.render()
    assert isinstance(text, str) and text.count("\n") == env.y_size - 1 and "L" in text and "V" in text
