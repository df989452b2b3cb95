"""Generate the golden vectors in this directory by running the UNMODIFIED reference
(/root/reference, imported through oracle/refharness.py).  Run once in the build container:

    python tests/golden/make_golden.py

Outputs (committed; the GPU box has no /root/reference):
  traces.json.gz    rows + action streams + every output of reset/step, step by step
  polyfit.json.gz   direct known answers of the reference's exact-fit search
  resets.json.gz    puzzle-cursor behaviour of reset(seed=...) / reset() / reset(options=...)
  ingest.json.gz    `_process_puzzles` results on rows that exercise the ingest quirks
  sparc_text.json.gz  observation='SPaRC' strings, step by step
  config1.json.gz   BASELINE configs[0]: the 10,000-step random rollout of one 4x4 gaps+dots puzzle, step by step
"""
from __future__ import annotations

import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE)),
                os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle")]

import refharness as rh                                     # noqa: E402
from fuzzrows import SMALL_SHAPES, fuzz_actions, fuzz_row   # noqa: E402
from tracing import _pack, _pack_bits, run_episode          # noqa: E402
from sparc_gym_b200.dataset import RULES_FULL, RULES_GAPS_DOTS, gen_puzzle  # noqa: E402


def dump(name, obj):
    raw = json.dumps(obj, separators=(",", ":")).encode()
    with gzip.GzipFile(os.path.join(HERE, name), "wb", mtime=0) as f:
        f.write(raw)
    print(f"{name}: {len(raw) / 1e6:.2f} MB raw, {os.path.getsize(os.path.join(HERE, name)) / 1e3:.0f} kB gz")


def golden_rows():
    rows = [fuzz_row(s, yaml_style="block" if s % 2 else "json") for s in range(40)]
    rows += [fuzz_row(1000 + s, max_cells=5, n_colors=8, collide=False) for s in range(10)]
    rows += [gen_puzzle(4, 4, 0, RULES_GAPS_DOTS),                  # BASELINE config 1
             gen_puzzle(3, 3, 1, RULES_FULL), gen_puzzle(4, 4, 2, RULES_FULL, extra_solutions=2),
             gen_puzzle(5, 5, 3, RULES_FULL), gen_puzzle(6, 6, 4, RULES_FULL),
             gen_puzzle(6, 6, 5, RULES_FULL, extra_solutions=1), gen_puzzle(6, 4, 6, RULES_FULL),
             gen_puzzle(2, 5, 7, RULES_FULL)]
    return rows


def make_traces():
    rows = golden_rows()
    episodes = []
    for tb in (False, True):
        ref = rh.make_reference_env(rows, observation="new", traceback=tb, max_steps=40)
        for i, row in enumerate(rows):
            planted = row["id"].startswith("syn")
            for ep in range(2):
                if planted and ep == 0:                   # replay the listed solution -> +1
                    from fuzzrows import solution_actions
                    acts = solution_actions(row, 0)[:39]
                else:
                    acts = fuzz_actions(row, 31 * i + ep + 1000 * tb, 45)
                trace = run_episode(ref, row["id"], acts, pre_reset=rh.zero_leak)
                episodes.append({"row": i, "traceback": tb, "max_steps": 40,
                                 "actions": acts[:max(len(trace) - 1, 0)], "trace": trace})
    dump("traces.json.gz", {"rows": rows, "episodes": episodes})


def obs_crc(obs):
    """CRC32 over the observation exactly as the reference holds it: base planes in the puzzle's key order
    (int32[X, Y] each), then color (int32), then additional_info (int64)."""
    import zlib
    c = 0
    for k in obs["base"]:
        c = zlib.crc32(np.ascontiguousarray(obs["base"][k], np.int32).tobytes(), c)
    c = zlib.crc32(np.ascontiguousarray(obs["color"], np.int32).tobytes(), c)
    return zlib.crc32(np.ascontiguousarray(obs["additional_info"], np.int64).tobytes(), c)


CONFIG1 = dict(w=4, h=4, seed=0, T=10000, max_steps=1000, action_seed=0)


def config1_actions():
    return np.random.default_rng(CONFIG1["action_seed"]).integers(0, 4, size=CONFIG1["T"], dtype=np.uint8)


def make_config1():
    """BASELINE configs[0]: one 4x4 gaps+dots puzzle, observation='new', traceback=False, max_steps=1000,
    10,000 random actions, reset onto the same puzzle when an episode ends (the reference's plane leak undone
    first).  Per step: reward (type + value), terminated, truncated, legal actions, agent location,
    current_step, all_rules_satisfied and a CRC32 of the whole observation."""
    row = gen_puzzle(CONFIG1["w"], CONFIG1["h"], CONFIG1["seed"], RULES_GAPS_DOTS)
    ref = rh.make_reference_env([row], observation="new", traceback=False, max_steps=CONFIG1["max_steps"])
    rh.zero_leak(ref)
    obs, info = ref.reset(options={"puzzle_id": row["id"]})
    out = {"reset_crc": obs_crc(obs), "keys": list(obs["base"].keys()), "reward": [], "term": [], "trunc": [],
           "legal": [], "agent": [], "step": [], "rules": [], "crc": [], "episodes": 0}
    for a in config1_actions().tolist():
        obs, r, te, tr, info = ref.step(a)
        out["reward"].append([type(r).__name__, float(r)])
        out["term"].append(int(bool(te))); out["trunc"].append(int(bool(tr)))
        out["legal"].append(sum(1 << int(x) for x in info["legal_actions"]))
        out["agent"].append([int(v) for v in info["agent_location"]])
        out["step"].append(int(info["current_step"]))
        out["rules"].append(int(bool(info["rule_status"]["all_rules_satisfied"]["passed"])))
        out["crc"].append(obs_crc(obs))
        if te or tr:
            out["episodes"] += 1
            rh.zero_leak(ref)
            obs, info = ref.reset(options={"puzzle_id": row["id"]})
            assert obs_crc(obs) == out["reset_crc"]
    dump("config1.json.gz", {"config": CONFIG1, "row": row, "golden": out})


def make_polyfit():
    """Call the reference's `_polyfit_region_exact` directly on stamped regions."""
    rng = np.random.default_rng(7)
    cases = []
    ref_mod = rh.load_reference_class()
    while len(cases) < 400:
        w, h = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        X, Y = 2 * w + 1, 2 * h + 1
        need = np.zeros((w, h), dtype=int)
        polys, ylops = [], []
        for _ in range(int(rng.integers(1, 5))):          # stamp polys (overlaps => multiplicity)
            arr = np.array(SMALL_SHAPES[int(rng.integers(len(SMALL_SHAPES) - 1))])
            if arr.shape[0] > w or arr.shape[1] > h:
                continue
            ox, oy = int(rng.integers(w - arr.shape[0] + 1)), int(rng.integers(h - arr.shape[1] + 1))
            need[ox:ox + arr.shape[0], oy:oy + arr.shape[1]] += arr
            polys.append(arr.tolist())
        for _ in range(int(rng.integers(0, 3))):          # dig ylops out of what is there
            arr = np.array(SMALL_SHAPES[int(rng.integers(4))])
            if arr.shape[0] > w or arr.shape[1] > h:
                continue
            ox, oy = int(rng.integers(w - arr.shape[0] + 1)), int(rng.integers(h - arr.shape[1] + 1))
            sub = need[ox:ox + arr.shape[0], oy:oy + arr.shape[1]]
            if (sub >= arr).all():
                sub -= arr
                ylops.append(arr.tolist())
        region = [(2 * i + 1, 2 * j + 1) for i in range(w) for j in range(h) if need[i, j] > 0]
        if not region or not polys:
            continue
        if rng.random() < 0.3:                            # perturb
            k = int(rng.integers(3))
            if k == 0 and len(region) > 1:
                region.pop(int(rng.integers(len(region))))
            elif k == 1:
                polys[int(rng.integers(len(polys)))] = SMALL_SHAPES[int(rng.integers(len(SMALL_SHAPES)))]
            else:
                polys.append(SMALL_SHAPES[int(rng.integers(3))])
        names = {}
        def nm(a):
            return names.setdefault(json.dumps(a), str(20 + len(names)))
        insts = [{"name": nm(a), "kind": "poly"} for a in polys] + \
                [{"name": nm(a), "kind": "ylop"} for a in ylops]
        shapes = {v: json.loads(k) for k, v in names.items()}
        env = ref_mod.SPaRC_Gym.__new__(ref_mod.SPaRC_Gym)
        env.x_size, env.y_size, env.polyshapes = X, Y, shapes
        reg = ref_mod.RegionData(id=0, cells=list(region), area=len(region), symbols={}, colors={})
        ok, det = env._polyfit_region_exact(reg, insts)
        cases.append({"X": X, "Y": Y, "region": [list(c) for c in region], "shapes": shapes,
                      "instances": insts, "expected": bool(ok),
                      "area_matched": det["net"] == len(region)})
    print("polyfit cases:", len(cases), "fit:", sum(c["expected"] for c in cases),
          "with ylops:", sum(any(i["kind"] == "ylop" for i in c["instances"]) for c in cases),
          "area-matched:", sum(c["area_matched"] for c in cases))
    dump("polyfit.json.gz", cases)


def make_resets():
    rows = [gen_puzzle(2, 2, s, RULES_GAPS_DOTS) for s in range(7)]
    ref = rh.make_reference_env(rows, observation="new", traceback=False, max_steps=10)
    script = [("none",), ("none",), ("seed", 0), ("none",), ("seed", 12345), ("seed", 12345),
              ("id", rows[3]["id"]), ("none",), ("id", "no-such-id"), ("seed+id", 5, rows[1]["id"]),
              ("none",), ("none",), ("none",), ("none",), ("seed", 2 ** 31 + 5)]
    out = []
    for s in script:
        rh.zero_leak(ref)
        if s[0] == "none":
            ref.reset()
        elif s[0] == "seed":
            ref.reset(seed=s[1])
        elif s[0] == "id":
            ref.reset(options={"puzzle_id": s[1]})
        else:
            ref.reset(seed=s[1], options={"puzzle_id": s[2]})
        out.append(int(ref.current_puzzle_index))
    dump("resets.json.gz", {"rows": rows, "script": [list(s) for s in script], "index": out})


def _props_row(cells, w=2, h=2, shapes=None, rid="q"):
    import yaml
    text = {"puzzle": {"start": {"x": 0, "y": 2 * h}, "end": {"x": 2 * w, "y": 0}, "cells": cells}}
    return {"id": rid, "difficulty_level": 1, "grid_size": {"width": w, "height": h},
            "solution_count": 0, "solutions": [], "polyshapes": json.dumps(shapes or {}),
            "text_visualization": yaml.safe_dump(text, sort_keys=False),
            "puzzle_array": [["+"] * (2 * w + 1) for _ in range(2 * h + 1)]}


def ingest_cases():
    """Row lists that poke the ingest quirks Q11-Q13 (SG.py:283-351)."""
    P = lambda x, y, **p: {"position": {"x": x, "y": y}, "properties": p}
    return {
        "plain": [_props_row([P(1, 1, type="star", color="red"), P(3, 3, type="square", color="blue"),
                              P(2, 1, dot=True), P(1, 2, gap=True)])],
        # second puzzle starts with a gap-only cell: inherits the previous puzzle's last symbol (Q12)
        "stale_across_puzzles": [
            _props_row([P(1, 1, type="square", color="red")], rid="a"),
            _props_row([P(1, 2, gap=True), P(2, 1, dot=True)], rid="b")],
        # color listed before type: the stale symbol of the previous cell gets the bit (Q11)
        "color_before_type": [_props_row([P(1, 1, type="star", color="red"),
                                          {"position": {"x": 3, "y": 3},
                                           "properties": {"color": "blue", "type": "square"}}])],
        # unknown colour, colourless star, triangle count 0, count colliding with a shape id (Q10)
        "odd_values": [_props_row([P(1, 1, type="star"), P(1, 3, type="star", color="pink"),
                                   P(3, 1, type="triangle", color="green", count=0),
                                   P(3, 3, type="poly", color="white", polyshape=2)],
                                  shapes={"2": [[1, 1]]})],
        # dot on a typed cell: two keys in one properties dict
        "dot_and_type": [_props_row([{"position": {"x": 1, "y": 1},
                                      "properties": {"type": "square", "color": "red", "dot": True}}])],
        # very first cell of the dataset is gap-only: the reference raises UnboundLocalError
        "gap_first_raises": [_props_row([P(1, 2, gap=True), P(1, 1, type="star", color="red")])],
        "no_cells": [_props_row([])],
    }


def make_ingest():
    out = {}
    for name, rows in ingest_cases().items():
        try:
            env = rh.make_reference_env(rows, observation="new", traceback=False, max_steps=10)
        except Exception as e:  # noqa: BLE001
            out[name] = {"rows": rows, "raises": type(e).__name__}
            continue
        pz = []
        rh.zero_leak(env)
        for p in env.puzzles:
            p["obs_array"]["target_location"][...] = 0      # set by _load_puzzle, not by ingest
            pz.append({"keys": list(p["obs_array"].keys()),
                       "planes": {k: _pack_bits(v) for k, v in p["obs_array"].items()},
                       "color": _pack(p["color_array"]), "additional_info": _pack(p["additional_info"]),
                       "start": list(p["start_location"]), "target": list(p["target_location"]),
                       "solution_count": int(p["solution_count"])})
        out[name] = {"rows": rows, "puzzles": pz}
    dump("ingest.json.gz", out)


def make_sparc_text():
    """observation='SPaRC' (JSON text of the [y][x] symbol grid with L / V / + / . overlays, SG.py:1149-1184)."""
    rows = [gen_puzzle(3, 3, 31, RULES_FULL), gen_puzzle(4, 4, 32, RULES_FULL), gen_puzzle(5, 4, 33, RULES_FULL)]
    episodes = []
    for tb in (False, True):
        ref = rh.make_reference_env(rows, observation="SPaRC", traceback=tb, max_steps=30)
        for i, row in enumerate(rows):
            for ep in range(2):
                acts = fuzz_actions(row, 77 * i + ep + 10 * tb, 35, p_follow=0.5 if ep else 0.95, n_act=4)
                rh.zero_leak(ref)
                obs, info = ref.reset(options={"puzzle_id": row["id"]})
                trace = [[obs, None, None, None, [int(a) for a in info["legal_actions"]]]]
                used = []
                for a in acts:
                    obs, r, te, tr, info = ref.step(a)
                    used.append(a)
                    trace.append([obs, [type(r).__name__, float(r)], bool(te), bool(tr), [int(x) for x in info["legal_actions"]]])
                    if te or tr:
                        break
                episodes.append({"row": i, "traceback": tb, "max_steps": 30, "actions": used, "trace": trace})
    dump("sparc_text.json.gz", {"rows": rows, "episodes": episodes})


def make_postdone():
    """the reference has no guard after done -- step() keeps moving the agent and keeps returning +-1.
    Episodes stepped 12 actions PAST their end (stop_at_done=False), both traceback modes, small max_steps so that
    truncation by step count happens too; the drop-in (autoreset mode CONTINUE) must reproduce every output."""
    rows = golden_rows()[40:]
    episodes = []
    for tb in (False, True):
        ref = rh.make_reference_env(rows, observation="new", traceback=tb, max_steps=14)
        for i, row in enumerate(rows):
            from fuzzrows import solution_actions
            planted = row["id"].startswith("syn")
            for ep in range(2):
                rng = np.random.default_rng([i, ep, int(tb), 77])
                if planted and ep == 0:
                    acts = solution_actions(row, 0)[:13] + [int(a) for a in rng.integers(0, 4, 12)]
                else:
                    acts = fuzz_actions(row, 17 * i + ep + 500 * tb, 26)
                trace = run_episode(ref, row["id"], acts, pre_reset=rh.zero_leak, stop_at_done=False)
                episodes.append({"row": i, "traceback": tb, "max_steps": 14, "actions": acts, "trace": trace})
    dump("postdone.json.gz", {"rows": rows, "episodes": episodes})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "postdone":      # (added in round 2; the other files are unchanged)
        make_postdone()
        sys.exit(0)
    if not rh.reference_available():
        sys.exit("reference not present; golden vectors can only be regenerated in the build container")
    make_sparc_text()
    make_ingest()
    make_resets()
    make_polyfit()
    make_traces()
    make_config1()
    make_postdone()
