"""Record everything an env (reference, oracle or the CUDA drop-in) returns, step by step, in
a JSON-able form, so traces can be stored as golden vectors and compared exactly."""
from __future__ import annotations

import numpy as np

RULE_NAMES = ("reached_target", "path_not_crossing", "no_gap_violations", "all_dots_collected",
              "square_color_separation", "star_pairing_exact", "triangles_edge_count",
              "poly_ylop_area", "all_rules_satisfied")


def _pack(arr):
    a = np.asarray(arr)
    return [list(a.shape), str(a.dtype), a.astype(np.int64).ravel().tolist()]


def _pack_bits(arr):
    a = np.asarray(arr)
    assert set(np.unique(a).tolist()) <= {0, 1}, "base planes must be one-hot"
    return [list(a.shape), str(a.dtype), np.packbits(a.astype(np.uint8).ravel()).tobytes().hex()]


def _reward_repr(r):
    # value AND Python type matter (the reference returns int 0/1/-1 and float +-0.01)
    return [type(r).__name__, float(r)]


def _canon(v):
    """JSON-able, type-insensitive form of a rule `detail` payload: numpy scalars -> int, tuples -> lists,
    dict keys -> str; container order is kept (it is part of what is compared)."""
    if isinstance(v, dict):
        return {str(_canon_key(k)): _canon(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_canon(x) for x in v]
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    return v


def _canon_key(k):
    return tuple(int(x) for x in k) if isinstance(k, tuple) else (int(k) if isinstance(k, (int, np.integer)) else k)


def snapshot(obs, info, reward=None, terminated=None, truncated=None, full_static=False):
    """One step's outputs.  Static planes (everything but visited / agent_location) are
    recorded only when `full_static` (at reset) -- they are still *checked* every step by
    `static_digest`."""
    base = obs["base"]
    s = {
        "keys": list(base.keys()),
        "visited": _pack_bits(base["visited"]),
        "agent_location": _pack_bits(base["agent_location"]),
        "static_digest": static_digest(obs),
        "legal": [int(a) for a in info["legal_actions"]],
        "step": int(info["current_step"]),
        "agent": [int(v) for v in info["agent_location"]],
        "rewards": [_reward_repr(info["Rewards"]["normal_reward"]),
                    _reward_repr(info["Rewards"]["outcome_reward"])],
        "flags": [bool(info["rule_status"][k]["passed"]) for k in RULE_NAMES],
        "details": {k: _canon(info["rule_status"][k]["detail"]) for k in RULE_NAMES + ("_terminated", "_truncated")},
        "regions": {str(k): [int(v["area"]), {kk: int(vv) for kk, vv in v["symbol_counts"].items()},
                             {str(int(kk)): int(vv) for kk, vv in v["colors"].items()}]
                    for k, v in info["rule_status"]["_regions"].items()},
        "meta": [int(info["solution_count"]), int(info["difficulty"]), int(info["grid_x_size"]),
                 int(info["grid_y_size"])],
    }
    if reward is not None:
        s["reward"] = _reward_repr(reward)
        s["terminated"] = bool(terminated)
        s["truncated"] = bool(truncated)
    if full_static:
        s["planes"] = {k: _pack_bits(v) for k, v in base.items()}
        s["color"] = _pack(obs["color"])
        s["additional_info"] = _pack(obs["additional_info"])
    return s


def static_digest(obs):
    """Cheap exact digest of the planes that never change during an episode."""
    import hashlib
    h = hashlib.sha1()
    for k, v in obs["base"].items():
        if k in ("visited", "agent_location"):
            continue
        h.update(k.encode())
        h.update(np.ascontiguousarray(v).astype("<i4").tobytes())
    h.update(np.ascontiguousarray(obs["color"]).astype("<i4").tobytes())
    h.update(np.ascontiguousarray(obs["additional_info"]).astype("<i8").tobytes())
    h.update(str((obs["color"].dtype, obs["additional_info"].dtype,
                  [str(v.dtype) for v in obs["base"].values()])).encode())
    return h.hexdigest()[:16]


def run_episode(env, puzzle_id, actions, pre_reset=None, stop_at_done=True):
    """reset(options={'puzzle_id'}) then step through `actions`; returns the list of snapshots."""
    if pre_reset is not None:
        pre_reset(env)
    obs, info = env.reset(options={"puzzle_id": puzzle_id})
    out = [snapshot(obs, info, full_static=True)]
    for a in actions:
        obs, r, te, tr, info = env.step(a)
        out.append(snapshot(obs, info, r, te, tr))
        if stop_at_done and (te or tr):
            break
    return out


def diff_traces(a, b, path=""):
    """First difference between two traces (nested lists / dicts), or None."""
    if type(a) is not type(b):
        return f"{path}: type {type(a).__name__} != {type(b).__name__} ({a!r} vs {b!r})"
    if isinstance(a, dict):
        if list(a.keys()) != list(b.keys()):
            return f"{path}: keys {list(a.keys())} != {list(b.keys())}"
        for k in a:
            d = diff_traces(a[k], b[k], f"{path}.{k}")
            if d:
                return d
        return None
    if isinstance(a, list):
        if len(a) != len(b):
            return f"{path}: len {len(a)} != {len(b)}"
        for i, (x, y) in enumerate(zip(a, b)):
            d = diff_traces(x, y, f"{path}[{i}]")
            if d:
                return d
        return None
    return None if a == b else f"{path}: {a!r} != {b!r}"
