"""Single-env Gymnasium drop-in: `gym.make('SPaRC-Gym', observation='new', traceback=..., max_steps=...)`.

Mirrors the reference class `SPaRC_Gym(gym.Env)` (reference SPaRC_Gym/SPaRC_Gym.py:44-1265, "SG.py"):
same constructor keywords (SG.py:46), `reset(seed=None, options={'puzzle_id': ...})` with the
reference's puzzle-cursor rules (SG.py:1075-1087), `step(action)` returning
`(obs, reward, terminated, truncated, info)` with the reference's Python types -- rewards are the
ints `0 / 1 / -1` and the floats `0.01 / -0.01`, the observation is the dict
`{'base': {plane: int32[X, Y]} (per-puzzle key order), 'color': int32[X, Y], 'additional_info':
int64[X, Y]}` and the arrays are owned by the env and updated in place (SG.py:979).

All dynamics, rewards, rule validation and the observation planes come from the CUDA library
(one env on the GPU); this class only converts device arrays into the reference's containers.
`info['rule_status'][rule]['detail']` (SG.py:488-710) is assembled on the host from what the
device returns: the rule flags, the region map and the per-region poly/ylop verdicts.
observation='SPaRC' (the JSON text grid, SG.py:1149-1184) is assembled on the host from the
device's dynamics.

One step = ONE call into the library (sparc_step_info_host: action in, dynamics, validation of the new state, one
packed record out); the numpy planes are touched in place where the agent moved (as SG.py:1146-1188 does) instead
of being copied from the device.  `info` is the reference's plain dict; with the opt-in `lazy_info=True` its
`rule_status` -- nine `detail` payloads and the region summaries, more Python time than the rest of the step -- is
filled in from a snapshot of that step when it is first looked up (`_LazyInfo`).

Deliberate differences from the reference:
* resets are clean (the reference leaks `visited` / `agent_location` across episodes of one puzzle);
* there are no pygame windows: `render_mode` is accepted as in the reference ('human', 'llm', anything else means no
  renderer, SG.py:68-74), `render()` draws the board as text -- printed for 'human', returned for 'llm' (the
  reference's render() docstring promises a text for 'llm'; its pygame renderers return None);
* like the reference, step() does NOT guard a finished episode (Q4): the agent keeps moving and every further step
  returns +-1 (tests/golden/postdone.json.gz, recorded from the reference).
"""
from __future__ import annotations

import numpy as np

from . import abi
from .ingest import FIXED_KEYS, ingest_rows
from .store import PuzzleStore

try:  # subclass the real thing when it exists; the GPU boxes do not have gymnasium
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _Env = _gym.Env
except Exception:  # pragma: no cover - exercised on the boxes without gymnasium
    _gym = None

    class _Env:  # minimal duck-typed stand-in for gymnasium.Env
        metadata: dict = {}
        np_random = None

        def reset(self, *, seed=None, options=None):
            if seed is not None or self.np_random is None:
                self.np_random = np.random.default_rng(seed)

        def close(self):
            pass

    class _Space:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k

    class _spaces:  # type: ignore[no-redef]
        Dict = type("Dict", (_Space,), {})
        Box = type("Box", (_Space,), {})
        Discrete = type("Discrete", (_Space,), {})
        Text = type("Text", (_Space,), {})

RULE_NAMES = ("reached_target", "path_not_crossing", "no_gap_violations", "all_dots_collected",
              "square_color_separation", "star_pairing_exact", "triangles_edge_count",
              "poly_ylop_area", "all_rules_satisfied")                         # SG.py:912-935


def _load_rows(df_name, df_split, df_set):
    """The reference downloads its puzzles from the HF hub (SG.py:77-78)."""
    from datasets import load_dataset
    return load_dataset(df_name, df_split, split=df_set).to_pandas().to_dict("records")


class _LazyInfo(dict):
    """`info` of one step with `lazy_info=True` (an opt-in of this implementation, not the reference's behaviour):
    'rule_status' is computed from the step's snapshot the first time it is looked up -- the nine `detail` payloads and
    the region summaries cost more Python time than the rest of the step.  Lookups (`info['rule_status']`, `.get`,
    `.items()`, `.values()`, `==`, `repr`, `copy`, pickling) fill it in; C-level copies of a dict subclass
    (`dict(info)`, `{**info}`) read the storage directly and would see the placeholder -- call `info.materialise()`
    first, or leave `lazy_info` off: the default `info` is the reference's plain dict."""
    _PENDING = object()

    def __init__(self, items, snap):
        super().__init__(items)
        self._snap = snap

    def materialise(self):
        snap, self._snap = self._snap, None
        if snap is not None:
            dict.__setitem__(self, "rule_status", snap.rule_status())
        return self

    def __getitem__(self, key):
        if key == "rule_status" and self._snap is not None:
            self.materialise()
        return dict.__getitem__(self, key)

    def get(self, key, default=None):
        if key == "rule_status" and self._snap is not None:
            self.materialise()
        return dict.get(self, key, default)

    def _all(name):
        def method(self, *a, **k):
            if self._snap is not None:
                self.materialise()
            return getattr(dict, name)(self, *a, **k)
        method.__name__ = name
        return method

    for _n in ("values", "items", "__eq__", "__ne__", "__repr__", "pop", "popitem", "setdefault", "__or__", "__ror__"):
        locals()[_n] = _all(_n)
    del _n, _all

    def copy(self):
        return dict(self.materialise().items())

    def __reduce_ex__(self, protocol):
        return (dict, (dict(self.materialise().items()),))


class _Details:
    """What `info['rule_status']` is made of (SG.py:456-710, 901-950), as a function of the attributes below -- of the
    env itself or of a per-step snapshot of them (`_Snap`)."""

    def rule_status(self):
        area, symbols, colors = self._region_symbols()
        details = self._rule_details(area, symbols)
        core = list(RULE_NAMES[:8])
        status = {k: {"passed": self._flags[k], "detail": details[k]} for k in core}
        status["all_rules_satisfied"] = {"passed": self._flags["all_rules_satisfied"],
                                         "detail": {"rules_checked": core}}            # SG.py:931-935
        status["_terminated"] = {"passed": True, "detail": False}              # SG.py:1011, 937-938 (Q16)
        status["_truncated"] = {"passed": True, "detail": False}
        status["_regions"] = {r: {"id": r, "area": area[r],                     # SG.py:949, 36-42
                                  "symbol_counts": {k: len(v) for k, v in symbols[r].items()},
                                  "colors": colors[r]} for r in range(len(area))}
        return status

    def _static_facts(self):
        """Everything `rule_status` needs that does not change while a puzzle is loaded, gathered once per load (the
        containers are fresh per load, SG.py:149-151): symbol cells in the order SG.py:456-482 meets them, dots,
        triangles, shape instances.  Looking these up in the numpy planes every step was most of the step's host time."""
        st = getattr(self, "_static", None)
        if st is not None and st["planes"] is self.obs_array:
            return st
        planes, color, info = self.obs_array, self.color_array, self.additional_info
        cells = []
        for layer, arr in planes.items():
            if layer in FIXED_KEYS:
                continue
            for x, y in zip(*np.where(arr == 1)):
                cells.append((layer, int(x), int(y), color[x, y]))
        st = {"planes": planes, "cells": cells}
        st["dots"] = [(int(x), int(y)) for x, y in zip(*np.where(planes["dot"] == 1))] if "dot" in planes else None
        sx, sy = self.start_location                          # the only path node that can sit on a gap (SG.py:508-519)
        st["gap_violations"] = [(int(sx), int(sy))] if planes["gaps"][sx, sy] == 1 else []
        st["triangles"] = None
        if "triangle" in planes:                                                        # SG.py:617-646
            tri = planes["triangle"]
            st["triangles"] = [(x, y, int(info[x, y])) for x in range(1, self.x_size - 1) for y in range(1, self.y_size - 1)
                               if tri[x, y] == 1 and int(info[x, y]) > 0]
        inst = []                                                                       # SG.py:648-736
        if isinstance(self.polyshapes, dict):
            for x in range(self.x_size):
                for y in range(self.y_size):
                    v = info[x, y]
                    if v != 0 and f"{v}" in self.polyshapes:
                        kind = "poly" if planes["poly"][x, y] == 1 else "ylop"
                        inst.append((x, y, kind, int(np.array(self.polyshapes[f"{v}"]).sum())))
        st["instances"] = inst
        st["has_instances"] = self._has_instances()
        st["target"] = self._target_location.tolist()
        self._static = st
        return st

    def _region_symbols(self):
        """Per region id: area, symbols[layer] = [(x, y), ...] and colour counts (SG.py:456-482), from the
        device's region map and the puzzle's symbol cells."""
        rm = self._region_map.tolist()
        self._rm_rows = rm
        area = np.bincount(self._region_map.ravel().astype(np.int64) + 1).tolist()[1:]    # region -1 = the path itself
        n = len(area)
        symbols = [dict() for _ in range(n)]
        colors = [dict() for _ in range(n)]
        for layer, x, y, c in self._static_facts()["cells"]:
            rid = rm[x][y]
            if rid == -1:
                continue
            sym = symbols[rid]
            if layer in sym:
                sym[layer].append((x, y))
            else:
                sym[layer] = [(x, y)]
            if c:
                col = colors[rid]
                col[c] = col.get(c, 0) + 1
        return area, symbols, colors

    def _rule_details(self, area, symbols):
        """The `detail` payloads of SG.py:488-710.  Every fact comes from the device (flags, region map,
        per-region poly/ylop verdicts, visited plane); this only arranges them in the reference's containers."""
        planes, color = self.obs_array, self.color_array
        st, rm = self._static_facts(), self._rm_rows
        vis = planes["visited"].tolist()
        d = {}
        d["reached_target"] = {"agent_loc": self._agent_location.tolist(),              # SG.py:493-496
                               "target_loc": list(st["target"])}
        d["path_not_crossing"] = {"duplicates": {}}          # path nodes are distinct (clean resets, legal moves)
        d["no_gap_violations"] = {"violations": list(st["gap_violations"])}
        if st["dots"] is not None:                                                      # SG.py:521-532
            d["all_dots_collected"] = {"total": len(st["dots"]), "collected": sum(1 for (x, y) in st["dots"] if vis[x][y] == 1)}
        else:
            d["all_dots_collected"] = {"total": 0, "collected": 0}
        if "square" in planes:                                                          # SG.py:534-552
            bad, det = [], []
            for rid, sym in enumerate(symbols):
                squares = sym.get("square", [])
                if not squares:
                    continue
                cs = set(color[x, y] for (x, y) in squares if color[x, y] != 0)
                if len(cs) > 1:
                    bad.append(rid)
                det.append({"region": rid, "square_count": len(squares), "colors": list(cs)})
            d["square_color_separation"] = {"violating_regions": bad, "region_square_details": det}
        else:
            d["square_color_separation"] = {"regions": []}
        if "star" in planes:                                                            # SG.py:554-614
            violations, per_region = [], []
            for rid, sym in enumerate(symbols):
                stars = sym.get("star", [])
                if not stars:
                    continue
                totals = {}
                for coords in sym.values():
                    for (x, y) in coords:
                        c = color[x, y]
                        if c != 0:
                            totals[c] = totals.get(c, 0) + 1
                star_colors = {}
                for (x, y) in stars:
                    c = color[x, y]
                    if c == 0:
                        violations.append({"region": rid, "color": 0, "found_total": 1})
                        continue
                    star_colors[c] = star_colors.get(c, 0) + 1
                region_ok, rows = True, []
                for c, n_star in star_colors.items():
                    total_c = totals.get(c, 0)
                    ok = total_c == 2
                    if not ok:
                        region_ok = False
                        violations.append({"region": rid, "color": c, "found_total": total_c, "star_cells": n_star})
                    rows.append({"color": c, "total_symbols_of_color": total_c, "star_cells": n_star, "ok": ok})
                per_region.append({"region": rid, "details": rows, "all_ok": region_ok})
            d["star_pairing_exact"] = {"violations": violations, "per_region": per_region}
        else:
            d["star_pairing_exact"] = {"regions": []}
        mism = []
        if st["triangles"] is not None:                                                 # SG.py:617-646
            for x, y, need in st["triangles"]:
                touches = vis[x + 1][y] + vis[x - 1][y] + vis[x][y - 1] + vis[x][y + 1]
                if touches != need:
                    mism.append({"x": x, "y": y, "required": need, "touches": touches})
        d["triangles_edge_count"] = {"mismatches": mism}
        by_region = {}                                                                  # SG.py:648-736
        for x, y, kind, shape_area in st["instances"]:
            rid = rm[x][y]
            if rid != -1:
                by_region.setdefault(rid, []).append((kind, shape_area))
        if not by_region and not st["has_instances"]:
            d["poly_ylop_area"] = {"regions": []}
        else:
            det = []
            for rid, insts in by_region.items():
                pa = sum(a for k, a in insts if k == "poly")
                ya = sum(a for k, a in insts if k == "ylop")
                net, verdict = pa - ya, int(self._region_fit[rid])
                area_ok = area[rid] == net
                assert area_ok == (verdict in (abi.FIT_OK, abi.FIT_NO_TILING)), "device / host area check disagree"
                row = {"region": rid, "area_check": {"region_area": area[rid], "poly_area": pa, "ylop_area": ya,
                                                     "net": net, "ok": area_ok}}
                if area_ok:
                    row["exact_fit"] = {"ok": verdict == abi.FIT_OK, "region_id": rid, "region_area": area[rid],
                                        "poly_area": pa, "ylop_area": ya, "net": net}
                else:
                    row["exact_fit"] = {"ok": False, "skipped": True}
                row["ok"] = area_ok and verdict == abi.FIT_OK
                det.append(row)
            d["poly_ylop_area"] = {"violations": [r["region"] for r in det if not r["ok"]], "region_details": det}
        return d

    def _has_instances(self):
        """SG.py:716-736: is any node (on a rule cell or not) a poly / ylop instance?"""
        if not isinstance(self.polyshapes, dict):
            return False
        nz = self.additional_info[self.additional_info != 0]
        return any(f"{v}" in self.polyshapes for v in nz)


class _Snap(_Details):
    """The state `rule_status` is a function of, frozen at one step: the static planes by reference (a puzzle load makes
    fresh ones), the two dynamic planes and the device's verdicts by value."""

    def __init__(self, env):
        self.obs_array = dict(env.obs_array)
        self.obs_array["visited"] = env.obs_array["visited"].copy()
        self.obs_array["agent_location"] = env.obs_array["agent_location"].copy()
        self.color_array, self.additional_info = env.color_array, env.additional_info
        self._agent_location, self._target_location = env._agent_location.copy(), env._target_location
        self.start_location, self.x_size, self.y_size = env.start_location, env.x_size, env.y_size
        self.polyshapes, self._flags = env.polyshapes, env._flags
        self._region_map, self._region_fit = env._region_map, env._region_fit
        st = dict(env._static_facts())
        st["planes"] = self.obs_array                      # (this snapshot's own dict; same static planes)
        self._static = st


class SPaRC_Gym(_Details, _Env):
    metadata = {"render_modes": ["human", "llm"], "render_fps": 30}            # SG.py:45

    def __init__(self, df_name="lkaesberg/SPaRC", df_split="all", df_set="test", render_mode=None,
                 observation="new", traceback=False, max_steps=2000, *, puzzles=None, device="cuda:0", lazy_info=False):
        if observation not in ("new", "SPaRC"):
            raise ValueError("Invalid observation type. Choose 'new' or 'SPaRC'.")     # SG.py:207
        from .vec_env import SparcVecEnv
        self.render_mode, self.observation, self.lazy_info = render_mode, observation, bool(lazy_info)
        self.traceback, self.max_steps = traceback, max_steps
        rows = puzzles if puzzles is not None else _load_rows(df_name, df_split, df_set)
        if rows is None or len(rows) == 0:
            raise ValueError("No valid dataframe provided")                    # SG.py:86-87
        self.puzzles = ingest_rows(rows, observation)
        self._store = PuzzleStore(self.puzzles)
        # 'continue': a finished episode is stepped like a live one, as in the reference (Q4); every call validates
        # the state it leaves behind (sparc_step_info_host), so the step itself need not
        self._vec = SparcVecEnv(self._store, 1, traceback=traceback, max_steps=max_steps,
                                autoreset_mode="continue", validate="never", device=device)
        self._fast = hasattr(self._vec, "step_info_host")      # (the CPU twin used by the host-logic tests has no such call)
        self._last_info = None
        self.current_puzzle_index = 0
        self.current_step = 0
        self.rule_status = {}
        self.action_space = _spaces.Discrete(4)                                # SG.py:210
        self._load_puzzle(0)                                                   # SG.py:90

    # -- device -> reference containers
    def _load_puzzle(self, index):
        import torch
        pz = self.puzzles[index]
        self._pz = pz
        self.x_size, self.y_size = pz.X, pz.Y
        self.difficulty, self.polyshapes = pz.difficulty, pz.polyshapes
        self.solution_paths, self.solution_count = pz.solution_paths, pz.solution_count
        self.start_location, self.target_location = pz.start, pz.target
        self.normal_reward = 0
        self.outcome_reward = 0
        rec = self._vec.reset_info_host([index]) if self._fast else None
        if rec is None:
            self._vec.reset(puzzle_idx=torch.tensor([index], dtype=torch.int32))
        # fresh containers per load, then updated in place while the episode runs (SG.py:149-151, 979)
        self.obs_array = {k: np.zeros((pz.X, pz.Y), np.int32) for k in pz.keys}
        self.color_array = np.zeros((pz.X, pz.Y), np.int32)
        self.additional_info = np.zeros((pz.X, pz.Y), np.int64)
        self._agent_location = np.zeros(2, np.int32)
        self._target_location = np.array(pz.target, dtype=np.int32)
        self._pull(rec, planes=True)
        if self.observation == "SPaRC":                                        # SG.py:153-164, 198-204
            self.observ = [[str(c) for c in row] for row in pz.observ]
            if any(len(r) != len(self.observ[0]) for r in self.observ):
                raise ValueError("Non-rectangular SPaRC grid")
            init_json = self._build_json_obs()
            self.observation_space = _spaces.Text(max_length=int(len(init_json) * 2),
                                                  charset="".join(sorted(set(init_json) | set("LV."))))
            return
        keys = list(self.obs_array.keys())
        self.observation_space = _spaces.Dict({                                # SG.py:190-196
            "base": _spaces.Dict({k: _spaces.Box(low=0, high=1, shape=(pz.X, pz.Y), dtype=np.int32) for k in keys}),
            "color": _spaces.Box(low=0, high=8, shape=(pz.X, pz.Y), dtype=np.int32),
            "additional_info": _spaces.Box(low=0, high=143632, shape=(pz.X, pz.Y), dtype=np.int64)})

    def _pull(self, rec=None, planes=False):
        """The device results of the last reset / step -> the reference's containers.  rec: the host record of
        sparc_*_info_host, or None = fetch field by field (the CPU twin of the tests).  planes: copy the observation
        planes from the device (after a reset); a step touches them in place instead (`_move`)."""
        v, st, pz = self._vec, self._store, self._pz
        if rec is None:
            flags, rmap, rfit = v.validate(region_map=True, region_fit=True)
            rec = {"rule_flags": flags.cpu().numpy().view(np.uint16), "region_map": rmap.cpu().numpy(), "region_fit": rfit.cpu().numpy(),
                   "agent_location": v.agent_xy.cpu().numpy(), "current_step": v.step_count.cpu().numpy(),
                   "legal_mask": v.legal_mask.cpu().numpy(), "reward_code": v.reward_code.cpu().numpy(),
                   "terminated": v.terminated.cpu().numpy(), "truncated": v.truncated.cpu().numpy()}
            planes = True
        if planes:
            base = v.obs["base"][0].cpu().numpy()
            for pi, key in enumerate(st.plane_keys):
                if key in self.obs_array:
                    self.obs_array[key][...] = base[pi, :pz.X, :pz.Y]
            self.color_array[...] = v.obs["color"][0, :pz.X, :pz.Y].cpu().numpy()
            self.additional_info[...] = v.obs["additional_info"][0, :pz.X, :pz.Y].cpu().numpy()
        self._agent_location[...] = rec["agent_location"][0]
        self.current_step = int(rec["current_step"][0])
        lm = int(rec["legal_mask"][0])
        self._legal = [a for a in range(4) if (lm >> a) & 1]
        f = int(rec["rule_flags"][0]) & 0xFFFF
        self._flags = {name: bool((f >> k) & 1) for k, name in enumerate(RULE_NAMES)}
        self._region_map = rec["region_map"][0, :pz.X, :pz.Y].copy()
        self._region_fit = rec["region_fit"][0].copy()
        return rec

    def _move(self, old, new):
        """In-place update of the planes after the agent moved (SG.py:1141-1188): a node it had visited before means a
        traceback undo (legality only allows path[-2], SG.py:1046-1048), which gives the left node back."""
        vis, ag = self.obs_array["visited"], self.obs_array["agent_location"]
        undo = vis[new] == 1
        ag[old] = 0
        if undo:
            vis[old] = 0
        vis[new] = 1
        ag[new] = 1
        return undo

    def _build_json_obs(self):                                                  # SG.py:988-992
        import json
        return json.dumps(self.observ, separators=(",", ":"))

    def _get_obs(self):                                                         # SG.py:956-983
        if self.observation == "SPaRC":
            return self._build_json_obs()
        return {"base": self.obs_array, "color": self.color_array, "additional_info": self.additional_info}

    def _get_info(self):                                                        # SG.py:994-1022
        items = {"solution_count": self.solution_count, "difficulty": self.difficulty,
                 "grid_x_size": self.x_size, "grid_y_size": self.y_size,
                 "legal_actions": list(self._legal), "current_step": self.current_step,
                 "agent_location": self._agent_location, "rule_status": _LazyInfo._PENDING,
                 "Rewards": {"normal_reward": self.normal_reward, "outcome_reward": self.outcome_reward}}
        if self.lazy_info:
            info = _LazyInfo(items, _Snap(self))
        else:                                   # the reference's plain dict, everything computed now
            info = items
            info["rule_status"] = _Details.rule_status(self)
        self._last_info = info
        return info

    @property
    def rule_status(self):                                                      # SG.py:64, 1010 (an attribute there)
        return self._last_info["rule_status"] if self._last_info is not None else {}

    @rule_status.setter
    def rule_status(self, value):
        pass

    # -- reference surface
    def reset(self, seed=None, options=None):                                   # SG.py:1057-1108
        super().reset(seed=seed)
        if options is not None:
            pid = options.get("puzzle_id", None)
            for idx, pz in enumerate(self.puzzles):
                if pz.id == pid:
                    self.current_puzzle_index = idx
                    break
        elif seed is not None:
            self.current_puzzle_index = int(self.np_random.integers(len(self.puzzles)))
        else:
            self.current_puzzle_index = (self.current_puzzle_index + 1) % len(self.puzzles)
        self._load_puzzle(self.current_puzzle_index)
        return self._get_obs(), self._get_info()

    def step(self, action):                                                     # SG.py:1111-1238
        import torch
        try:
            a = int(action)
            if a != action:
                a = -1
        except (TypeError, ValueError):
            a = -1                                                              # `action in legal` is False (SG.py:1137)
        v = self._vec
        old = (int(self._agent_location[0]), int(self._agent_location[1]))
        if self._fast:
            rec = v.step_info_host([a])
            self._pull(rec)
        else:                                    # field-by-field fetch: the planes come from the device, already moved
            old_len = int(self.obs_array["visited"].sum())
            v.step(torch.tensor([a], dtype=torch.int64))
            rec = self._pull(None)
        code = int(rec["reward_code"][0])
        terminated, truncated = bool(int(rec["terminated"][0])), bool(int(rec["truncated"][0]))
        self.normal_reward = abi.REWARD_VALUES[code]                            # ints 0/1/-1, floats +-0.01
        if code in (3, 4):                                                      # SG.py:1204-1215
            self.outcome_reward = 1 if code == 3 else -1
        elif not (terminated or truncated):
            self.outcome_reward = 0          # (a final step with reward 0: past the end after a win, outcome stays 1)
        new = (int(self._agent_location[0]), int(self._agent_location[1]))
        if new != old:
            undo = self._move(old, new) if self._fast else int(self.obs_array["visited"].sum()) < old_len
            if self.observation == "SPaRC":                                     # text overlays, SG.py:1149-1184
                self.observ[old[1]][old[0]] = ("." if self.obs_array["gaps"][old] == 1 else "+") if undo else "V"
                self.observ[new[1]][new[0]] = "L"
        return self._get_obs(), self.normal_reward, terminated, truncated, self._get_info()

    def render(self):                                                           # SG.py:1269-1299
        """The board as text, [y][x] like the reference's 'SPaRC' grid: 'L' agent, 'V' visited, 'E' target, 'G' gap,
        '.' dot, symbols as <kind initial><count / shape id>-<colour id>, '+' free node.  'human': printed, returns None
        (the reference opens a pygame window); 'llm': returned (the reference's docstring promises a text)."""
        if self.render_mode not in ("human", "llm"):
            return None
        p, col, info = self.obs_array, self.color_array, self.additional_info
        rows = []
        for y in range(self.y_size):
            cells = []
            for x in range(self.x_size):
                c = "+"
                for key in p:
                    if key not in FIXED_KEYS and key != "dot" and p[key][x, y] == 1:
                        c = key[0].upper() + (str(int(info[x, y])) if info[x, y] else "") + (f"-{int(col[x, y])}" if col[x, y] else "")
                if "dot" in p and p["dot"][x, y] == 1:
                    c = "."
                if p["gaps"][x, y] == 1 and x % 2 + y % 2 == 1:
                    c = "G"
                if p["target_location"][x, y] == 1:
                    c = "E"
                if p["visited"][x, y] == 1:
                    c = "V"
                if p["agent_location"][x, y] == 1:
                    c = "L"
                cells.append(c)
            rows.append(cells)
        wid = max(len(c) for r in rows for c in r)
        text = "\n".join(" ".join(c.ljust(wid) for c in r).rstrip() for r in rows)
        if self.render_mode == "human":
            print(text)
            return None
        return text

    def close(self):
        if getattr(self, "_vec", None) is not None:
            self._vec.close()
            self._vec = None


def register_envs():
    """Register both spellings: the reference registers 'SPaRC-Gym' (register_env.py:5-8), its README and
    BASELINE.json say 'SPaRc-Gym' (SURVEY Q14)."""
    if _gym is None:
        return False
    from gymnasium.envs.registration import register, registry
    for env_id in ("SPaRC-Gym", "SPaRc-Gym"):
        if env_id not in registry:
            register(id=env_id, entry_point="sparc_gym_b200.gym_env:SPaRC_Gym")
    return True


def make(env_id="SPaRC-Gym", **kwargs):
    """`gym.make` stand-in for machines without gymnasium."""
    if env_id not in ("SPaRC-Gym", "SPaRc-Gym"):
        raise ValueError(f"unknown env id {env_id!r}")
    return SPaRC_Gym(**kwargs)
