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
Deliberate differences: resets are clean and there are no pygame renderers.
observation='SPaRC' (the JSON text grid, SG.py:1149-1184) is assembled on the host from the
device's dynamics.
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


class SPaRC_Gym(_Env):
    metadata = {"render_modes": ["human", "llm"], "render_fps": 30}            # SG.py:45

    def __init__(self, df_name="lkaesberg/SPaRC", df_split="all", df_set="test", render_mode=None,
                 observation="new", traceback=False, max_steps=2000, *, puzzles=None, device="cuda:0"):
        if observation not in ("new", "SPaRC"):
            raise ValueError("Invalid observation type. Choose 'new' or 'SPaRC'.")     # SG.py:207
        if render_mode is not None:
            raise ValueError("the B200 environment has no renderers (render_mode must be None)")
        from .vec_env import SparcVecEnv
        self.render_mode, self.observation = render_mode, observation
        self.traceback, self.max_steps = traceback, max_steps
        rows = puzzles if puzzles is not None else _load_rows(df_name, df_split, df_set)
        if rows is None or len(rows) == 0:
            raise ValueError("No valid dataframe provided")                    # SG.py:86-87
        self.puzzles = ingest_rows(rows, observation)
        self._store = PuzzleStore(self.puzzles)
        self._vec = SparcVecEnv(self._store, 1, traceback=traceback, max_steps=max_steps,
                                autoreset_mode="disabled", validate="never", device=device)   # _pull() validates on demand
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
        self._vec.reset(puzzle_idx=torch.tensor([index], dtype=torch.int32))
        # fresh containers per load, then updated in place while the episode runs (SG.py:149-151, 979)
        self.obs_array = {k: np.zeros((pz.X, pz.Y), np.int32) for k in pz.keys}
        self.color_array = np.zeros((pz.X, pz.Y), np.int32)
        self.additional_info = np.zeros((pz.X, pz.Y), np.int64)
        self._agent_location = np.zeros(2, np.int32)
        self._target_location = np.array(pz.target, dtype=np.int32)
        self._pull()
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

    def _pull(self):
        """Copy the device results of the last reset/step into the reference's containers."""
        v, st, pz = self._vec, self._store, self._pz
        flags, rmap, rfit = v.validate(region_map=True, region_fit=True)
        base = v.obs["base"][0].cpu().numpy()
        for pi, key in enumerate(st.plane_keys):
            if key in self.obs_array:
                self.obs_array[key][...] = base[pi, :pz.X, :pz.Y]
        self.color_array[...] = v.obs["color"][0, :pz.X, :pz.Y].cpu().numpy()
        self.additional_info[...] = v.obs["additional_info"][0, :pz.X, :pz.Y].cpu().numpy()
        self._agent_location[...] = v.agent_xy[0].cpu().numpy()
        self.current_step = int(v.step_count[0])
        self._legal = [a for a in range(4) if (int(v.legal_mask[0]) >> a) & 1]
        f = int(flags[0]) & 0xFFFF
        self._flags = {name: bool((f >> k) & 1) for k, name in enumerate(RULE_NAMES)}
        self._region_map = rmap[0, :pz.X, :pz.Y].cpu().numpy()
        self._region_fit = rfit[0].cpu().numpy()

    def _region_symbols(self):
        """Per region id: area, symbols[layer] = [(x, y), ...] and colour counts (SG.py:456-482), from the
        device's region map and the puzzle's planes."""
        rmap = self._region_map
        n = int(rmap.max()) + 1
        area = [int((rmap == r).sum()) for r in range(n)]
        symbols = [dict() for _ in range(n)]
        colors = [dict() for _ in range(n)]
        for layer, arr in self.obs_array.items():
            if layer in FIXED_KEYS:
                continue
            for x, y in zip(*np.where(arr == 1)):
                rid = int(rmap[x, y])
                if rid == -1:
                    continue
                symbols[rid].setdefault(layer, []).append((int(x), int(y)))
                c = self.color_array[x, y]
                if c:
                    colors[rid][c] = colors[rid].get(c, 0) + 1
        return area, symbols, colors

    def _rule_details(self, area, symbols):
        """The `detail` payloads of SG.py:488-710.  Every fact comes from the device (flags, region map,
        per-region poly/ylop verdicts, visited plane); this only arranges them in the reference's containers."""
        planes, color, info = self.obs_array, self.color_array, self.additional_info
        d = {}
        d["reached_target"] = {"agent_loc": self._agent_location.tolist(),              # SG.py:493-496
                               "target_loc": self._target_location.tolist()}
        d["path_not_crossing"] = {"duplicates": {}}          # path nodes are distinct (clean resets, legal moves)
        sx, sy = self.start_location                          # the only path node that can sit on a gap (SG.py:508-519)
        d["no_gap_violations"] = {"violations": [(int(sx), int(sy))] if planes["gaps"][sx, sy] == 1 else []}
        if "dot" in planes:                                                             # SG.py:521-532
            dots = planes["dot"] == 1
            d["all_dots_collected"] = {"total": int(dots.sum()),
                                       "collected": int((dots & (planes["visited"] == 1)).sum())}
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
        if "triangle" in planes:                                                        # SG.py:617-646
            tri, vis, mism = planes["triangle"], planes["visited"], []
            for x in range(1, self.x_size - 1):
                for y in range(1, self.y_size - 1):
                    if tri[x, y] == 1 and int(info[x, y]) > 0:
                        touches = int(vis[x + 1, y] + vis[x - 1, y] + vis[x, y - 1] + vis[x, y + 1])
                        if touches != int(info[x, y]):
                            mism.append({"x": x, "y": y, "required": int(info[x, y]), "touches": touches})
            d["triangles_edge_count"] = {"mismatches": mism}
        else:
            d["triangles_edge_count"] = {"mismatches": []}
        by_region = {}                                                                  # SG.py:648-736
        if isinstance(self.polyshapes, dict):
            for x in range(self.x_size):
                for y in range(self.y_size):
                    v = info[x, y]
                    if v != 0 and f"{v}" in self.polyshapes:
                        rid = int(self._region_map[x, y])
                        if rid != -1:
                            kind = "poly" if planes["poly"][x, y] == 1 else "ylop"
                            by_region.setdefault(rid, []).append((kind, int(np.array(self.polyshapes[f"{v}"]).sum())))
        if not by_region and not self._has_instances():
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

    def _build_json_obs(self):                                                  # SG.py:988-992
        import json
        return json.dumps(self.observ, separators=(",", ":"))

    def _get_obs(self):                                                         # SG.py:956-983
        if self.observation == "SPaRC":
            return self._build_json_obs()
        return {"base": self.obs_array, "color": self.color_array, "additional_info": self.additional_info}

    def _get_info(self):                                                        # SG.py:994-1022
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
        self.rule_status = status
        return {"solution_count": self.solution_count, "difficulty": self.difficulty,
                "grid_x_size": self.x_size, "grid_y_size": self.y_size,
                "legal_actions": list(self._legal), "current_step": self.current_step,
                "agent_location": self._agent_location, "rule_status": status,
                "Rewards": {"normal_reward": self.normal_reward, "outcome_reward": self.outcome_reward}}

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
        if int(v.terminated[0]) or int(v.truncated[0]):
            raise RuntimeError("episode is over: call reset() (the B200 env does not step finished episodes, SURVEY Q4)")
        old = (int(self._agent_location[0]), int(self._agent_location[1]))
        old_len = int(self.obs_array["visited"].sum())
        v.step(torch.tensor([a], dtype=torch.int64))
        code = int(v.reward_code[0])
        terminated, truncated = bool(int(v.terminated[0])), bool(int(v.truncated[0]))
        self.normal_reward = abi.REWARD_VALUES[code]                            # ints 0/1/-1, floats +-0.01
        self.outcome_reward = {3: 1, 4: -1}.get(code, 0)                        # SG.py:1204-1215
        self._pull()
        new = (int(self._agent_location[0]), int(self._agent_location[1]))
        if self.observation == "SPaRC" and new != old:                          # text overlays, SG.py:1149-1184
            if int(self.obs_array["visited"].sum()) < old_len:                  # traceback undo restores the node
                self.observ[old[1]][old[0]] = "." if self.obs_array["gaps"][old] == 1 else "+"
            else:
                self.observ[old[1]][old[0]] = "V"
            self.observ[new[1]][new[0]] = "L"
        return self._get_obs(), self.normal_reward, terminated, truncated, self._get_info()

    def render(self):                                                           # SG.py:1269-1299
        return None

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
