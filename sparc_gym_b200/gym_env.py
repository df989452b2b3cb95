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
Deliberate differences: resets are clean, `info['rule_status'][rule]['detail']` is
reduced to what the device computes, and there are no pygame renderers.  observation='SPaRC'
(the JSON text grid, SG.py:1149-1184) is assembled on the host from the device's dynamics.
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
        flags, rmap = v.validate(region_map=True)
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

    def _regions_summary(self):
        """`_regions` of SG.py:949 from the device's region map + the puzzle's planes (SG.py:456-482)."""
        rmap = self._region_map
        n = int(rmap.max()) + 1
        out = {r: {"id": r, "area": int((rmap == r).sum()), "symbol_counts": {}, "colors": {}} for r in range(n)}
        for layer, arr in self.obs_array.items():
            if layer in FIXED_KEYS:
                continue
            for x, y in zip(*np.where(arr == 1)):
                rid = int(rmap[x, y])
                if rid == -1:
                    continue
                sc = out[rid]["symbol_counts"]
                sc[layer] = sc.get(layer, 0) + 1
                c = self.color_array[x, y]
                if c:
                    out[rid]["colors"][c] = out[rid]["colors"].get(c, 0) + 1
        return out

    def _build_json_obs(self):                                                  # SG.py:988-992
        import json
        return json.dumps(self.observ, separators=(",", ":"))

    def _get_obs(self):                                                         # SG.py:956-983
        if self.observation == "SPaRC":
            return self._build_json_obs()
        return {"base": self.obs_array, "color": self.color_array, "additional_info": self.additional_info}

    def _get_info(self):                                                        # SG.py:994-1022
        status = {k: {"passed": v, "detail": {}} for k, v in self._flags.items()}
        status["_terminated"] = {"passed": True, "detail": False}              # SG.py:1011, 937-938 (Q16)
        status["_truncated"] = {"passed": True, "detail": False}
        status["_regions"] = self._regions_summary()
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
