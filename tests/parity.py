"""Lockstep parity harness: N oracle envs (oracle/sparc_oracle.py) driven with the batched
semantics of SparcVecEnv (autoreset modes, cyclic puzzle stride), and exact comparison of every
array a batched env returns.  Works on numpy arrays, so the same checks run against the CPU
emulation of the device code (tests/emu) and against the CUDA library on a GPU."""
from __future__ import annotations

import copy

import numpy as np

import sparc_oracle as so
from sparc_gym_b200 import abi

RULE_BITS = {name: 1 << k for k, name in enumerate(so.RULE_NAMES)}
REWARD_CODE = {(int, 0): 0, (float, 0.01): 1, (float, -0.01): 2, (int, 1): 3, (int, -1): 4}


def to_np(x):
    if isinstance(x, np.ndarray):
        return x
    return x.detach().cpu().numpy()


class OracleLockstep:
    def __init__(self, rows, num_envs, plane_keys, slab, traceback, max_steps, autoreset_mode="disabled",
                 puzzle_stride=1):
        self.rows, self.N = rows, num_envs
        self.plane_keys, self.slab = list(plane_keys), tuple(slab)
        self.mode, self.stride = autoreset_mode, puzzle_stride
        proto = so.OracleEnv(rows, observation="new", traceback=traceback, max_steps=max_steps)
        self.envs = []
        for i in range(num_envs):
            e = copy.copy(proto)
            e.puzzles = copy.deepcopy(proto.puzzles)
            self.envs.append(e)
        self.puzzle = [i % len(rows) for i in range(num_envs)]
        self.done = [False] * num_envs
        self.term_trunc = [(False, False)] * num_envs
        self.last = [None] * num_envs

    def _load(self, i, idx):
        e = self.envs[i]
        e.current_puzzle_index = idx
        e.current_step = 0
        e._load(idx)
        self.puzzle[i] = idx
        self.done[i] = False
        self.term_trunc[i] = (False, False)
        self.last[i] = (e._obs(), 0, False, False, e._info())

    def reset(self, puzzle_idx=None, mask=None):
        for i in range(self.N):
            if mask is None or mask[i]:
                self._load(i, int(puzzle_idx[i]) % len(self.rows) if puzzle_idx is not None else self.puzzle[i])
            else:   # untouched env: same state, reward fields zero
                o, _, _, _, info = self.last[i]
                self.last[i] = (o, 0, self.term_trunc[i][0], self.term_trunc[i][1], info)
        return self.expected(flags_valid=False)

    def step(self, actions):
        valid = [False] * self.N
        for i in range(self.N):
            e = self.envs[i]
            if self.done[i]:
                if self.mode == "next_step":
                    self._load(i, (self.puzzle[i] + self.stride) % len(self.rows))
                else:   # frozen
                    o, _, _, _, info = self.last[i]
                    self.last[i] = (o, 0, self.term_trunc[i][0], self.term_trunc[i][1], info)
                continue
            o, r, te, tr, info = e.step(int(actions[i]))
            self.last[i] = (o, r, te, tr, info)
            valid[i] = True
            if te or tr:
                self.done[i] = True
                self.term_trunc[i] = (te, tr)
        return valid

    def expected(self, flags_valid=None):
        """Arrays the batched env must hold now (call right after reset/step, BEFORE same_step resets)."""
        P, (SX, SY) = len(self.plane_keys), self.slab
        base = np.zeros((self.N, P, SX, SY), np.int32)
        color = np.zeros((self.N, SX, SY), np.int32)
        info_pl = np.zeros((self.N, SX, SY), np.int64)
        out = {k: [] for k in ("reward_code", "terminated", "truncated", "legal", "step", "agent", "flags")}
        for i in range(self.N):
            o, r, te, tr, info = self.last[i]
            X, Y = o["color"].shape
            for pi, key in enumerate(self.plane_keys):
                if key in o["base"]:
                    assert o["base"][key].dtype == np.int32
                    base[i, pi, :X, :Y] = o["base"][key]
            assert o["color"].dtype == np.int32 and o["additional_info"].dtype == np.int64
            color[i, :X, :Y] = o["color"]
            info_pl[i, :X, :Y] = o["additional_info"]
            out["reward_code"].append(REWARD_CODE[(type(r), r)])
            out["terminated"].append(int(te))
            out["truncated"].append(int(tr))
            out["legal"].append(sum(1 << a for a in info["legal_actions"]))
            out["step"].append(info["current_step"])
            out["agent"].append([int(v) for v in info["agent_location"]])
            out["flags"].append(sum(RULE_BITS[k] for k in so.RULE_NAMES if info["rule_status"][k]["passed"]))
        return {"base": base, "color": color, "additional_info": info_pl,
                **{k: np.array(v) for k, v in out.items()}}


def compare(env_out, exp, what, check_flags=None, check_obs=True, only=None):
    """env_out = (obs, reward, terminated, truncated, info) of the batched env (numpy or torch)."""
    obs, reward, term, trunc, info = env_out
    sel = slice(None) if only is None else np.asarray(only)
    def eq(name, got, want):
        got, want = to_np(got)[sel], want[sel]
        if got.shape != want.shape or not np.array_equal(got, want):
            bad = np.argwhere(np.asarray(got != want).reshape(len(want), -1).any(1)).ravel()[:5] if got.shape == want.shape else "shape"
            raise AssertionError(f"{what}: {name} mismatch at envs {bad}\n got={got[bad[0]] if len(bad) else got}\nwant={want[bad[0]] if len(bad) else want}")
    if check_obs:
        for k in ("base", "color", "additional_info"):
            assert to_np(obs[k]).dtype == exp[k].dtype, (k, to_np(obs[k]).dtype)
            eq(k, obs[k], exp[k])
    eq("reward_code", info["reward_code"], exp["reward_code"].astype(np.int8))
    eq("reward", reward, np.array(abi.REWARD_VALUES, np.float32)[exp["reward_code"]])
    eq("terminated", term, exp["terminated"].astype(np.uint8))
    eq("truncated", trunc, exp["truncated"].astype(np.uint8))
    eq("legal_mask", info["legal_mask"], exp["legal"].astype(np.uint8))
    eq("current_step", info["current_step"], exp["step"].astype(np.int32))
    eq("agent_location", info["agent_location"], exp["agent"].astype(np.uint8))
    if check_flags is not None:
        got = to_np(info["rule_flags"]).astype(np.uint16)
        for i in np.arange(len(got))[sel]:
            if check_flags[i]:
                assert got[i] & abi.RULE_VALID, f"{what}: env {i} flags not marked valid"
                assert (got[i] & 0x1FF) == exp["flags"][i], \
                    f"{what}: env {i} rule flags {got[i] & 0x1FF:09b} != oracle {exp['flags'][i]:09b}"
            else:
                assert got[i] == 0, f"{what}: env {i} has stale rule flags"


# ---- BASELINE configs[0]: the reference's own CPU-runnable case, recorded step by step (tests/golden/config1.json.gz) ----
def obs_crc(obs):
    """CRC32 over the observation as the reference holds it: base planes in the puzzle's key order (int32[X, Y]
    each), then color (int32), then additional_info (int64).  (Same function as tests/golden/make_golden.py.)"""
    import zlib
    c = 0
    for k in obs["base"]:
        c = zlib.crc32(np.ascontiguousarray(to_np(obs["base"][k]), np.int32).tobytes(), c)
    c = zlib.crc32(np.ascontiguousarray(to_np(obs["color"]), np.int32).tobytes(), c)
    return zlib.crc32(np.ascontiguousarray(to_np(obs["additional_info"]), np.int64).tobytes(), c)


def replay_config1(env, g, n_steps=None):
    """Drive a single env with the reference's call surface (reset(options=...) / step(a)) through the recorded
    config-1 rollout and compare every step with what the unmodified reference returned: reward value AND type,
    terminated, truncated, legal actions, agent location, current_step, all_rules_satisfied and the CRC of the
    whole observation; resets onto the same puzzle whenever an episode ends, like the recording."""
    cfg, row, gold = g["config"], g["row"], g["golden"]
    acts = np.random.default_rng(cfg["action_seed"]).integers(0, 4, size=cfg["T"], dtype=np.uint8).tolist()
    obs, info = env.reset(options={"puzzle_id": row["id"]})
    assert list(obs["base"].keys()) == gold["keys"] and obs_crc(obs) == gold["reset_crc"]
    episodes = 0
    for t, a in enumerate(acts[:n_steps]):
        obs, r, te, tr, info = env.step(a)
        got = ([type(r).__name__, float(r)], int(bool(te)), int(bool(tr)), sum(1 << int(x) for x in info["legal_actions"]),
               [int(v) for v in info["agent_location"]], int(info["current_step"]),
               int(bool(info["rule_status"]["all_rules_satisfied"]["passed"])), obs_crc(obs))
        want = (gold["reward"][t], gold["term"][t], gold["trunc"][t], gold["legal"][t], gold["agent"][t], gold["step"][t],
                gold["rules"][t], gold["crc"][t])
        assert got == want, (t, got, want)
        if te or tr:
            episodes += 1
            obs, info = env.reset(options={"puzzle_id": row["id"]})
            assert obs_crc(obs) == gold["reset_crc"], t
    return episodes
