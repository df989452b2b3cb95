"""TEST-ONLY numpy twin of sparc_gym_b200.vec_env.SparcVecEnv that drives tests/emu/libsparc_emu.so
(the product's device code compiled for the CPU).  Used by the CPU test-suite to check the
kernels' logic against the oracle without a GPU; never imported by the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sparc_gym_b200 import abi
from sparc_gym_b200.ingest import ingest_rows
from sparc_gym_b200.store import PuzzleStore

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsparc_emu.so")


def build_emu():
    if os.environ.get("SPARC_EMU_LIB"):          # e.g. the ASan/UBSan build of tests/test_emu_sanitizers.py
        return os.environ["SPARC_EMU_LIB"]
    src = os.path.join(HERE, "emu.cpp")
    deps = [src, os.path.join(abi.CSRC, "sparc_core.cuh"), os.path.join(abi.REPO_ROOT, "include", "sparc_b200.h")]
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", LIB, src], check=True)
    return LIB


class EmuVecEnv:
    def __init__(self, puzzles, num_envs, traceback=False, max_steps=2000, autoreset_mode="next_step",
                 puzzle_stride=1, validate="terminal", slab=None, planes="auto", tile=8, nthr=128, defer_validation=False,
                 obs_format="reference", obs_update="rewrite"):
        if not isinstance(puzzles, PuzzleStore):
            if puzzles and isinstance(puzzles[0], dict):
                puzzles = ingest_rows(puzzles)
            puzzles = PuzzleStore(puzzles, slab=slab, planes=planes)
        self.store = st = puzzles
        self.lib = C.CDLL(build_emu())
        self.desc, self._keep = abi.store_desc(st)
        N = self.num_envs = int(num_envs)
        self.cfg = (int(bool(traceback)), int(max_steps), abi.AUTORESET[autoreset_mode], int(puzzle_stride),
                    abi.VALIDATE[validate])
        self.tile, self.nthr, self.defer = tile, nthr, int(defer_validation)
        self.core = np.zeros((N, 4), np.uint32)
        self.core[:, 0] = np.arange(N) % st.n_puzzles
        self.vis = np.zeros((8, N), np.uint32)
        self.ep_ret = np.zeros(N, np.int32)
        self.moves = np.zeros((12, N), np.uint32) if traceback else None
        self.stats_acc = np.zeros(abi.N_STATS, np.uint64)
        t_base, t_color, t_info = {"reference": (np.int32, np.int32, np.int64), "compact": (np.uint8, np.uint8, np.int32)}[obs_format]
        self.obs = {"base": np.full((N, st.n_planes, st.slab_x, st.slab_y), 249, t_base),
                    "color": np.full((N, st.slab_x, st.slab_y), 249, t_color),
                    "additional_info": np.full((N, st.slab_x, st.slab_y), -7, t_info)}
        self.reward = np.zeros(N, np.float32)
        self.reward_code = np.zeros(N, np.int8)
        self.terminated = np.zeros(N, np.uint8)
        self.truncated = np.zeros(N, np.uint8)
        self.legal_mask = np.zeros(N, np.uint8)
        self.rule_flags = np.zeros(N, np.uint16)
        self.step_count = np.zeros(N, np.int32)
        self.agent_xy = np.zeros((N, 2), np.uint8)
        p = lambda a: a.ctypes.data
        self._obs_out = abi.ObsOut(p(self.obs["base"]), p(self.obs["color"]), p(self.obs["additional_info"]),
                                   abi.OBS_FORMATS[obs_format], abi.OBS_UPDATES[obs_update])
        self._step_out = abi.StepOut(p(self.reward), p(self.reward_code), p(self.terminated), p(self.truncated),
                                     p(self.legal_mask), p(self.rule_flags), p(self.step_count), p(self.agent_xy))

    def _launch(self, mode, actions=None, dtype=0, idx=None, mask=None):
        p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
        rc = self.lib.emu_launch(C.byref(self.desc), p(self.core), p(self.vis), p(self.ep_ret), p(self.moves),
                                 p(self.stats_acc), C.c_longlong(self.num_envs), *self.cfg, mode, self.tile, self.nthr,
                                 p(actions), dtype, p(idx), p(mask), C.byref(self._step_out), C.byref(self._obs_out), self.defer)
        assert rc == 0

    def _info(self):
        return {"legal_mask": self.legal_mask, "rule_flags": self.rule_flags, "reward_code": self.reward_code,
                "current_step": self.step_count, "agent_location": self.agent_xy}

    def reset(self, puzzle_idx=None, mask=None):
        idx = None if puzzle_idx is None else np.ascontiguousarray(puzzle_idx, np.int32)
        m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
        self._launch(1, idx=idx, mask=m)
        return self.obs, self._info()

    def step(self, actions):
        a = np.ascontiguousarray(actions)
        name = str(a.dtype)
        if name not in abi.ACT_DTYPES:
            a, name = a.astype(np.int64), "int64"
        self._launch(0, actions=a, dtype=abi.ACT_DTYPES[name])
        return self.obs, self.reward, self.terminated, self.truncated, self._info()

    def validate(self, mask=None, region_map=False, region_fit=False):
        flags = np.zeros(self.num_envs, np.uint16)
        rmap = np.zeros((self.num_envs, self.store.slab_x, self.store.slab_y), np.int8) if region_map else None
        rfit = np.zeros((self.num_envs, abi.MAX_REGIONS), np.int8) if region_fit else None
        p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
        self.lib.emu_validate(C.byref(self.desc), p(self.core), p(self.vis), p(self.ep_ret),
                              C.c_longlong(self.num_envs), p(flags), p(rmap), p(rfit))
        out = (flags,) + ((rmap,) if region_map else ()) + ((rfit,) if region_fit else ())
        return out if len(out) > 1 else flags

    def stats(self, clear=False):
        out = self.stats_acc.astype(np.int64)
        if clear:
            self.stats_acc[:] = 0
        return out
