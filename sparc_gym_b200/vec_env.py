"""Batched SPaRC environment on one B200: the host-side mirror of the reference's env class
(SPaRC_Gym/SPaRC_Gym.py:44-1265) for N lockstep instances.

The reference is single-env; this class keeps its vocabulary (`reset`, `step`, observation
'new' = {'base', 'color', 'additional_info'}, `traceback`, `max_steps`, rewards
{0, +-0.01, +-1}) and adds the batch dimension in front of every array.  All compute happens in
libsparc_b200.so (hand-written sm_100a CUDA) through the C ABI in include/sparc_b200.h; torch
only owns the tensors the kernels write into and the CUDA stream they run on.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi
from .ingest import ingest_rows
from .store import PuzzleStore


_ACT_CODES = {torch.uint8: abi.ACT_DTYPES["uint8"], torch.int32: abi.ACT_DTYPES["int32"], torch.int64: abi.ACT_DTYPES["int64"]}


class DeviceStore:
    """A PuzzleStore uploaded to one GPU (SparcStore handle)."""

    def __init__(self, store: PuzzleStore, device):
        self.host = store
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise abi.SparcError("sparc_gym_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = abi.load_library()
        desc, keep = abi.store_desc(store)
        handle = C.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()   # bare 'cuda' = the current device
        self.device = torch.device("cuda", index)
        abi.check(self.lib.sparc_store_create(C.byref(desc), index, C.byref(handle)))
        self.handle = handle

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sparc_store_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def _as_store(puzzles, **kw):
    if isinstance(puzzles, PuzzleStore):
        return puzzles
    if puzzles and isinstance(puzzles[0], dict):
        puzzles = ingest_rows(puzzles)
    return PuzzleStore(puzzles, **kw)


class SparcVecEnv:
    """N lockstep SPaRC envs on one GPU.

    puzzles        dataset rows / ingested puzzles / a PuzzleStore (shared pool; env i starts on
                   puzzle i mod pool)
    traceback, max_steps   as the reference ctor (SG.py:46)
    autoreset_mode 'next_step' (gymnasium >= 1.0 vector default), 'same_step' or 'disabled'
    puzzle_stride  on autoreset puzzle p -> (p + stride) mod pool (1 = the reference's cyclic
                   cursor, SG.py:1087; 0 = same puzzle again)
    validate       'terminal' | 'every_step' | 'never': when the rule validator
                   (SG.py:941-950) runs inside step(); its flags land in info['rule_flags']
    obs_format     'reference' (int32 planes, int32 colour, int64 additional_info: the reference's dtypes,
                   SG.py:272-280) or 'compact' (uint8 / uint8 / int32: same values, 3.5x fewer bytes)
    obs_update     'rewrite' (every step writes every observation in full) or 'delta': `self.obs` is the envs'
                   persistent observation, as the reference's aliased planes are (SG.py:979, 1146-1188) -- a step
                   touches the visited / agent_location cells that changed and rewrites an env in full only when it
                   starts a new episode.  Same tensors, same contents after every call; do not write into them.
    """

    def __init__(self, puzzles, num_envs, traceback=False, max_steps=2000, autoreset_mode="next_step",
                 puzzle_stride=1, validate="terminal", device="cuda:0", slab=None, planes="auto",
                 write_obs=True, obs_format="reference", obs_update="rewrite"):
        self.store = _as_store(puzzles, slab=slab, planes=planes) if not isinstance(puzzles, DeviceStore) else puzzles.host
        self.dstore = puzzles if isinstance(puzzles, DeviceStore) else DeviceStore(self.store, device)
        self.device = self.dstore.device
        self.lib = self.dstore.lib
        self.num_envs = int(num_envs)
        self.traceback, self.max_steps = bool(traceback), int(max_steps)
        self.autoreset_mode, self.validate_mode = autoreset_mode, validate
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            abi.check(self.lib.sparc_env_create(self.dstore.handle, self.num_envs, int(self.traceback),
                                                self.max_steps, abi.AUTORESET[autoreset_mode], int(puzzle_stride),
                                                abi.VALIDATE[validate], C.byref(handle)))
        self.handle = handle
        N, st, dev = self.num_envs, self.store, self.device
        self.obs_format = obs_format
        t_base, t_color, t_info = {"reference": (torch.int32, torch.int32, torch.int64),
                                   "compact": (torch.uint8, torch.uint8, torch.int32)}[obs_format]
        self.obs = {
            "base": torch.empty((N, st.n_planes, st.slab_x, st.slab_y), dtype=t_base, device=dev),
            "color": torch.empty((N, st.slab_x, st.slab_y), dtype=t_color, device=dev),
            "additional_info": torch.empty((N, st.slab_x, st.slab_y), dtype=t_info, device=dev),
        }
        self.reward = torch.zeros(N, dtype=torch.float32, device=dev)
        self.reward_code = torch.zeros(N, dtype=torch.int8, device=dev)
        self.terminated = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.legal_mask = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.rule_flags = torch.zeros(N, dtype=torch.int16, device=dev)     # bit pattern of uint16
        self.step_count = torch.zeros(N, dtype=torch.int32, device=dev)
        self.agent_xy = torch.zeros((N, 2), dtype=torch.uint8, device=dev)
        self._stats = torch.zeros(abi.N_STATS, dtype=torch.int64, device=dev)
        self._obs_out = abi.ObsOut(self.obs["base"].data_ptr(), self.obs["color"].data_ptr(),
                                   self.obs["additional_info"].data_ptr(), abi.OBS_FORMATS[obs_format],
                                   abi.OBS_UPDATES[obs_update])
        self.obs_update = obs_update
        self._no_obs = abi.ObsOut(None, None, None, 0, 0)
        self._step_out = abi.StepOut(self.reward.data_ptr(), self.reward_code.data_ptr(),
                                     self.terminated.data_ptr(), self.truncated.data_ptr(),
                                     self.legal_mask.data_ptr(), self.rule_flags.data_ptr(),
                                     self.step_count.data_ptr(), self.agent_xy.data_ptr())
        self.write_obs = write_obs
        self._obs_p, self._no_obs_p = C.byref(self._obs_out), C.byref(self._no_obs)
        self._step_p = C.byref(self._step_out)

    # -- helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _info(self):
        return {"legal_mask": self.legal_mask, "rule_flags": self.rule_flags, "reward_code": self.reward_code,
                "current_step": self.step_count, "agent_location": self.agent_xy}

    # -- reference surface, batched
    def reset(self, puzzle_idx=None, mask=None):
        """Clean reset (SG.py:1057-1108 without the plane leak).  `puzzle_idx`:
        int32[N] puzzle index per env (None = keep each env's current puzzle); `mask`: bool/uint8[N]
        which envs to reset (None = all)."""
        idx_ptr = mask_ptr = None
        if puzzle_idx is not None:
            if not (isinstance(puzzle_idx, torch.Tensor) and puzzle_idx.is_cuda):      # host data: range-check before uploading
                chk = np.asarray(puzzle_idx)
                if chk.size and (chk.min() < 0 or chk.max() >= self.store.n_puzzles):
                    raise ValueError("puzzle index out of range")
            self._idx = torch.as_tensor(puzzle_idx, dtype=torch.int32, device=self.device).contiguous()
            idx_ptr = self._idx.data_ptr()
        if mask is not None:
            self._mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            mask_ptr = self._mask.data_ptr()
        with torch.cuda.device(self.device):
            abi.check(self.lib.sparc_reset(self.handle, idx_ptr, mask_ptr, C.byref(self._step_out),
                                           C.byref(self._obs_out if self.write_obs else self._no_obs), self._stream()))
        return self.obs, self._info()

    def step(self, actions):
        """One lockstep step (SG.py:1111-1238).  `actions`: uint8/int32/int64 CUDA tensor [N]."""
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda):
            actions = torch.as_tensor(actions, device=self.device)
        code = _ACT_CODES.get(actions.dtype)
        if code is None:
            actions, code = actions.to(torch.int64), abi.ACT_DTYPES["int64"]
        if not actions.is_contiguous():
            actions = actions.contiguous()
        if actions.numel() != self.num_envs:
            raise ValueError("actions must have one entry per env")
        # (no torch.cuda.device() context: the library selects the store's device itself; on launch-bound batches the
        # context manager, the dtype-name string and two fresh byref objects were a third of the call)
        rc = self.lib.sparc_step(self.handle, actions.data_ptr(), code, self._step_p,
                                 self._obs_p if self.write_obs else self._no_obs_p, self._stream())
        if rc:
            abi.check(rc)
        return self.obs, self.reward, self.terminated, self.truncated, self._info()

    def step_host(self, actions_host, reward_host, terminated_host, truncated_host):
        """End-to-end step with HOST buffers (numpy / pinned torch CPU tensors): actions are copied
        to the device, results copied back, stream synchronised (sparc_step_host)."""
        ptr = lambda a: a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data
        # (no torch.cuda.device() context here: the library selects the store's device itself, and on a launch-bound
        # batch the context manager alone was a tenth of the call)
        rc = self.lib.sparc_step_host(self.handle, ptr(actions_host), ptr(reward_host), ptr(terminated_host), ptr(truncated_host),
                                      self._obs_p if self.write_obs else self._no_obs_p, self._stream())
        if rc:
            abi.check(rc)

    # -- host-record calls (the single-env drop-in): one launch sequence, ONE device->host copy, no torch ops
    def _info_record(self):
        if getattr(self, "_rec", None) is None:
            n, S = self.num_envs, self.store.slab_x * self.store.slab_y
            nbytes = self.lib.sparc_info_record_bytes(self.handle)
            self._rec_t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            self._ints_t = torch.zeros(n, dtype=torch.int32).pin_memory()
            buf, self._ints = self._rec_t.numpy(), self._ints_t.numpy()
            view = lambda off, cnt, dt, shape: buf[off:off + cnt * np.dtype(dt).itemsize].view(dt).reshape(shape)
            self._rec = {"current_step": view(0, n, np.int32, (n,)), "rule_flags": view(4 * n, n, np.uint16, (n,)),
                         "agent_location": view(6 * n, 2 * n, np.uint8, (n, 2)), "reward_code": view(8 * n, n, np.int8, (n,)),
                         "terminated": view(9 * n, n, np.uint8, (n,)), "truncated": view(10 * n, n, np.uint8, (n,)),
                         "legal_mask": view(11 * n, n, np.uint8, (n,)),
                         "region_fit": view(12 * n, abi.MAX_REGIONS * n, np.int8, (n, abi.MAX_REGIONS)),
                         "region_map": view((12 + abi.MAX_REGIONS) * n, S * n, np.int8, (n, self.store.slab_x, self.store.slab_y))}
        return self._rec

    def step_info_host(self, actions):
        """sparc_step_info_host: step every env with the int32 `actions` (host), validate the resulting states and return
        the host record (numpy views of one pinned buffer, overwritten by the next call): current_step, rule_flags,
        agent_location, reward_code, terminated, truncated, legal_mask, region_fit, region_map."""
        rec = self._info_record()
        self._ints[...] = actions
        abi.check(self.lib.sparc_step_info_host(self.handle, self._ints_t.data_ptr(), self._rec_t.data_ptr(),
                                                C.byref(self._obs_out if self.write_obs else self._no_obs), self._stream()))
        return rec

    def reset_info_host(self, puzzle_idx=None):
        """sparc_reset_info_host: clean reset of every env (on `puzzle_idx`, int32 host array, or its current puzzle) and
        the same host record as step_info_host."""
        rec = self._info_record()
        ptr = None
        if puzzle_idx is not None:
            idx = np.asarray(puzzle_idx)
            if idx.min() < 0 or idx.max() >= self.store.n_puzzles:
                raise ValueError("puzzle index out of range")
            self._ints[...] = idx
            ptr = self._ints_t.data_ptr()
        abi.check(self.lib.sparc_reset_info_host(self.handle, ptr, self._rec_t.data_ptr(),
                                                 C.byref(self._obs_out if self.write_obs else self._no_obs), self._stream()))
        return rec

    def encode(self):
        """Rewrite the observation tensors from the current state without stepping (sparc_encode)."""
        with torch.cuda.device(self.device):
            abi.check(self.lib.sparc_encode(self.handle, C.byref(self._obs_out), self._stream()))
        return self.obs

    def validate(self, mask=None, region_map=False, region_fit=False):
        """Rule validation of the current state (SG.py:941-950): uint16 flags (as int16 bit
        patterns), optionally the region map int8[N, X, Y] (-1 off-cell, SG.py:429-454) and the
        per-region poly/ylop verdicts int8[N, abi.MAX_REGIONS] (abi.FIT_*; SG.py:674-702)."""
        flags = torch.zeros(self.num_envs, dtype=torch.int16, device=self.device)
        rmap = torch.empty((self.num_envs, self.store.slab_x, self.store.slab_y), dtype=torch.int8,
                           device=self.device) if region_map else None
        rfit = torch.empty((self.num_envs, abi.MAX_REGIONS), dtype=torch.int8, device=self.device) if region_fit else None
        mask_ptr = None
        if mask is not None:
            self._vmask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            mask_ptr = self._vmask.data_ptr()
        with torch.cuda.device(self.device):
            abi.check(self.lib.sparc_validate(self.handle, mask_ptr, flags.data_ptr(),
                                              rmap.data_ptr() if region_map else None,
                                              rfit.data_ptr() if region_fit else None, self._stream()))
        out = (flags,) + ((rmap,) if region_map else ()) + ((rfit,) if region_fit else ())
        return out if len(out) > 1 else flags

    def stats(self, clear=False):
        """Episode statistics since creation / last clear, int64[8] on the device (abi.STAT_NAMES)."""
        with torch.cuda.device(self.device):
            abi.check(self.lib.sparc_stats(self.handle, self._stats.data_ptr(), int(clear), self._stream()))
        return self._stats

    def state_dict(self):
        n = self.lib.sparc_state_bytes(self.handle)
        blob = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            abi.check(self.lib.sparc_state_save(self.handle, blob.data_ptr(), self._stream()))
        return {"state": blob, "num_envs": self.num_envs}

    def load_state_dict(self, sd):
        blob = sd["state"].to(self.device).contiguous()
        if blob.numel() != self.lib.sparc_state_bytes(self.handle):
            raise ValueError("state blob does not match this env")
        with torch.cuda.device(self.device):
            abi.check(self.lib.sparc_state_load(self.handle, blob.data_ptr(), self._stream()))
        if self.obs_update == "delta" and self.write_obs:
            self.encode()                 # the persistent observation must show the restored state

    @property
    def kernel_launches(self):
        n = C.c_int64()
        abi.check(self.lib.sparc_kernel_launches(self.handle, C.byref(n)))
        return n.value

    def close(self):
        if getattr(self, "handle", None):
            torch.cuda.synchronize(self.device)
            self.lib.sparc_env_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class SparcVectorEnv(SparcVecEnv):
    """`gymnasium.vector.VectorEnv`-shaped face of SparcVecEnv (duck-typed: gymnasium is optional).

    `reset(seed=None, options=None) -> (obs, info)`, `step(actions) -> (obs, reward, terminated,
    truncated, info)` with batched CUDA tensors; `options={'puzzle_id': [...]}` or
    `{'puzzle_index': tensor}` picks puzzles like the reference's `reset(options=...)`
    (SG.py:1076-1082); a seed draws the per-env puzzle indices from `np.random.default_rng(seed)`
    (the reference draws one index for its one env, SG.py:1084-1085).  `info` carries
    `legal_mask` (bit a = action a legal; callers of the reference gate on
    `info["legal_actions"]`, human_play.py:55), `rule_flags`, `current_step`, `agent_location`.
    """

    metadata = {"autoreset_mode": None}

    def __init__(self, puzzles, num_envs, **kwargs):
        super().__init__(puzzles, num_envs, **kwargs)
        from .gym_env import _spaces
        st = self.store
        self.metadata = {"autoreset_mode": self.autoreset_mode}
        self.single_action_space = _spaces.Discrete(4)
        d_base, d_color, d_info = (np.int32, np.int32, np.int64) if self.obs_format == "reference" else (np.uint8, np.uint8, np.int32)
        self.single_observation_space = _spaces.Dict({
            "base": _spaces.Box(low=0, high=1, shape=(st.n_planes, st.slab_x, st.slab_y), dtype=d_base),
            "color": _spaces.Box(low=0, high=8, shape=(st.slab_x, st.slab_y), dtype=d_color),
            "additional_info": _spaces.Box(low=0, high=143632, shape=(st.slab_x, st.slab_y), dtype=d_info)})
        self.action_space, self.observation_space = self.single_action_space, self.single_observation_space
        self.plane_keys = list(st.plane_keys)

    def reset(self, *, seed=None, options=None, mask=None):
        """The reference's cursor rules (SG.py:1075-1087), per env: `options` given -> the env whose `puzzle_id` is in
        the pool goes there, an unknown id keeps the env's current puzzle, the seed is ignored; else a seed draws the
        indices; else every env advances its cursor by one (cyclically).  The cursors live on the host and follow
        explicit resets only (autoresets move the device-side puzzle by `puzzle_stride` on their own)."""
        n = self.store.n_puzzles
        if getattr(self, "_cursor", None) is None:
            self._cursor = np.arange(self.num_envs, dtype=np.int64) % n
        sel = np.ones(self.num_envs, bool) if mask is None else np.asarray(torch.as_tensor(mask).cpu()).astype(bool)
        cur = self._cursor.copy()
        if options is not None:
            if "puzzle_index" in options:
                cur = np.asarray(torch.as_tensor(options["puzzle_index"]).cpu()).astype(np.int64)
            elif "puzzle_id" in options:
                ids = options["puzzle_id"]
                ids = [ids] * self.num_envs if isinstance(ids, str) else list(ids)
                cur = np.array([self.store.id_to_index.get(i, int(c)) for i, c in zip(ids, cur)], dtype=np.int64)
        elif seed is not None:
            cur = np.random.default_rng(seed).integers(n, size=self.num_envs).astype(np.int64)
        else:
            cur = (cur + 1) % n
        self._cursor = np.where(sel, cur, self._cursor)
        return super().reset(puzzle_idx=torch.from_numpy(self._cursor.astype(np.int32)), mask=mask)
