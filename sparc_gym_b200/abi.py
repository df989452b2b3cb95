"""ctypes binding of the C ABI in include/sparc_b200.h (libsparc_b200.so).

No torch types cross the boundary: callers pass raw device pointers (`tensor.data_ptr()`),
sizes and a `cudaStream_t` handle.  There is no CPU fallback -- if the CUDA library is missing
or no device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
# SPARC_B200_LIB selects another build of the same CUDA library (kernel tuning experiments)
LIB_PATH = os.environ.get("SPARC_B200_LIB") or os.path.join(CSRC, "libsparc_b200.so")
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

MAX_PLANES = 10
N_STATS = 8
STAT_NAMES = ("episodes", "solved", "failed", "truncated", "steps", "return_cents", "length", "rules_pass")
RULE_VALID = 1 << 15
MAX_REGIONS = 64                              # SPARC_MAX_REGIONS
FIT_NONE, FIT_AREA_MISMATCH, FIT_NO_TILING, FIT_OK = 0, 1, 2, 3   # SPARC_FIT_*: per-region poly/ylop verdicts
AUTORESET = {"disabled": 0, "next_step": 1, "same_step": 2, "continue": 3}   # continue: the reference's no-guard stepping (Q4)
VALIDATE = {"never": 0, "terminal": 1, "every_step": 2}
ACT_DTYPES = {"uint8": 0, "int32": 1, "int64": 2}
REWARD_VALUES = (0, 0.01, -0.01, 1, -1)      # reward code -> the reference's Python value (SG.py:1204-1223)

EXPORTS = ("sparc_abi_version", "sparc_last_error", "sparc_store_create", "sparc_store_destroy",
           "sparc_env_create", "sparc_env_destroy", "sparc_reset", "sparc_step", "sparc_encode", "sparc_validate",
           "sparc_stats", "sparc_step_host", "sparc_info_record_bytes", "sparc_step_info_host", "sparc_reset_info_host",
           "sparc_state_bytes", "sparc_state_save", "sparc_state_load", "sparc_env_num_envs", "sparc_kernel_launches")


class SparcError(RuntimeError):
    pass


class StoreDesc(C.Structure):
    _fields_ = [("n_puzzles", C.c_int32), ("slab_x", C.c_int32), ("slab_y", C.c_int32),
                ("n_planes", C.c_int32), ("plane_ids", C.c_int32 * MAX_PLANES),
                ("words_per_puzzle", C.c_int32), ("n_shapes_total", C.c_int32),
                ("n_trie_nodes", C.c_int32),
                ("headers", C.c_void_p), ("node_boards", C.c_void_p), ("cell_boards", C.c_void_p),
                ("static_bits", C.c_void_p), ("color4", C.c_void_p), ("info_nz", C.c_void_p), ("info", C.c_void_p),
                ("shapes", C.c_void_p), ("trie", C.c_void_p)]


OBS_UPDATES = {"rewrite": 0, "delta": 1}         # SPARC_OBS_REWRITE (full rewrite every call) | SPARC_OBS_DELTA (in place)
OBS_FORMATS = {"reference": 0, "compact": 1}     # SPARC_OBS_*: int32/int32/int64 (the reference's dtypes) | uint8/uint8/int32


class ObsOut(C.Structure):
    _fields_ = [("base_dev", C.c_void_p), ("color_dev", C.c_void_p), ("info_dev", C.c_void_p),
                ("format", C.c_int32), ("update", C.c_int32)]


class StepOut(C.Structure):
    _fields_ = [("reward_dev", C.c_void_p), ("reward_code_dev", C.c_void_p),
                ("terminated_dev", C.c_void_p), ("truncated_dev", C.c_void_p),
                ("legal_mask_dev", C.c_void_p), ("rule_flags_dev", C.c_void_p),
                ("step_count_dev", C.c_void_p), ("agent_xy_dev", C.c_void_p)]


def build_library(force=False, verbose=False):
    """Compile csrc/sparc_b200.cu for sm_100a into csrc/libsparc_b200.so (in-tree)."""
    src = os.path.join(CSRC, "sparc_b200.cu")
    deps = [src, os.path.join(CSRC, "sparc_core.cuh"), os.path.join(CSRC, "sparc_val_warp.cuh"), os.path.join(REPO_ROOT, "include", "sparc_b200.h")]
    if not force and os.path.exists(LIB_PATH) and \
            os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "-o", LIB_PATH, src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SparcError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SparcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(the B200 environment has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.sparc_abi_version.restype = i32
    lib.sparc_last_error.restype = C.c_char_p
    lib.sparc_store_create.argtypes = [C.POINTER(StoreDesc), i32, C.POINTER(vp)]
    lib.sparc_store_destroy.argtypes = [vp]
    lib.sparc_env_create.argtypes = [vp, i64, i32, i32, i32, i32, i32, C.POINTER(vp)]
    lib.sparc_env_destroy.argtypes = [vp]
    lib.sparc_reset.argtypes = [vp, vp, vp, C.POINTER(StepOut), C.POINTER(ObsOut), vp]
    lib.sparc_step.argtypes = [vp, vp, i32, C.POINTER(StepOut), C.POINTER(ObsOut), vp]
    lib.sparc_encode.argtypes = [vp, C.POINTER(ObsOut), vp]
    lib.sparc_validate.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.sparc_stats.argtypes = [vp, vp, i32, vp]
    lib.sparc_step_host.argtypes = [vp, vp, vp, vp, vp, C.POINTER(ObsOut), vp]
    lib.sparc_info_record_bytes.argtypes = [vp]
    lib.sparc_info_record_bytes.restype = i64
    lib.sparc_step_info_host.argtypes = [vp, vp, vp, C.POINTER(ObsOut), vp]
    lib.sparc_reset_info_host.argtypes = [vp, vp, vp, C.POINTER(ObsOut), vp]
    lib.sparc_state_bytes.argtypes = [vp]
    lib.sparc_state_bytes.restype = i64
    lib.sparc_state_save.argtypes = [vp, vp, vp]
    lib.sparc_state_load.argtypes = [vp, vp, vp]
    lib.sparc_env_num_envs.argtypes = [vp]
    lib.sparc_env_num_envs.restype = i64
    lib.sparc_kernel_launches.argtypes = [vp, C.POINTER(i64)]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("sparc_abi_version",):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SparcError(f"libsparc_b200: error {rc}: {load_library().sparc_last_error().decode()}")


def store_desc(store):
    """PuzzleStore (host numpy) -> (StoreDesc, keepalive list)."""
    from .store import PLANE_IDS
    d = StoreDesc()
    d.n_puzzles, d.slab_x, d.slab_y, d.n_planes = store.n_puzzles, store.slab_x, store.slab_y, store.n_planes
    for i in range(MAX_PLANES):
        d.plane_ids[i] = PLANE_IDS[store.plane_keys[i]] if i < store.n_planes else -1
    d.words_per_puzzle, d.n_shapes_total, d.n_trie_nodes = store.words_per_puzzle, store.n_shapes_total, len(store.trie)
    keep = []
    for field, arr in (("headers", store.headers), ("node_boards", store.node_boards),
                       ("cell_boards", store.cell_boards), ("static_bits", store.static_bits),
                       ("color4", store.color4), ("info_nz", store.info_nz), ("info", store.info), ("shapes", store.shapes),
                       ("trie", store.trie)):
        a = np.ascontiguousarray(arr)
        keep.append(a)
        setattr(d, field, a.ctypes.data)
    return d, keep
