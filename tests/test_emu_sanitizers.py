"""The device code (sparc_core.cuh compiled for the host) under AddressSanitizer + UBSan, with the
emulated shared memory sized exactly like the kernels' dynamic shared memory: out-of-bounds row /
board accesses, over-wide shifts and signed overflows in the dynamics, validator and encoder show
up here (compute-sanitizer is not available on the GPU pool)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SCRIPT = r"""
import sys
sys.path[:0] = [%r, %r, %r]
import numpy as np
from emu.emu_env import EmuVecEnv
from fuzzrows import fuzz_row
from sparc_gym_b200.dataset import RULES_FULL, gen_pool, gen_puzzle, RULES_GAPS_DOTS
pools = [[fuzz_row(s) for s in range(400, 430)], gen_pool(6, [(1, 1), (2, 3), (6, 6), (7, 7)], RULES_FULL, seed0=3),
         [gen_puzzle(1, 1, 0, RULES_GAPS_DOTS)]]
for rows in pools:
    for tile, nthr, N in ((4, 32, 37), (32, 128, 70), (128, 128, 130)):
        for tb in (False, True):
            env = EmuVecEnv(rows, N, traceback=tb, max_steps=17, autoreset_mode="same_step", validate="every_step",
                            tile=tile, nthr=nthr)
            env.reset()
            rng = np.random.default_rng(tile)
            for t in range(40):
                env.step(rng.integers(0, 5, N).astype(np.int64))
            env.validate(region_map=True, region_fit=True)
    # compact observation format (uint8 / uint8 / int32), deferred validation queue
    for tile, N in ((16, 37), (64, 130)):
        env = EmuVecEnv(rows, N, traceback=True, max_steps=17, autoreset_mode="same_step", validate="terminal",
                        tile=tile, obs_format="compact", defer_validation=True)
        env.reset()
        rng = np.random.default_rng(tile)
        for t in range(40):
            env.step(rng.integers(0, 4, N).astype(np.uint8))
    # delta observation update: in-place plane writes + the per-env rewrite of k_encode_list (rows of tile = 1)
    for fmt, defer, tile in (("reference", True, 8), ("compact", True, 32), ("reference", False, 4)):
        env = EmuVecEnv(rows, 45, traceback=True, max_steps=11, autoreset_mode="same_step", validate="terminal",
                        tile=tile, obs_format=fmt, obs_update="delta", defer_validation=defer)
        env.reset()
        rng = np.random.default_rng(9)
        for t in range(40):
            env.step(rng.integers(0, 4, 45).astype(np.uint8))
print("sanitizer run ok")
""" % (ROOT, HERE, os.path.join(ROOT, "oracle"))


def test_device_code_under_asan_ubsan(tmp_path):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    ubsan = subprocess.run(["gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    if not (os.path.isabs(asan) and os.path.exists(asan)):
        pytest.skip("libasan not available")
    lib = str(tmp_path / "libsparc_emu_asan.so")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-fsanitize=address,undefined",
                    "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-Wno-unknown-pragmas", "-o", lib,
                    os.path.join(HERE, "emu", "emu.cpp")], check=True)
    env = dict(os.environ, SPARC_EMU_LIB=lib, LD_PRELOAD=f"{asan}:{ubsan}" if os.path.exists(ubsan) else asan,
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
    res = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "sanitizer run ok" in res.stdout, res.stderr[-3000:]
