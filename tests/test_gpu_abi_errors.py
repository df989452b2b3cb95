"""GPU: error behaviour of the C ABI (include/sparc_b200.h): bad arguments come back as negative codes with a
message from sparc_last_error(), nothing is thrown across the boundary, handles survive a failed call, and the
gymnasium-facing classes turn codes into Python exceptions the way the reference raises its own (SG.py:87, 207)."""
import ctypes as C

import numpy as np
import pytest
import torch

from sparc_gym_b200 import abi
from sparc_gym_b200.dataset import RULES_FULL, gen_pool
from sparc_gym_b200.ingest import ingest_rows
from sparc_gym_b200.store import PuzzleStore

pytestmark = pytest.mark.gpu


def test_bad_arguments_return_codes_and_messages():
    lib = abi.load_library()
    st = PuzzleStore(ingest_rows(gen_pool(3, [(3, 3)], RULES_FULL, seed0=1)))
    desc, keep = abi.store_desc(st)
    hs, he = C.c_void_p(), C.c_void_p()
    assert lib.sparc_store_create(None, 0, C.byref(hs)) < 0 and b"null" in lib.sparc_last_error()
    bad = abi.StoreDesc.from_buffer_copy(desc)
    bad.slab_x = 99
    assert lib.sparc_store_create(C.byref(bad), 0, C.byref(hs)) < 0 and b"geometry" in lib.sparc_last_error()
    bad = abi.StoreDesc.from_buffer_copy(desc)
    bad.words_per_puzzle += 1
    assert lib.sparc_store_create(C.byref(bad), 0, C.byref(hs)) < 0
    bad = abi.StoreDesc.from_buffer_copy(desc)
    bad.n_puzzles = 2 ** 31 - 1                     # the encoder indexes the pool with 32-bit word offsets: refused up front
    assert lib.sparc_store_create(C.byref(bad), 0, C.byref(hs)) < 0 and b"too large" in lib.sparc_last_error()
    assert lib.sparc_store_create(C.byref(desc), 0, C.byref(hs)) == 0
    assert lib.sparc_env_create(hs, 0, 0, 10, 0, 1, 1, C.byref(he)) < 0            # n_envs <= 0
    assert lib.sparc_env_create(hs, 8, 0, 10, 7, 1, 1, C.byref(he)) < 0            # autoreset mode
    assert lib.sparc_env_create(hs, 8, 0, 10, 0, 1, 9, C.byref(he)) < 0 and b"mode" in lib.sparc_last_error()
    assert lib.sparc_env_create(hs, 8, 0, 10, 0, 1, 1, C.byref(he)) == 0
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    acts = torch.zeros(8, dtype=torch.uint8, device="cuda")
    assert lib.sparc_step(he, None, 0, None, None, stream) < 0                     # no actions
    assert lib.sparc_step(he, acts.data_ptr(), 5, None, None, stream) < 0          # action dtype
    obs = abi.ObsOut(None, None, None, 7, 0)
    assert lib.sparc_step(he, acts.data_ptr(), 0, None, C.byref(obs), stream) < 0 and b"format" in lib.sparc_last_error()
    assert lib.sparc_encode(he, None, stream) < 0
    assert lib.sparc_stats(he, None, 0, stream) < 0
    assert lib.sparc_state_save(he, None, stream) < 0 and lib.sparc_state_load(he, None, stream) < 0
    assert lib.sparc_reset(None, None, None, None, None, stream) < 0
    # the handle is still good after all those refusals
    assert lib.sparc_reset(he, None, None, None, None, stream) == 0
    assert lib.sparc_step(he, acts.data_ptr(), 0, None, None, stream) == 0
    out = torch.zeros(abi.N_STATS, dtype=torch.int64, device="cuda")
    assert lib.sparc_stats(he, out.data_ptr(), 0, stream) == 0
    torch.cuda.synchronize()
    assert int(out[4]) == 8                                                        # eight env-steps were counted
    assert lib.sparc_env_num_envs(he) == 8 and lib.sparc_env_num_envs(None) == 0
    assert lib.sparc_env_destroy(he) == 0 and lib.sparc_env_destroy(None) == 0
    assert lib.sparc_store_destroy(hs) == 0 and lib.sparc_store_destroy(None) == 0


def test_python_surface_raises_like_the_reference():
    import sparc_gym_b200
    rows = gen_pool(2, [(3, 3)], RULES_FULL, seed0=2)
    with pytest.raises(ValueError, match="Invalid observation type"):             # SG.py:207
        sparc_gym_b200.SPaRC_Gym(puzzles=rows, observation="pixels")
    with pytest.raises(ValueError, match="No valid dataframe"):                   # SG.py:86-87
        sparc_gym_b200.SPaRC_Gym(puzzles=[])
    env = sparc_gym_b200.SPaRC_Gym(puzzles=rows, max_steps=3)
    env.reset()
    for a in (0, 1, 2):
        _, _, te, tr, _ = env.step(a)
    assert te or tr                                                               # max_steps = 3
    _, r, te, tr, info = env.step(0)                  # like the reference, a finished episode is stepped on
    assert tr and r in (1, -1) and type(r) is int and info["current_step"] == 4
    assert sparc_gym_b200.SPaRC_Gym(puzzles=rows, render_mode="human").render_mode == "human"      # accepted, SG.py:68-74
    from sparc_gym_b200.vec_env import SparcVecEnv
    with pytest.raises(KeyError):
        SparcVecEnv(rows, 4, obs_format="float16")
    with pytest.raises(KeyError):
        SparcVecEnv(rows, 4, autoreset_mode="sometimes")


def test_nvtx_ranges_do_not_change_results():
    """SPARC_NVTX=1 wraps every ABI call in an NVTX range (tracing); a fresh process, because the
    switch is read once."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); import torch\n"
            "from sparc_gym_b200.dataset import RULES_FULL, gen_pool\n"
            "from sparc_gym_b200.vec_env import SparcVecEnv\n"
            "rows = gen_pool(4, [(3, 3)], RULES_FULL, seed0=2)\n"
            "out = []\n"
            "for n in (64, 9000):\n"
            "    env = SparcVecEnv(rows, n, max_steps=9, autoreset_mode='same_step'); env.reset()\n"
            "    g = torch.Generator(device='cuda').manual_seed(1)\n"
            "    for t in range(20): env.step(torch.randint(0, 4, (n,), dtype=torch.uint8, device='cuda', generator=g))\n"
            "    out.append(env.stats().cpu().tolist())\n"
            "print(out)\n") % root
    runs = [subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, SPARC_NVTX=v)) for v in ("0", "1")]
    assert all(r.returncode == 0 for r in runs), runs[1].stderr[-400:]
    assert runs[0].stdout == runs[1].stdout and "[[" in runs[0].stdout
