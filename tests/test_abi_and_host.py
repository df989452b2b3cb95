"""CPU tests of the boundary and host logic: the C-ABI library loads and exports every symbol
include/sparc_b200.h declares (no compute without a GPU), struct layouts agree between the
header, numpy and ctypes, the store packer is consistent with the oracle's planes, and the
multi-GPU sharding + statistics reduction works over gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import sparc_oracle as so
from fuzzrows import fuzz_row
from sparc_gym_b200 import abi
from sparc_gym_b200.dataset import RULES_FULL, gen_pool, gen_puzzle
from sparc_gym_b200.ingest import ingest_rows
from sparc_gym_b200.runner import shard_puzzles, shard_range
from sparc_gym_b200.store import HEADER_DTYPE, SHAPE_DTYPE, PuzzleStore

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "sparc_b200.h")).read()


def test_library_exports_every_declared_symbol():
    abi.build_library()
    declared = set(re.findall(r"\b(sparc_[a-z_]+)\s*\(", HEADER))
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    lib = abi.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sparc_abi_version() == int(re.search(r"#define SPARC_ABI_VERSION (\d+)", HEADER).group(1))
    nm = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert all(re.search(rf"\bT {n}\b", nm) for n in declared)


def test_no_compute_without_a_device_and_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    from sparc_gym_b200.vec_env import SparcVecEnv
    rows = [gen_puzzle(2, 2, 0, RULES_FULL)]
    with pytest.raises(abi.SparcError):
        SparcVecEnv(rows, 4, device="cpu")
    lib = abi.load_library()
    desc, keep = abi.store_desc(PuzzleStore(ingest_rows(rows)))
    h = C.c_void_p()
    assert lib.sparc_store_create(C.byref(desc), 0, C.byref(h)) == -2          # SPARC_E_CUDA
    assert b"cuda" in lib.sparc_last_error().lower()


def test_product_package_never_touches_the_checkers():
    """The oracle, the host build of the device code (tests/emu), the staged reference (baseline/_ref) and
    /root/reference are test / bench infrastructure: nothing under sparc_gym_b200/ or in the CUDA sources may import,
    load or name them -- a product path that routes through a checker voids every parity claim."""
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sparc_gym_b200")
    bad = re.compile(r"oracle|libsparc_emu|tests[/.]emu|emu_env|baseline|_ref\b|/root/reference|refharness")
    hits = []
    for dirpath, _, names in os.walk(root):
        for name in names:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                path = os.path.join(dirpath, name)
                for ln, line in enumerate(open(path, encoding="utf-8"), 1):
                    line = line.split("#" if name.endswith(".py") else "//")[0]          # (comments may explain the set-up)
                    if bad.search(line):
                        hits.append(f"{os.path.relpath(path, root)}:{ln}: {line.strip()[:100]}")
    assert not hits, "\n".join(hits)


def test_struct_layouts_match_the_header():
    assert HEADER_DTYPE.itemsize == 32 and SHAPE_DTYPE.itemsize == 32
    # StoreDesc: 17 int32 (padded to 72 bytes) + 9 pointers, exactly as a C compiler lays it out
    assert abi.StoreDesc.headers.offset == 72 and C.sizeof(abi.StoreDesc) == 72 + 9 * 8
    nocomment = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)

    def members(struct):
        body = re.search(r"typedef struct \{([^}]*)\} %s;" % struct, nocomment, re.S).group(1)
        out = []
        for decl in body.split(";"):
            for part in decl.split(","):
                m = re.search(r"(\w+)\s*(?:\[\w+\])?\s*$", part.strip())
                if m:
                    out.append(m.group(1))
        return out

    for struct, cls in (("SparcStoreDesc", abi.StoreDesc), ("SparcObsOut", abi.ObsOut), ("SparcStepOut", abi.StepOut)):
        assert members(struct) == [f[0] for f in cls._fields_], struct
    enums = dict(re.findall(r"(SPARC_(?:NB|CB)_[A-Z0-9]+) = (\d+)", HEADER))
    from sparc_gym_b200 import store
    assert int(enums["SPARC_NB_TRIBAD"]) == store.NB_TRIBAD and int(enums["SPARC_CB_COLOR1"]) == store.CB_COLOR1


def test_product_ingest_equals_oracle_ingest():
    rows = [fuzz_row(s) for s in range(300, 340)] + gen_pool(6, [(3, 3), (6, 6)], RULES_FULL, seed0=5)
    a, b = ingest_rows(rows), so.process_rows(rows)
    for p, q in zip(a, b):
        assert p.keys == list(q["obs_array"].keys())
        for k in p.keys:
            assert p.planes[k].dtype == np.int32 and np.array_equal(p.planes[k], q["obs_array"][k])
        assert np.array_equal(p.color, q["color_array"]) and np.array_equal(p.info, q["additional_info"])
        assert p.info.dtype == np.int64 and (p.start, p.target) == (q["start_location"], q["target_location"])


def test_store_static_bits_decode_back_to_the_planes():
    rows = gen_pool(5, [(2, 3), (4, 4), (6, 5)], RULES_FULL, seed0=9)
    st = PuzzleStore(ingest_rows(rows))
    S = st.slab_x * st.slab_y
    for i, pz in enumerate(st.puzzles):
        bits = np.unpackbits(st.static_bits[i].view(np.uint8), bitorder="little")[:st.n_planes * S]
        planes = bits.reshape(st.n_planes, st.slab_x, st.slab_y)
        for pi, key in enumerate(st.plane_keys):
            want = np.zeros((st.slab_x, st.slab_y), np.uint8)
            if key == "target_location":
                want[pz.target] = 1
            elif key in pz.planes and key not in ("visited", "agent_location"):
                want[:pz.X, :pz.Y] = pz.planes[key]
            assert np.array_equal(planes[pi], want), (pz.id, key)
        nib = np.unpackbits(st.color4[i].view(np.uint8), bitorder="little")[:4 * S].reshape(S, 4)
        assert np.array_equal((nib * [1, 2, 4, 8]).sum(1).reshape(st.slab_x, st.slab_y), st.color[i])
        nz = np.unpackbits(st.info_nz[i].view(np.uint8), bitorder="little")[:S].reshape(st.slab_x, st.slab_y)
        assert np.array_equal(nz.astype(bool), st.info[i] != 0)


def test_store_rejects_what_the_reference_cannot_run():
    row = gen_puzzle(2, 2, 3, RULES_FULL)
    row["solution_count"] = 5                      # more than listed: IndexError in SG.py:1205-1206
    with pytest.raises(ValueError):
        PuzzleStore(ingest_rows([row]))
    big = gen_puzzle(2, 2, 3, RULES_FULL)
    big["grid_size"] = {"width": 8, "height": 2}
    with pytest.raises(ValueError):
        PuzzleStore(ingest_rows([big]))


def test_sharding_covers_every_env_once():
    for n, world in ((1 << 20, 8), (1000, 3), (7, 8)):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        cat = np.concatenate([shard_puzzles(n, r, world, 13) for r in range(world)])
        assert np.array_equal(cat, np.arange(n) % 13)


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from sparc_gym_b200.runner import reduce_stats
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    local = torch.arange(8, dtype=torch.int64) * (rank + 1) - (5 if rank else 0)    # negative return cents too
    out[rank] = reduce_stats(local)
    dist.destroy_process_group()


def test_stats_reduction_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 1000
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    want = {n: k * 1 + (k * 2 - 5) for k, n in enumerate(abi.STAT_NAMES)}
    assert dict(out[0]) == want and dict(out[1]) == want


def _gloo_async_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from sparc_gym_b200.runner import AsyncStats
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    counters = torch.zeros(8, dtype=torch.int64)
    a = AsyncStats(lambda stream: counters, every=4)
    seen = []
    for t in range(1, 21):
        counters += torch.arange(8, dtype=torch.int64) * (rank + 1)          # this rank's "step"
        a.tick()
        if t % 4 == 0:
            # an all-reduce on the DEFAULT group in between: the statistics travel on a communicator of their own
            x = torch.ones(1)
            dist.all_reduce(x)
    got = a.latest(wait=True)
    seen.append(got)
    out[rank] = (a.reductions, got[0], got[1])
    dist.destroy_process_group()


def test_async_stats_reduction_over_gloo_world_size_2():
    """AsyncStats (ShardedRunner.stats_async): periodic reduction over a process group of its own, every 4th tick; the
    newest completed result is the sum over ranks of the counters at tick 20."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 30600 + os.getpid() % 1000
    mp.spawn(_gloo_async_worker, args=(2, port, out), nprocs=2, join=True)
    want = {n: k * 20 * 3 for k, n in enumerate(abi.STAT_NAMES)}            # ranks add k and 2k per tick
    for r in (0, 1):
        n_red, tick, stats = out[r]
        assert n_red == 5 and tick == 20 and stats == want


def test_rows_survive_a_parquet_round_trip(tmp_path):
    """Real SPaRC rows arrive as pandas/pyarrow objects (numpy arrays of dicts, numpy scalars): the
    loader + ingest must give the same planes as the plain-Python rows."""
    import pandas as pd
    from sparc_gym_b200.dataset import load_rows
    rows = gen_pool(5, [(3, 3), (5, 5)], RULES_FULL, seed0=31) + [fuzz_row(s) for s in range(600, 605)]
    for r in rows:
        r.pop("puzzle_array", None)            # ragged object column; only observation='SPaRC' reads it
    path = tmp_path / "sparc.parquet"
    pd.DataFrame(rows).to_parquet(path)
    back = load_rows(path)
    a, b = ingest_rows(rows), ingest_rows(back)
    sa, sb = PuzzleStore(a), PuzzleStore(b)
    for name in ("headers", "node_boards", "cell_boards", "static_bits", "color4", "info_nz", "info", "shapes", "trie"):
        assert np.array_equal(getattr(sa, name), getattr(sb, name)), name
    assert [p.keys for p in a] == [p.keys for p in b]


def test_import_compat_shim():
    import SPaRC_Gym as shim
    import sparc_gym_b200
    assert shim.SPaRC_Gym is sparc_gym_b200.SPaRC_Gym
