"""CPU tests of the product's device code (compiled for the host by tests/emu) against the
oracle: dynamics, rewards, truncation, traceback, validator, encoder -- everything but the
GPU execution itself (tests/test_gpu_parity.py does that on the B200)."""
import numpy as np
import pytest

from emu.emu_env import EmuVecEnv
from fuzzrows import fuzz_actions, fuzz_row
from parity import OracleLockstep, compare
from sparc_gym_b200.dataset import RULES_FULL, RULES_GAPS_DOTS, gen_pool, gen_puzzle


def run_lockstep(make_env, rows, N, T, traceback, max_steps, mode, validate, seed=0, follow=0.5, stride=1):
    env = make_env(rows, N, traceback=traceback, max_steps=max_steps, autoreset_mode=mode,
                   validate=validate, puzzle_stride=stride)
    orc = OracleLockstep(rows, N, env.store.plane_keys, (env.store.slab_x, env.store.slab_y), traceback,
                         max_steps, autoreset_mode="next_step" if mode == "next_step" else "disabled",
                         puzzle_stride=stride)
    exp = orc.reset()
    out = env.reset()
    compare((out[0], env.reward, env.terminated, env.truncated, out[1]), exp, "reset")
    rng = np.random.default_rng(seed)
    acts = np.stack([np.array(fuzz_actions(rows[i % len(rows)], seed * 1000 + i, T, p_follow=follow)) for i in range(N)], 1)
    n_done = 0
    for t in range(T):
        a = acts[t].astype(np.uint8 if t % 3 == 0 else np.int32 if t % 3 == 1 else np.int64)
        valid = orc.step(a)
        exp = orc.expected()
        res = env.step(a)
        if validate == "every_step":
            cf = valid
        elif validate == "terminal":
            cf = [v and (exp["terminated"][i] or exp["truncated"][i]) for i, v in enumerate(valid)]
        else:
            cf = [False] * N
        compare(res, exp, f"step {t}", check_flags=cf)
        n_done += int(sum(orc.done))
    return env, orc


@pytest.mark.parametrize("traceback", [False, True])
def test_gaps_dots_lockstep(traceback):
    rows = gen_pool(6, [(4, 4), (3, 3)], RULES_GAPS_DOTS)
    run_lockstep(EmuVecEnv, rows, 16, 70, traceback, 50, "disabled", "every_step")


@pytest.mark.parametrize("traceback", [False, True])
def test_fuzz_rows_every_step_validation(traceback):
    rows = [fuzz_row(s) for s in range(200, 236)]
    run_lockstep(EmuVecEnv, rows, 36, 45, traceback, 40, "disabled", "every_step", seed=3)


@pytest.mark.parametrize("defer", [False, True])
def test_full_rules_mixed_sizes_terminal_validation(defer):
    rows = gen_pool(8, [(3, 3), (4, 4), (5, 5), (6, 6)], RULES_FULL, seed0=40, extra_solutions=1)
    make = lambda *a, **k: EmuVecEnv(*a, defer_validation=defer, **k)
    env, _ = run_lockstep(make, rows, 24, 90, True, 80, "disabled", "terminal", seed=5, follow=0.97)
    assert env.stats()[7] > 0                     # some episodes ended with every rule satisfied


def test_deferred_validation_with_same_step_autoreset():
    """The queue snapshot must hold the TERMINAL state although the env is reset in the same step."""
    rows = gen_pool(6, [(2, 2), (3, 3), (4, 4)], RULES_FULL, seed0=21)
    a = EmuVecEnv(rows, 40, traceback=False, max_steps=30, autoreset_mode="same_step", validate="terminal")
    b = EmuVecEnv(rows, 40, traceback=False, max_steps=30, autoreset_mode="same_step", validate="terminal",
                  defer_validation=True, tile=32)
    a.reset(); b.reset()
    rng = np.random.default_rng(3)
    for t in range(80):
        act = rng.integers(0, 4, 40).astype(np.uint8)
        ra, rb = a.step(act), b.step(act)
        assert np.array_equal(ra[4]["rule_flags"], rb[4]["rule_flags"]), t
        assert np.array_equal(ra[0]["base"], rb[0]["base"]) and np.array_equal(ra[1], rb[1])
    assert np.array_equal(a.stats(), b.stats()) and a.stats()[0] > 50


def test_autoreset_next_step_cycles_puzzles():
    rows = gen_pool(5, [(2, 2), (3, 3)], RULES_FULL, seed0=7)
    env, orc = run_lockstep(EmuVecEnv, rows, 10, 120, False, 25, "next_step", "terminal", seed=9, follow=0.3)
    st = env.stats()
    assert st[0] > 10 and st[4] > 0          # episodes finished, steps counted


def test_encoder_is_tiling_invariant_and_handles_tiny_lattices():
    """Same envs, different CTA tilings / thread counts (incl. ragged last tiles and a 3x3
    lattice whose planes are shorter than one 32-bit word) -> identical arrays, all checked
    against the oracle."""
    for rows in ([gen_puzzle(1, 1, s, RULES_GAPS_DOTS) for s in range(3)],
                 gen_pool(5, [(1, 2), (2, 1), (3, 3)], RULES_FULL, seed0=70)):
        N, T = 37, 12
        ref = None
        for tile, nthr in ((4, 32), (8, 128), (32, 128), (128, 128), (64, 96)):
            env = EmuVecEnv(rows, N, traceback=True, max_steps=9, autoreset_mode="next_step", validate="terminal",
                            tile=tile, nthr=nthr)
            orc = OracleLockstep(rows, N, env.store.plane_keys, (env.store.slab_x, env.store.slab_y), True, 9,
                                 autoreset_mode="next_step")
            exp = orc.reset()
            out = env.reset()
            compare((out[0], env.reward, env.terminated, env.truncated, out[1]), exp, f"reset tile={tile}")
            rng = np.random.default_rng(5)
            for t in range(T):
                a = rng.integers(0, 4, N).astype(np.uint8)
                orc.step(a)
                compare(env.step(a), orc.expected(), f"tile={tile} step {t}")


def gap_start_rows():
    import yaml
    rows = []
    for seed in range(6):
        row = dict(gen_puzzle(3, 3, 40 + seed, RULES_GAPS_DOTS))
        tv = yaml.safe_load(row["text_visualization"])
        st = tv["puzzle"]["start"]
        tv["puzzle"]["cells"] = [c for c in tv["puzzle"]["cells"] if c["position"] != st] + \
            [{"position": dict(st), "properties": {"gap": True}}]
        row["text_visualization"] = yaml.safe_dump(tv)
        row["id"] = f"gapstart-{seed}"
        rows.append(row)
    return rows


def test_start_node_on_a_gap_cannot_be_traced_back_to():
    """SG.py:1024-1051 tests `gaps` for every candidate move, the traceback target included: an episode whose start node
    is itself marked as a gap can leave it but never undo back onto it (legality works on walk & ~visited plus a
    walkable-`path[-2]` test: this is the one case where that test says no)."""
    env, orc = run_lockstep(EmuVecEnv, gap_start_rows(), 24, 40, True, 30, "next_step", "every_step", seed=8, follow=0.2)
    assert env.stats()[0] > 0


def max_size_rows():
    """Pools at the ingest limit (7x7 cells = 15x15 lattice, SPARC_MAX_DIM) and non-square boards next to it:
    generated full-rule puzzles plus fuzz rows (shape ids colliding with triangle counts, Q10); fuzz rows over
    the documented ylop-instance limit are what ingest rejects, so they are left out here."""
    from sparc_gym_b200.ingest import ingest_rows
    from sparc_gym_b200.store import PuzzleStore
    gen = gen_pool(6, [(7, 7), (7, 5), (2, 7), (6, 7)], RULES_FULL, seed0=300)
    fuzz = []
    for s in range(500, 560):
        r = fuzz_row(s, max_cells=7, n_colors=8)
        try:
            PuzzleStore(ingest_rows([r]))
        except ValueError as ex:
            assert "ylop instances" in str(ex), ex
            continue
        fuzz.append(r)
    assert len(fuzz) > 40 and any(r["grid_size"]["width"] == 7 or r["grid_size"]["height"] == 7 for r in fuzz)
    return gen, fuzz


def test_maximum_lattice_and_non_square_boards():
    gen, fuzz = max_size_rows()
    env, _ = run_lockstep(EmuVecEnv, gen, 12, 140, True, 120, "disabled", "every_step", seed=9, follow=0.97)
    assert (env.store.slab_x, env.store.slab_y) == (15, 15) and env.stats()[1] > 0      # some listed solutions walked
    run_lockstep(EmuVecEnv, fuzz, len(fuzz), 60, True, 50, "disabled", "every_step", seed=4)


def test_dense_polyomino_boards_vs_oracle():
    """Stress puzzles for the exact-fit search (one region = the whole board, a dozen small pieces, 1e2..1e4
    search nodes): the listed border path is replayed to the end, then the terminal validation must agree with
    the oracle's restatement of the reference's search."""
    from sparc_gym_b200.dataset import gen_dense_poly_puzzle
    rows = [gen_dense_poly_puzzle(w, h, s) for (w, h, s) in [(3, 3, 0), (4, 4, 0), (4, 4, 3), (5, 5, 1), (5, 5, 2), (6, 6, 5)]]
    make = lambda *a, **k: EmuVecEnv(*a, defer_validation=True, **k)
    env, _ = run_lockstep(make, rows, 6, 26, False, 60, "disabled", "terminal", seed=1, follow=1.0)
    assert env.stats()[0] == 6 and env.stats()[1] == 6          # six episodes, six listed solutions walked (+1)
    assert 0 < env.stats()[7] < 6                               # some boards tile, some do not


@pytest.mark.parametrize("tile,sizes", [(16, [(3, 3), (4, 4), (5, 5), (6, 6)]), (32, [(1, 1), (2, 3)]), (48, [(4, 4)])])
def test_compact_observation_format_holds_the_same_values(tile, sizes):
    """obs_format='compact' (uint8 planes, uint8 colour, int32 additional_info): same values as the reference
    format, element for element, at every step -- incl. ragged last tiles and lattices whose env stride is odd."""
    rows = gen_pool(7, sizes, RULES_FULL, seed0=60)
    N = 2 * tile + 5
    kw = dict(traceback=True, max_steps=25, autoreset_mode="same_step", validate="terminal")
    a = EmuVecEnv(rows, N, tile=8 if tile == 16 else 4, **kw)
    b = EmuVecEnv(rows, N, tile=tile, obs_format="compact", **kw)
    assert b.obs["base"].dtype == np.uint8 and b.obs["color"].dtype == np.uint8 and b.obs["additional_info"].dtype == np.int32
    a.reset(); b.reset()
    rng = np.random.default_rng(4)
    for t in range(40):
        for k in a.obs:
            assert np.array_equal(a.obs[k], b.obs[k]), (t, k)
        act = rng.integers(0, 4, N).astype(np.uint8)
        a.step(act); b.step(act)
    assert np.array_equal(a.stats(), b.stats())


@pytest.mark.parametrize("fmt", ["reference", "compact"])
@pytest.mark.parametrize("mode,traceback,defer", [("same_step", False, True), ("same_step", True, False),
                                                  ("next_step", True, True), ("disabled", False, False)])
def test_delta_observation_equals_full_rewrite(fmt, mode, traceback, defer):
    """obs_update='delta' (the persistent observation is touched in place, as the reference mutates its aliased
    planes, SG.py:979 / 1146-1188; an env is rewritten in full only when it starts a new episode) leaves the same
    arrays behind as the full rewrite -- every step, every autoreset mode, both launch shapes (fused tile decision /
    k_dyn + k_encode_list), both formats; incl. a masked reset and a checkpoint restore in between."""
    rows = gen_pool(6, [(1, 1), (2, 3), (4, 4), (6, 6)], RULES_FULL, seed0=33)
    N = 53
    kw = dict(traceback=traceback, max_steps=18, autoreset_mode=mode, validate="terminal", obs_format=fmt,
              defer_validation=defer, tile=32 if fmt == "compact" else 8)
    a = EmuVecEnv(rows, N, **kw)
    b = EmuVecEnv(rows, N, obs_update="delta", **kw)
    a.reset(); b.reset()
    rng = np.random.default_rng(12)
    acts = np.stack([np.array(fuzz_actions(rows[i % len(rows)], 70 + i, 60, p_follow=0.7)) for i in range(N)], 1)
    for t in range(60):
        if t == 25:                                   # a masked reset onto other puzzles in the middle
            m = rng.integers(0, 2, N).astype(np.uint8)
            pz = rng.integers(0, len(rows), N).astype(np.int32)
            a.reset(pz, m); b.reset(pz, m)
        act = acts[t].astype(np.uint8)
        ra, rb = a.step(act), b.step(act)
        for k in a.obs:
            assert np.array_equal(a.obs[k], b.obs[k]), (t, k, np.argwhere(a.obs[k] != b.obs[k])[:4])
        assert np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2]) and np.array_equal(ra[3], rb[3])
        assert np.array_equal(ra[4]["rule_flags"], rb[4]["rule_flags"]) and np.array_equal(ra[4]["legal_mask"], rb[4]["legal_mask"])
    assert np.array_equal(a.stats(), b.stats()) and (mode == "disabled" or a.stats()[0] > N)


def test_more_than_six_ylops_run_on_wide_counters():
    """A puzzle with seven ylop instances is accepted (the reference has no limit), flagged for the 6-bit demand counters
    and validated: the tiling variant passes every rule on its listed solution, the variant whose areas do not balance
    fails the poly rule only (independent answers for such boards: tests/test_exact_fit.py)."""
    from fuzzrows import solution_actions
    from sparc_gym_b200.dataset import gen_many_ylops_puzzle
    from sparc_gym_b200.ingest import ingest_rows
    from sparc_gym_b200.store import PuzzleStore
    rows = [gen_many_ylops_puzzle(s, solvable=(s % 2 == 0)) for s in range(8)] + gen_pool(4, [(3, 3)], RULES_FULL, seed0=3)
    st = PuzzleStore(ingest_rows(rows))
    assert [int(f) & 4 for f in st.headers["flags"]] == [4] * 8 + [0] * 4
    for defer in (False, True):
        env = EmuVecEnv(rows, len(rows), max_steps=100, autoreset_mode="disabled", validate="terminal", defer_validation=defer)
        env.reset()
        sols = [solution_actions(r, 0) for r in rows]
        flags = np.zeros(len(rows), np.uint16)
        for t in range(max(map(len, sols))):
            a = np.array([s[t] if t < len(s) else 4 for s in sols], np.uint8)
            _, r, te, tr, info = env.step(a)
            flags = np.where(info["rule_flags"] != 0, info["rule_flags"], flags)
        assert [hex(int(f)) for f in flags[:8]] == ["0x81ff", "0x807f"] * 4, [hex(int(f)) for f in flags]
        assert all(int(f) == 0x81ff for f in flags[8:])
