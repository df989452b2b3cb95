"""GPU parity tests: the CUDA library (through the C ABI, via SparcVecEnv) against the oracle on
the same seeded puzzles and action streams -- bit-exact on every returned array, every step."""
import numpy as np
import pytest
import torch

from fuzzrows import fuzz_row
from parity import to_np
from test_emu_parity import max_size_rows, run_lockstep
from sparc_gym_b200.dataset import RULES_FULL, RULES_GAPS_DOTS, gen_pool

pytestmark = pytest.mark.gpu


def make_env(rows, N, **kw):
    from sparc_gym_b200.vec_env import SparcVecEnv
    return SparcVecEnv(rows, N, device="cuda:0", **kw)


@pytest.mark.parametrize("traceback", [False, True])
def test_gaps_dots_lockstep(traceback):
    rows = gen_pool(6, [(4, 4), (3, 3)], RULES_GAPS_DOTS)
    run_lockstep(make_env, rows, 16, 70, traceback, 50, "disabled", "every_step")


@pytest.mark.parametrize("traceback", [False, True])
def test_fuzz_rows_every_step_validation(traceback):
    rows = [fuzz_row(s) for s in range(200, 236)]
    run_lockstep(make_env, rows, 36, 45, traceback, 40, "disabled", "every_step", seed=3)


def test_start_node_on_a_gap_cannot_be_traced_back_to():
    from test_emu_parity import gap_start_rows
    env, _ = run_lockstep(make_env, gap_start_rows(), 24, 40, True, 30, "next_step", "every_step", seed=8, follow=0.2)
    assert int(env.stats()[0]) > 0


def test_full_rules_mixed_sizes_terminal_validation():
    rows = gen_pool(8, [(3, 3), (4, 4), (5, 5), (6, 6)], RULES_FULL, seed0=40, extra_solutions=1)
    run_lockstep(make_env, rows, 24, 90, True, 80, "disabled", "terminal", seed=5, follow=0.97)


def test_autoreset_next_step_cycles_puzzles():
    rows = gen_pool(5, [(2, 2), (3, 3)], RULES_FULL, seed0=7)
    env, orc = run_lockstep(make_env, rows, 10, 120, False, 25, "next_step", "terminal", seed=9, follow=0.3)
    st = to_np(env.stats())
    assert st[0] > 10 and st[4] > 0


def test_maximum_lattice_and_non_square_boards():
    """15x15 lattices (SPARC_MAX_DIM) and non-square boards, generated and fuzzed, under the fused and the split step."""
    gen, fuzz = max_size_rows()
    env, _ = run_lockstep(make_env, gen, 12, 140, True, 120, "disabled", "every_step", seed=9, follow=0.97)
    assert (env.store.slab_x, env.store.slab_y) == (15, 15) and to_np(env.stats())[1] > 0
    run_lockstep(make_env, fuzz, len(fuzz), 60, True, 50, "disabled", "every_step", seed=4)
    # the split step (>= 8,192 envs) on the 15x15 slab against the CPU build of the same device code
    from emu.emu_env import EmuVecEnv
    N, kw = 8192 + 5, dict(traceback=True, max_steps=30, autoreset_mode="same_step", validate="terminal")
    g, c = make_env(gen + fuzz, N, **kw), EmuVecEnv(gen + fuzz, N, tile=32, **kw)
    g.reset(); c.reset()
    rng = np.random.default_rng(1)
    for t in range(40):
        a = rng.integers(0, 4, N).astype(np.uint8)
        go, gr, gt, gu, gi = g.step(torch.from_numpy(a).cuda())
        co, cr, ct, cu, ci = c.step(a)
        for k in go:
            assert np.array_equal(to_np(go[k]), co[k]), (t, k)
        assert np.array_equal(to_np(gi["reward_code"]), ci["reward_code"]) and np.array_equal(to_np(gt), ct), t
        assert np.array_equal(to_np(gu), cu) and np.array_equal(to_np(gi["rule_flags"]).astype(np.uint16), ci["rule_flags"]), t
    assert np.array_equal(to_np(g.stats()), c.stats()) and c.stats()[0] > 500


@pytest.fixture(params=[("0", "8"), ("0", "32"), ("0", "128"), ("0", "20"), ("3", "8"), ("4", "32"), ("1", "4"), ("1", "36"), ("1", "128")],
                ids=lambda p: f"split{p[0]}-tile{p[1]}")
def kernel_variant(request, monkeypatch):
    """Every launch strategy: fused k_env at several CTA tilings (incl. tiles that are not multiples of 16, which the
    compact format must round up, and the largest tile, which needs > 48 KB of dynamic shared memory), and the
    k_dyn -> k_val || k_encode split (side stream + events), single- and multi-chunk."""
    monkeypatch.setenv("SPARC_SPLIT", request.param[0])
    monkeypatch.setenv("SPARC_TILE", request.param[1])
    monkeypatch.setenv("SPARC_ENC_TILE", request.param[1])
    return request.param


def test_lockstep_under_every_kernel_variant(kernel_variant):
    rows = gen_pool(6, [(2, 2), (4, 4), (6, 6)], RULES_FULL, seed0=90, extra_solutions=1)
    run_lockstep(make_env, rows, 70, 40, True, 30, "next_step", "terminal", seed=2, follow=0.8)


def test_gpu_matches_cpu_emulation_at_scale(kernel_variant):
    """65,536-env class sizes cannot be replayed by the Python oracle; compare the CUDA run with the
    same device code compiled for the CPU (itself pinned to the oracle by test_emu_parity.py)."""
    from emu.emu_env import EmuVecEnv
    rows = gen_pool(32, [(3, 3), (4, 4), (5, 5), (6, 6)], RULES_FULL, seed0=300)
    N, T = 4096 + 3, 60          # ragged tail tile on purpose
    kw = dict(traceback=True, max_steps=40, autoreset_mode="same_step", validate="terminal", puzzle_stride=3)
    g = make_env(rows, N, **kw)
    c = EmuVecEnv(rows, N, tile=32, **kw)
    g.reset(); c.reset()
    rng = np.random.default_rng(0)
    for t in range(T):
        a = rng.integers(0, 4, N).astype(np.uint8)
        go, gr, gt, gu, gi = g.step(torch.from_numpy(a).cuda())
        co, cr, ct, cu, ci = c.step(a)
        for k in go:
            assert np.array_equal(to_np(go[k]), co[k]), (t, k)
        assert np.array_equal(to_np(gi["reward_code"]), ci["reward_code"]), t
        assert np.array_equal(to_np(gt), ct) and np.array_equal(to_np(gu), cu), t
        assert np.array_equal(to_np(gi["legal_mask"]), ci["legal_mask"]), t
        assert np.array_equal(to_np(gi["rule_flags"]).astype(np.uint16), ci["rule_flags"]), t
    assert np.array_equal(to_np(g.stats()), c.stats())
    assert c.stats()[0] > 1000


def test_checkpoint_restore_and_standalone_encode():
    """state_dict -> steps -> load_state_dict -> encode() gives back the checkpointed observation, and
    the restored env continues exactly like a twin that never left (sparc_state_save/load, sparc_encode)."""
    rows = gen_pool(6, [(3, 3), (5, 5)], RULES_FULL, seed0=77)
    a, b = make_env(rows, 300, traceback=True, max_steps=25), make_env(rows, 300, traceback=True, max_steps=25)
    a.reset(); b.reset()
    rng = np.random.default_rng(8)
    acts = [torch.from_numpy(rng.integers(0, 4, 300).astype(np.uint8)).cuda() for _ in range(30)]
    for t in range(10):
        a.step(acts[t]); b.step(acts[t])
    ck = a.state_dict()
    want = {k: v.clone() for k, v in a.obs.items()}
    for t in range(10, 20):
        a.step(acts[t])
    a.load_state_dict(ck)
    got = a.encode()
    for k in want:
        assert torch.equal(got[k], want[k]), k
    for t in range(10, 30):
        ra, rb = a.step(acts[t]), b.step(acts[t])
        assert torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2]) and torch.equal(ra[3], rb[3])
        assert all(torch.equal(ra[0][k], rb[0][k]) for k in ra[0])


@pytest.mark.parametrize("obs_update", ["rewrite", "delta"])
def test_step_host_matches_device_step(obs_update):
    """sparc_step_host (host action / result buffers: packed launch + completion flag on the fused kernel; result copies
    beside the encoder on the split path; results written into the pinned arrays by k_dyn itself in the delta step)
    returns exactly what the device-buffer step returns."""
    rows = gen_pool(16, [(4, 4), (6, 6)], RULES_FULL, seed0=500)
    for n in (40000, 300):                      # split pipeline / fused kernel
        a = make_env(rows, n, traceback=False, max_steps=20, autoreset_mode="same_step", obs_update=obs_update)
        b = make_env(rows, n, traceback=False, max_steps=20, autoreset_mode="same_step", obs_update=obs_update)
        a.reset(); b.reset()
        rng = np.random.default_rng(n)
        r = torch.empty(n, dtype=torch.float32).pin_memory()
        te = torch.empty(n, dtype=torch.uint8).pin_memory()
        tr = torch.empty(n, dtype=torch.uint8).pin_memory()
        keep = torch.empty(n, dtype=torch.uint8).pin_memory()
        for t in range(30):
            act = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8))
            if t % 9 == 4:
                act[t] = 9                                      # an illegal move (SG.py:1137): not expressible in two bits
            if t < 10:
                a.step_host(act.pin_memory(), r, te, tr)       # a new host buffer every call
            elif t < 25:
                keep.copy_(act)
                a.step_host(keep, r, te, tr)                   # the same buffers
            else:
                a.step_host(act.numpy(), r, te, tr)            # pageable memory: the plain call sequence
            ob, rb, tb, ub, _ = b.step(act.cuda())
            assert torch.equal(r, rb.cpu()) and torch.equal(te, tb.cpu()) and torch.equal(tr, ub.cpu()), (n, t)
            assert all(torch.equal(a.obs[k], ob[k]) for k in ob)


@pytest.mark.parametrize("n", [1, 5, 16, 40])
@pytest.mark.parametrize("mode,validate", [("continue", "never"), ("same_step", "terminal")])
def test_info_host_calls_match_step_plus_validate(n, mode, validate):
    """sparc_step_info_host / sparc_reset_info_host (up to 16 envs: ONE launch of k_info with the arguments in the kernel
    parameters and a completion flag; more: the launch + copy sequence) return exactly what a device step followed by
    sparc_validate of the new state returns -- every field of the host record, the observation, the statistics; with
    illegal actions (-1, 7), resets on chosen puzzles and both validation settings."""
    rows = gen_pool(6, [(3, 3), (4, 4), (6, 6)], RULES_FULL, seed0=300)
    kw = dict(traceback=True, max_steps=18, autoreset_mode=mode, validate=validate)
    a, b = make_env(rows, n, **kw), make_env(rows, n, **kw)
    rng = np.random.default_rng(n)
    idx = rng.integers(0, len(rows), n).astype(np.int32)

    def check(rec, t):
        flags, rmap, rfit = b.validate(region_map=True, region_fit=True)
        torch.cuda.synchronize()
        assert np.array_equal(rec["rule_flags"], flags.cpu().numpy().view(np.uint16)), t
        assert np.array_equal(rec["region_map"], rmap.cpu().numpy()) and np.array_equal(rec["region_fit"], rfit.cpu().numpy()), t
        for k, ten in (("current_step", b.step_count), ("agent_location", b.agent_xy), ("reward_code", b.reward_code),
                       ("terminated", b.terminated), ("truncated", b.truncated), ("legal_mask", b.legal_mask)):
            assert np.array_equal(rec[k], ten.cpu().numpy()), (t, k)
        assert all(torch.equal(a.obs[k], b.obs[k]) for k in a.obs), t

    rec = a.reset_info_host(idx)
    b.reset(puzzle_idx=idx)
    check(rec, "reset")
    for t in range(60):
        act = rng.integers(0, 4, n).astype(np.int32)
        if t % 7 == 3:
            act[rng.integers(0, n)] = -1 if t % 2 else 7
        rec = a.step_info_host(act)
        b.step(torch.from_numpy(act).cuda())
        check(rec, t)
        if t == 30:
            rec = a.reset_info_host(None)                   # NULL = every env's current puzzle
            b.reset()
            check(rec, "reset 2")
    assert torch.equal(a.stats(), b.stats()) and int(a.stats()[4]) == 60 * n


def test_long_episodes_hit_max_steps_like_the_emulation():
    """1,100 steps with max_steps=1000 and traceback: deep move stacks, the `current_step >= max_steps`
    truncation (SG.py:1134) and the cyclic puzzle cursor wrapping around the pool, GPU vs the CPU build."""
    from emu.emu_env import EmuVecEnv
    rows = gen_pool(7, [(3, 3), (4, 4), (6, 6)], RULES_FULL, seed0=900)
    N, T = 1024, 1100
    kw = dict(traceback=True, max_steps=1000, autoreset_mode="next_step", validate="terminal", puzzle_stride=5)
    g, c = make_env(rows, N, **kw), EmuVecEnv(rows, N, tile=32, **kw)
    g.reset(); c.reset()
    rng = np.random.default_rng(11)
    hit_max = 0
    for t in range(T):
        a = rng.integers(0, 4, N).astype(np.uint8)
        go, gr, gt, gu, gi = g.step(torch.from_numpy(a).cuda())
        co, cr, ct, cu, ci = c.step(a)
        if t % 50 == 0 or t > 990:
            for k in go:
                assert np.array_equal(to_np(go[k]), co[k]), (t, k)
        assert np.array_equal(to_np(gi["reward_code"]), ci["reward_code"]), t
        assert np.array_equal(to_np(gt), ct) and np.array_equal(to_np(gu), cu), t
        assert np.array_equal(to_np(gi["current_step"]), ci["current_step"]), t
        hit_max += int(((ci["current_step"] >= 1000) & (cu == 1)).sum())
    assert hit_max > 0, "no env survived to max_steps: the truncation path was not exercised"
    assert np.array_equal(to_np(g.stats()), c.stats())


def test_vector_env_face_reset_options():
    from sparc_gym_b200.vec_env import SparcVectorEnv
    rows = gen_pool(6, [(3, 3), (4, 4)], RULES_FULL, seed0=40)
    env = SparcVectorEnv(rows, 12, max_steps=20, autoreset_mode="next_step")
    obs, info = env.reset(options={"puzzle_id": rows[4]["id"]})
    one = {k: v[0].clone() for k, v in obs.items()}
    assert all(torch.equal(v[i], one[k]) for k, v in obs.items() for i in range(12))       # every env on puzzle 4
    obs, info = env.reset(seed=3)
    want = np.random.default_rng(3).integers(6, size=12)
    solo = SparcVectorEnv(rows, 12, max_steps=20)
    o2, _ = solo.reset(options={"puzzle_index": torch.tensor(want, dtype=torch.int32)})
    assert all(torch.equal(obs[k], o2[k]) for k in obs)
    assert env.single_observation_space is not None and env.num_envs == 12 and env.plane_keys[:4] == ["visited", "gaps", "agent_location", "target_location"]
    out = env.step(torch.zeros(12, dtype=torch.int64, device="cuda"))
    assert len(out) == 5 and out[1].shape == (12,)
    # the reference's cursor rules (SG.py:1075-1087), per env: no seed / no options advances the cursor by one,
    # an unknown puzzle_id keeps the env's current puzzle, a seed is ignored when options are given
    env.reset(options={"puzzle_index": torch.arange(12, dtype=torch.int32) % 6})
    obs, _ = env.reset()
    o2, _ = solo.reset(options={"puzzle_index": (torch.arange(12, dtype=torch.int32) + 1) % 6})
    assert all(torch.equal(obs[k], o2[k]) for k in obs)
    obs, _ = env.reset(seed=5, options={"puzzle_id": "no-such-puzzle"})
    assert all(torch.equal(obs[k], o2[k]) for k in obs)
    with pytest.raises(ValueError):
        env.reset(options={"puzzle_index": np.array([0] * 11 + [6], dtype=np.int32)})


def test_dense_polyomino_boards_on_the_gpu(monkeypatch):
    """Dense polyomino boards (1e2..1e4 exact-fit search nodes each) through the deferred validation queue:
    flags and statistics must equal the CPU build of the same code."""
    from emu.emu_env import EmuVecEnv
    from fuzzrows import solution_actions
    from sparc_gym_b200.dataset import gen_dense_poly_puzzle
    monkeypatch.setenv("SPARC_SPLIT", "1")
    rows = [gen_dense_poly_puzzle(6, 6, s) for s in range(24)] + [gen_dense_poly_puzzle(5, 5, s) for s in range(8)]
    rows += gen_pool(8, [(4, 4), (6, 6)], RULES_FULL, seed0=500)
    N = 640                                   # 16 envs per puzzle
    kw = dict(traceback=False, max_steps=200, autoreset_mode="disabled", validate="terminal")
    g = make_env(rows, N, **kw)
    c = EmuVecEnv(rows, N, tile=32, **kw)
    g.reset(); c.reset()
    sols = [solution_actions(r, 0) for r in rows]
    T = max(map(len, sols))
    for t in range(T):
        a = np.array([sols[i % len(rows)][t] if t < len(sols[i % len(rows)]) else 0 for i in range(N)], np.uint8)
        go, gr, gt, gu, gi = g.step(torch.from_numpy(a).cuda())
        co, cr, ct, cu, ci = c.step(a)
        assert np.array_equal(to_np(gi["rule_flags"]).astype(np.uint16), ci["rule_flags"]), t
        assert np.array_equal(to_np(gt), ct) and np.array_equal(to_np(gi["reward_code"]), ci["reward_code"]), t
    st = to_np(g.stats())
    assert np.array_equal(st, c.stats())
    assert st[0] == N and st[1] == N and 0 < st[7] < N      # every env walked its listed solution; some boards tile


@pytest.mark.parametrize("n_envs,tile", [(333, None), (9000, None), (333, "20"), (9000, "36"), (9000, "128"), (333, "128")])
def test_compact_observation_format_on_the_gpu(n_envs, tile, monkeypatch):
    """obs_format='compact' (uint8 / uint8 / int32) holds the same values as the reference-dtype observation,
    element for element, through the fused kernel (333 envs) and the split k_dyn / k_encode step (9,000 envs);
    sparc_encode rewrites it from the state alike.  Tile overrides that are not multiples of 16 are rounded up (16-byte
    stores of uint8 streams), the largest tile needs more than 48 KB of dynamic shared memory."""
    if tile:
        monkeypatch.setenv("SPARC_TILE", tile)
        monkeypatch.setenv("SPARC_ENC_TILE", tile)
    rows = gen_pool(12, [(3, 3), (4, 4), (5, 5), (6, 6)], RULES_FULL, seed0=60)
    kw = dict(traceback=True, max_steps=25, autoreset_mode="same_step", validate="terminal")
    a = make_env(rows, n_envs, **kw)
    b = make_env(rows, n_envs, obs_format="compact", **kw)
    assert b.obs["base"].dtype == torch.uint8 and b.obs["additional_info"].dtype == torch.int32
    a.reset(); b.reset()
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(45):
        for k in a.obs:
            assert torch.equal(a.obs[k].to(torch.int64), b.obs[k].to(torch.int64)), (t, k)
        act = torch.randint(0, 4, (n_envs,), dtype=torch.uint8, device="cuda", generator=gen)
        a.step(act); b.step(act)
    snap = {k: v.clone() for k, v in b.obs.items()}
    for v in b.obs.values():
        v.fill_(77)
    b.encode()
    for k in snap:
        assert torch.equal(snap[k], b.obs[k]), k
    assert torch.equal(a.stats(), b.stats())


@pytest.mark.parametrize("fmt", ["reference", "compact"])
@pytest.mark.parametrize("n_envs,mode", [(70, "next_step"), (8192 + 37, "same_step"), (20000, "disabled")])
def test_delta_observation_equals_full_rewrite_on_the_gpu(fmt, n_envs, mode):
    """obs_update='delta' against the full rewrite, tensor for tensor after every step: the fused kernel's per-tile
    decision (70 envs) and k_dyn + k_encode_list of the split step (>= 8,192 envs), with a masked reset and a
    checkpoint restore in between (see tests/test_emu_parity.py for the CPU twin of this test)."""
    rows = gen_pool(6, [(1, 1), (2, 3), (4, 4), (6, 6)], RULES_FULL, seed0=33)
    kw = dict(traceback=True, max_steps=18, autoreset_mode=mode, validate="terminal", obs_format=fmt)
    a, b = make_env(rows, n_envs, **kw), make_env(rows, n_envs, obs_update="delta", **kw)
    a.reset(); b.reset()
    rng = np.random.default_rng(12)
    ck = None
    for t in range(50):
        if t == 20:
            m = torch.from_numpy(rng.integers(0, 2, n_envs).astype(np.uint8)).cuda()
            pz = torch.from_numpy(rng.integers(0, len(rows), n_envs).astype(np.int32)).cuda()
            a.reset(pz, m); b.reset(pz, m)
        if t == 30:
            ck = (a.state_dict(), b.state_dict())
        if t == 40:
            a.load_state_dict(ck[0]); b.load_state_dict(ck[1])
            a.encode()
        act = torch.from_numpy(rng.integers(0, 4, n_envs).astype(np.uint8)).cuda()
        ra, rb = a.step(act), b.step(act)
        for k in a.obs:
            assert torch.equal(a.obs[k], b.obs[k]), (t, k)
        assert torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2]) and torch.equal(ra[3], rb[3]), t
        assert torch.equal(ra[4]["rule_flags"], rb[4]["rule_flags"]) and torch.equal(ra[4]["legal_mask"], rb[4]["legal_mask"]), t
    assert torch.equal(a.stats(), b.stats())

