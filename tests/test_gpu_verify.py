"""GPU: batched path verification (the LLM-evaluation caller's env work, llm_host.py:63-242)."""
import numpy as np
import pytest

import sparc_oracle as so
from sparc_gym_b200.dataset import RULES_FULL, gen_pool

pytestmark = pytest.mark.gpu


def test_verify_paths_against_oracle():
    from sparc_gym_b200.verify import verify_paths
    rows = gen_pool(24, [(3, 3), (4, 4), (5, 5)], RULES_FULL, seed0=700, extra_solutions=1)
    rng = np.random.default_rng(5)
    idx, paths = [], []
    for i, r in enumerate(rows):
        sol = [(p["x"], p["y"]) for p in r["solutions"][0]["path"]]
        idx += [i, i, i]
        paths += [sol, sol[:max(2, len(sol) // 2)], sol[:3] + [sol[1]] + sol[3:]]      # solution, prefix, a detour that breaks it
    res = verify_paths(rows, idx, paths)
    assert res["solved"][0::3].all() and res["all_rules_satisfied"][0::3].all() and res["terminated"][0::3].all()
    assert not res["solved"][1::3].any() and not res["reached_target"][1::3].any()
    # cross-check every verdict with the oracle replaying the same actions one env at a time
    from sparc_gym_b200.verify import path_to_actions
    env = so.OracleEnv(rows, max_steps=10 ** 6)
    for k, (i, p) in enumerate(zip(idx, paths)):
        env.reset(options={"puzzle_id": rows[i]["id"]})
        total, n, te, tr, info = 0, 0, False, False, env._info()
        for a in path_to_actions(p):
            _, r, te, tr, info = env.step(a)
            total += round(r * 100)
            n += 1
            if te or tr:
                break
        assert res["return_cents"][k] == total and res["steps"][k] == n, k
        assert bool(res["terminated"][k]) == te and bool(res["truncated"][k]) == tr, k
        for name in so.RULE_NAMES:
            assert bool(res[name][k]) == info["rule_status"][name]["passed"], (k, name)


def test_all_rows_sweep_over_dataset_files(tmp_path):
    """tools/verify_dataset.py on JSON / JSONL / parquet copies of a pool in the dataset's schema: the refused row is
    reported, every listed solution of the generated rows earns +1 on the GPU (CPU twin: tests/test_dropin_host_logic.py)."""
    import importlib.util
    import os
    from sparc_gym_b200.dataset import load_rows
    from test_dropin_host_logic import _dataset_files
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("verify_dataset", os.path.join(root, "tools", "verify_dataset.py"))
    vd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vd)
    rows, paths = _dataset_files(tmp_path)
    for ext, p in paths.items():
        rep = vd.sweep(load_rows(p), device="cuda:0")
        assert rep["accepted"] == len(rows) and [r["id"] for r in rep["rejected"]] == ["too-wide"], ext
        assert rep["solutions_rewarded_plus_one"] >= 20 and rep["solutions_satisfying_all_rules"] >= 10
        assert all(f["id"].startswith("fuzz") for f in rep["reward_failures"])
