"""Live differential test: oracle vs the UNMODIFIED reference, run whenever /root/reference
exists (the build container).  Skipped on the GPU box, where the committed golden vectors
(test_oracle_golden.py) carry the same evidence."""
import pytest

import refharness as rh
import sparc_oracle as so
from fuzzrows import fuzz_actions, fuzz_row
from tracing import diff_traces, run_episode
from sparc_gym_b200.dataset import RULES_FULL, gen_puzzle

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("traceback", [False, True])
def test_fuzz_against_reference(traceback):
    rows = [fuzz_row(5000 + s, yaml_style="block" if s % 3 else "json") for s in range(45)]
    rows += [gen_puzzle(3, 3, 11, RULES_FULL), gen_puzzle(5, 4, 12, RULES_FULL, extra_solutions=2)]
    ref = rh.make_reference_env(rows, observation="new", traceback=traceback, max_steps=50)
    orc = so.OracleEnv(rows, observation="new", traceback=traceback, max_steps=50)
    steps = 0
    for i, row in enumerate(rows):
        acts = fuzz_actions(row, 17 * i + traceback, 60)
        a = run_episode(ref, row["id"], acts, pre_reset=rh.zero_leak)
        b = run_episode(orc, row["id"], acts)
        d = diff_traces(a, b)
        assert d is None, f"{row['id']}: {d}"
        steps += len(a)
    assert steps > 500


def test_sparc_text_observation_matches_reference():
    rows = [gen_puzzle(3, 3, 21, RULES_FULL), gen_puzzle(4, 4, 22, RULES_FULL)]
    for tb in (False, True):
        ref = rh.make_reference_env(rows, observation="SPaRC", traceback=tb, max_steps=30)
        orc = so.OracleEnv(rows, observation="SPaRC", traceback=tb, max_steps=30)
        for i, row in enumerate(rows):
            rh.zero_leak(ref)
            o1, _ = ref.reset(options={"puzzle_id": row["id"]})
            o2, _ = orc.reset(options={"puzzle_id": row["id"]})
            assert o1 == o2
            for a in fuzz_actions(row, i, 30, p_follow=0.3, n_act=4):
                o1, r1, te1, tr1, _ = ref.step(a)
                o2, r2, te2, tr2, _ = orc.step(a)
                assert (o1, r1, te1, tr1) == (o2, r2, te2, tr2)
                if te1 or tr1:
                    break
