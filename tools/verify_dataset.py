#!/usr/bin/env python
"""All-rows sweep over a locally cached copy of the SPaRC dataset: ingest every row the way
`_process_puzzles` does (SG.py:219-368), upload the pool, replay EVERY listed solution of every puzzle on the GPU in one
batched call (verify.verify_paths) and report

  * rows the B200 store refuses, with the reason (limits: lattice <= 15x15, <= 32 distinct shapes and <= 6 ylop
    instances per puzzle, unknown symbol kinds; rows the reference itself crashes on, e.g. an instance without a
    'poly' plane) -- such rows are dropped from the pool, the sweep goes on;
  * per listed solution: +0.01 on every move and +1 / terminated at the end (SG.py:1205-1221: the reward comes from
    the listed paths), and whether the rule validator agrees that all rules are satisfied (the dataset's own claim).

    python tools/verify_dataset.py path/to/sparc.{parquet,json,jsonl} [--report out.json] [--device cuda:0]
    (run under gpurun; there is no CPU fallback)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sweep(rows, device="cuda:0"):
    import numpy as np
    from sparc_gym_b200.ingest import ingest_rows
    from sparc_gym_b200.store import PuzzleStore
    from sparc_gym_b200.verify import verify_paths
    # rows one by one first: a bad row must not take the pool down (the stale-`symbol` quirks Q11/Q12 make the
    # planes of a row depend on the rows before it, so the accepted rows are re-ingested together afterwards)
    rejects, keep = [], []
    for i, row in enumerate(rows):
        try:
            PuzzleStore(ingest_rows(rows[max(0, i - 1):i + 1])[-1:])
            keep.append(i)
        except Exception as e:                                   # noqa: BLE001 -- the report wants every reason
            rejects.append({"row": i, "id": str(row.get("id", i)), "error": f"{type(e).__name__}: {e}"[:300]})
    good = [rows[i] for i in keep]
    report = {"rows": len(rows), "accepted": len(good), "rejected": rejects}
    if not good:
        return report
    puzzles = ingest_rows(good)
    store = PuzzleStore(puzzles)
    report["slab"] = [store.slab_x, store.slab_y]
    report["planes"] = list(store.plane_keys)
    idx, paths = [], []
    for k, pz in enumerate(puzzles):
        for s in range(pz.solution_count):
            idx.append(k)
            paths.append(pz.solution_paths[s])
    report["listed_solutions"] = len(paths)
    if paths:
        res = verify_paths(store, idx, paths, device=device)
        lens = np.array([len(p) - 1 for p in paths])
        ok = res["solved"] & res["terminated"] & (res["steps"] == lens) & (res["return_cents"] == lens - 1 + 100)
        report["solutions_rewarded_plus_one"] = int(ok.sum())
        report["solutions_satisfying_all_rules"] = int(res["all_rules_satisfied"].sum())
        bad = np.nonzero(~ok)[0][:50]
        report["reward_failures"] = [{"id": puzzles[idx[b]].id, "solved": bool(res["solved"][b]), "terminated": bool(res["terminated"][b]),
                                      "steps": int(res["steps"][b]), "path_moves": int(lens[b])} for b in bad]
        worse = np.nonzero(ok & ~res["all_rules_satisfied"])[0][:50]
        report["rule_disagreements"] = [{"id": puzzles[idx[b]].id,
                                         "failed": [n for n in res if n.endswith(("_target", "_crossing", "_violations", "_collected", "_separation",
                                                                                   "_exact", "_count", "_area")) and not res[n][b]]} for b in worse]
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--report", default=None)
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args()
    from sparc_gym_b200.dataset import load_rows
    rows = load_rows(a.path)
    rep = sweep(rows, a.device)
    text = json.dumps(rep, indent=1)
    if a.report:
        with open(a.report, "w") as f:
            f.write(text)
    print(text if len(text) < 6000 else text[:6000] + "\n... (see --report)")
    return 0 if not rep["rejected"] and rep.get("solutions_rewarded_plus_one") == rep.get("listed_solutions") else 1


if __name__ == "__main__":
    sys.exit(main())
