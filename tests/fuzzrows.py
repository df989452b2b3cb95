"""Adversarial puzzle rows + action streams for the differential tests.

Unlike `sparc_gym_b200.dataset.gen_puzzle` (which plants a rule-satisfying solution) these
rows are deliberately messy: random typed cells (so most rule checks FAIL in interesting
ways), triangle counts that collide with polyshape ids (Q10), stacked ylops, shapes that do
not fit, random start / end, 0-3 listed solutions (some not ending at the target, some
shorter), and cells lists ordered so that the reference's stale-`symbol` ingest quirks
(Q11/Q12) are exercised without crashing it.
"""
from __future__ import annotations

import json

import numpy as np
import yaml

COLORS = ["red", "blue", "yellow", "green", "black", "purple", "orange", "white"]
SMALL_SHAPES = [
    [[1]], [[1, 1]], [[1], [1]], [[1, 1], [1, 0]], [[1, 1], [0, 1]], [[1, 1, 1]], [[1], [1], [1]],
    [[1, 1], [1, 1]], [[0, 1], [1, 1]], [[1, 0], [1, 1], [1, 0]], [[0, 1, 0], [1, 1, 1]],
    [[0, 0], [0, 0]],
]
DIRS = ((1, 0), (0, -1), (-1, 0), (0, 1))


def _walk(X, Y, start, blocked, rng, target=None, max_len=200):
    """Random self-avoiding node walk from `start`; stops at `target`, at a dead end or at max_len."""
    path = [start]
    seen = {start}
    while len(path) < max_len:
        x, y = path[-1]
        if target is not None and (x, y) == target:
            break
        opts = []
        for dx, dy in DIRS:
            n = (x + dx, y + dy)
            if 0 <= n[0] < X and 0 <= n[1] < Y and n not in seen and n not in blocked \
                    and not (n[0] % 2 and n[1] % 2):
                opts.append(n)
        if not opts:
            break
        if target is not None and rng.random() < 0.6:   # drift towards the target
            opts.sort(key=lambda n: abs(n[0] - target[0]) + abs(n[1] - target[1]))
            n = opts[0]
        else:
            n = opts[int(rng.integers(len(opts)))]
        path.append(n)
        seen.add(n)
    return path


def fuzz_row(seed, max_cells=4, n_colors=4, collide=True, yaml_style="block"):
    rng = np.random.default_rng([int(seed), 0xF022])
    w, h = int(rng.integers(1, max_cells + 1)), int(rng.integers(1, max_cells + 1))
    X, Y = 2 * w + 1, 2 * h + 1
    walk_nodes = [(x, y) for x in range(X) for y in range(Y) if not (x % 2 and y % 2)]
    edge_nodes = [n for n in walk_nodes if (n[0] + n[1]) % 2 == 1]
    gaps = [n for n in edge_nodes if rng.random() < 0.12]
    free = [n for n in walk_nodes if n not in gaps]
    start = free[int(rng.integers(len(free)))]
    end = start
    while end == start and len(free) > 1:
        end = free[int(rng.integers(len(free)))]
    dots = [n for n in walk_nodes if rng.random() < 0.08 and n not in gaps]

    # shape table: ids 1,2,3 collide with triangle counts on purpose
    n_shapes = int(rng.integers(0, 5))
    shapes = {}
    for k in range(n_shapes):
        sid = (k + 1) if collide else (20 + k)
        shapes[str(sid)] = SMALL_SHAPES[int(rng.integers(len(SMALL_SHAPES)))]
    typed = []
    kinds = ["star", "square", "triangle"] + (["poly", "ylop"] if shapes else [])
    has_poly = False
    for x in range(1, X, 2):
        for y in range(1, Y, 2):
            if rng.random() < 0.55:
                kind = kinds[int(rng.integers(len(kinds)))]
                props = {"type": kind}
                if rng.random() < 0.92:
                    props["color"] = COLORS[int(rng.integers(n_colors))]
                if kind == "triangle":
                    props["count"] = int(rng.integers(0, 5))
                elif kind in ("poly", "ylop"):
                    props["polyshape"] = int(list(shapes)[int(rng.integers(len(shapes)))])
                    has_poly |= kind == "poly"
                typed.append({"position": {"x": x, "y": y}, "properties": props})
    # the reference raises KeyError('poly') when a shape instance exists without any poly
    # plane (SG.py:734): keep the row runnable by forcing one poly cell in that case
    def is_instance(c):
        p = c["properties"]
        v = p.get("count") if p["type"] == "triangle" else p.get("polyshape")
        return bool(v) and str(v) in shapes
    if any(is_instance(c) for c in typed) and not has_poly:
        for c in typed:
            if is_instance(c):
                c["properties"] = {"type": "poly", "color": c["properties"].get("color", "red"),
                                   "polyshape": int(list(shapes)[0])}
                break
    cells = list(typed)
    cells += [{"position": {"x": x, "y": y}, "properties": {"dot": True}} for x, y in dots]
    gap_cells = [{"position": {"x": x, "y": y}, "properties": {"gap": True}} for x, y in gaps]
    if cells:
        # interleave some gaps (stale-symbol path of Q11) but never put one first (Q12 crash)
        for g in gap_cells:
            cells.insert(int(rng.integers(1, len(cells) + 1)), g)
    # listed solutions: random walks; most reach `end`, some do not
    sols = []
    for _ in range(int(rng.integers(0, 4))):
        p = _walk(X, Y, start, set(gaps), rng, target=end if rng.random() < 0.85 else None,
                  max_len=int(rng.integers(2, 3 * (X + Y))))
        if p not in sols:
            sols.append(p)
    text = {"puzzle": {"start": {"x": start[0], "y": start[1]}, "end": {"x": end[0], "y": end[1]},
                       "cells": cells}}
    tv = yaml.safe_dump(text, sort_keys=False) if yaml_style == "block" else json.dumps(text)
    return {
        "id": f"fuzz-{int(seed):06d}",
        "difficulty_level": int(rng.integers(1, 6)),
        "grid_size": {"width": w, "height": h},
        "solution_count": len(sols),
        "solutions": [{"path": [{"x": int(x), "y": int(y)} for x, y in s]} for s in sols],
        "polyshapes": json.dumps(shapes),
        "text_visualization": tv,
        "puzzle_array": [["+"] * X for _ in range(Y)],
    }


def solution_actions(row, k=0):
    p = [(q["x"], q["y"]) for q in row["solutions"][k]["path"]]
    return [DIRS.index((b[0] - a[0], b[1] - a[1])) for a, b in zip(p, p[1:])]


def fuzz_actions(row, seed, T, p_follow=0.6, n_act=5):
    """Action stream that follows a listed solution with prob. p_follow, else uniform in
    0..n_act-1 (4 = an invalid action id)."""
    rng = np.random.default_rng([int(seed), 0xAC7])
    acts = rng.integers(0, n_act, size=T)
    if row["solutions"]:
        k = int(rng.integers(len(row["solutions"])))
        sol = solution_actions(row, k)
        follow = rng.random(T) < p_follow
        j = 0
        for t in range(T):
            if follow[t] and j < len(sol):
                acts[t] = sol[j]
                j += 1
            else:
                break
    return [int(a) for a in acts]
