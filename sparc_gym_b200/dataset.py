"""Synthetic SPaRC puzzle rows (and a loader for locally cached rows).

There is no network, so the HF dataset the reference downloads
(`load_dataset('lkaesberg/SPaRC', ...)`, reference SPaRC_Gym/SPaRC_Gym.py:77-78)
is replaced by a seeded "planted solution" generator that emits rows in the
*same column schema* the reference reads in `_process_puzzles`
(SPaRC_Gym/SPaRC_Gym.py:239-363):

    id, difficulty_level, grid_size{width,height}, solution_count,
    solutions[{path:[{x,y}...]}], polyshapes (JSON/YAML string),
    text_visualization (YAML string: puzzle.start / puzzle.end / puzzle.cells),
    puzzle_array ([y][x] symbol grid, only read for observation='SPaRC').

A `w x h`-cell puzzle lives on a `(2w+1) x (2h+1)` node lattice indexed [x, y]
(SPaRC_Gym.py:243-248); cells are the (odd, odd) nodes.

Generator recipe: random self-avoiding vertex walk
start -> end, gap edges off the walk, dots on the walk, then per region
(reference flood-fill semantics, SPaRC_Gym.py:422-454) symbols that the planted
walk satisfies: one square colour per region, star pairs, triangles whose
count equals the touched edges, a polyomino partition of the region (shape
arrays stored with index 0 = grid x, SPaRC_Gym.py:861-870) and optional
poly+ylop cancelling pairs.  `properties` always lists `type` first (the
reference's ingest depends on YAML key order, SPaRC_Gym.py:288-306).
"""
from __future__ import annotations

import json
from collections import deque

import numpy as np
import yaml

_Dumper = getattr(yaml, "CSafeDumper", yaml.SafeDumper)
_Loader = getattr(yaml, "CSafeLoader", yaml.SafeLoader)

COLOR_NAMES = ["red", "blue", "yellow", "green", "black", "purple", "orange", "white"]

#: rule mixes used by the benchmark configs (BASELINE.json `configs`)
RULES_GAPS_DOTS = dict(gaps=3, dots=2)
RULES_FULL = dict(gaps=2, dots=2, stars=True, squares=True, triangles=True,
                  polys=True, ylops=True, colors=8)

_DIRS = ((1, 0), (0, -1), (-1, 0), (0, 1))  # right, up, left, down on [x, y]


def _vertex_walk(w, h, start, end, rng):
    """Random self-avoiding walk on the (w+1) x (h+1) vertex grid (randomised DFS)."""
    sv = (start[0] // 2, start[1] // 2)
    ev = (end[0] // 2, end[1] // 2)
    seen = {sv}
    stack = [(sv, list(rng.permutation(4)))]
    while stack:
        (vx, vy), order = stack[-1]
        if (vx, vy) == ev:
            return [v for v, _ in stack]
        moved = False
        while order:
            d = int(order.pop())
            nx, ny = vx + _DIRS[d][0], vy + _DIRS[d][1]
            if 0 <= nx <= w and 0 <= ny <= h and (nx, ny) not in seen:
                seen.add((nx, ny))
                stack.append(((nx, ny), list(rng.permutation(4))))
                moved = True
                break
        if not moved:
            stack.pop()  # dead end: backtrack (vertex stays in `seen`)
    raise RuntimeError("no walk found")  # unreachable on a rectangular grid


def _node_path(vwalk):
    """Vertex walk -> node path (vertex, edge midpoint, vertex, ...)."""
    out = []
    for i, (vx, vy) in enumerate(vwalk):
        if i:
            px, py = vwalk[i - 1]
            out.append((px + vx, py + vy))  # midpoint of (2px,2py)-(2vx,2vy)
        out.append((2 * vx, 2 * vy))
    return out


def lattice_regions(X, Y, blocked):
    """Connected components of rule cells under the reference's lattice semantics.

    A node is passable iff it is a rule cell (odd, odd) or not in `blocked`
    (= gaps U path); components are listed in raster order (x outer, y inner)
    of their first cell -- SPaRC_Gym.py:422-454.
    Returns (list of cell lists, region_map[X][Y] with -1 off-cell).
    """
    rmap = [[-1] * Y for _ in range(X)]
    regions = []
    for x in range(1, X, 2):
        for y in range(1, Y, 2):
            if rmap[x][y] != -1:
                continue
            rid = len(regions)
            cells = []
            seen = {(x, y)}
            q = deque([(x, y)])
            while q:
                cx, cy = q.popleft()
                if cx % 2 == 1 and cy % 2 == 1:
                    rmap[cx][cy] = rid
                    cells.append((cx, cy))
                for dx, dy in _DIRS:
                    nx, ny = cx + dx, cy + dy
                    if not (0 <= nx < X and 0 <= ny < Y) or (nx, ny) in seen:
                        continue
                    is_cell = nx % 2 == 1 and ny % 2 == 1
                    if is_cell or (nx, ny) not in blocked:
                        seen.add((nx, ny))
                        q.append((nx, ny))
            cells.sort()
            regions.append(cells)
    return regions, rmap


def _partition_polyominoes(cells, rng, max_piece=5, max_box=4):
    """Randomly partition a set of rule cells into edge-connected pieces (bbox <= max_box)."""
    todo = set(cells)
    pieces = []
    while todo:
        seed = min(todo)
        piece = [seed]
        todo.discard(seed)
        want = int(rng.integers(1, max_piece + 1))
        while len(piece) < want:
            frontier = []
            for (cx, cy) in piece:
                for dx, dy in _DIRS:
                    n = (cx + 2 * dx, cy + 2 * dy)
                    if n in todo:
                        xs = [p[0] for p in piece] + [n[0]]
                        ys = [p[1] for p in piece] + [n[1]]
                        if (max(xs) - min(xs)) // 2 < max_box and (max(ys) - min(ys)) // 2 < max_box:
                            frontier.append(n)
            if not frontier:
                break
            n = frontier[int(rng.integers(len(frontier)))]
            piece.append(n)
            todo.discard(n)
        pieces.append(sorted(piece))
    return pieces


def _shape_array(piece):
    """Cells (node coords) -> 0/1 array with index 0 = x, index 1 = y (reference orientation)."""
    xs = [p[0] for p in piece]
    ys = [p[1] for p in piece]
    x0, y0 = min(xs), min(ys)
    arr = [[0] * ((max(ys) - y0) // 2 + 1) for _ in range((max(xs) - x0) // 2 + 1)]
    for (x, y) in piece:
        arr[(x - x0) // 2][(y - y0) // 2] = 1
    return arr


def shape_id(arr, base=16):
    """Deterministic integer id of a shape array (>= `base`, so it never collides with a
    triangle count 1..4 -- the reference treats any additional_info value whose str() is a
    polyshape key as a shape instance, SPaRC_Gym.py:727-731)."""
    v = 0
    for i, row in enumerate(arr):
        for j, b in enumerate(row):
            if b:
                v |= 1 << (i * 4 + j)
    return base + v


def gen_puzzle(w, h, seed, rules=None, *, start=None, end=None, extra_solutions=0,
               fill=0.6, puzzle_id=None, text_format="yaml"):
    """One synthetic dataset row. `rules` keys: gaps, dots, stars, squares, triangles,
    polys, ylops (bools / counts) and colors (1..8).  `text_format`: 'yaml' writes
    `text_visualization` as block-style YAML like the HF dataset; 'json' writes the same
    document as JSON (a YAML subset the reference's yaml.safe_load reads identically), which
    both generates and ingests an order of magnitude faster -- used for the big bench pools."""
    rules = dict(rules or {})
    rng = np.random.default_rng([int(seed), w, h, 0x5BA2C])
    X, Y = 2 * w + 1, 2 * h + 1
    start = tuple(start) if start is not None else (0, 2 * h)
    end = tuple(end) if end is not None else (2 * w, 0)
    ncol = int(rules.get("colors", 4))

    vwalk = _vertex_walk(w, h, start, end, rng)
    path = _node_path(vwalk)
    pset = set(path)

    # gap edges: edge midpoints not on the walk
    edges = [(x, y) for x in range(X) for y in range(Y) if (x + y) % 2 == 1 and (x, y) not in pset]
    k_gap = min(int(rules.get("gaps", 0)), len(edges))
    gaps = [edges[i] for i in rng.permutation(len(edges))[:k_gap]] if k_gap else []
    # dots: on walk nodes
    k_dot = min(int(rules.get("dots", 0)), len(path))
    dots = [path[i] for i in sorted(rng.permutation(len(path))[:k_dot])] if k_dot else []

    regions, _ = lattice_regions(X, Y, pset | set(gaps))

    cells_out = []   # (x, y, properties) for typed cells
    shapes = {}      # str(id) -> array
    any_poly = False
    n_poly_regions = n_poly_pieces = n_ylop_pairs = 0
    for cells in regions:
        free = [cells[i] for i in rng.permutation(len(cells))]
        used_colors = set()

        def take():
            return free.pop() if free else None

        # polyomino partition first (needs one symbol cell per piece).  Density follows
        # Witness-style puzzles: a few shapes per puzzle, not a tiling of the whole board
        # (rules: poly_regions / poly_pieces / ylop_pairs caps per puzzle).
        if rules.get("polys") and n_poly_regions < int(rules.get("poly_regions", 2)) and rng.random() < 0.6:
            pieces = _partition_polyominoes(cells, rng)
            extra = 2 if (rules.get("ylops") and n_ylop_pairs < int(rules.get("ylop_pairs", 1))
                          and rng.random() < 0.5) else 0
            if len(pieces) + extra <= len(free) and \
                    n_poly_pieces + len(pieces) + extra // 2 <= int(rules.get("poly_pieces", 6)):
                n_poly_regions += 1
                n_poly_pieces += len(pieces) + extra // 2
                n_ylop_pairs += extra // 2
                for piece in pieces:
                    arr = _shape_array(piece)
                    sid = shape_id(arr)
                    shapes[str(sid)] = arr
                    cx, cy = take()
                    col = COLOR_NAMES[int(rng.integers(ncol))]
                    used_colors.add(col)
                    cells_out.append((cx, cy, {"type": "poly", "color": col, "polyshape": sid}))
                    any_poly = True
                if extra:
                    # cancelling pair: a ylop digs shape S somewhere, an extra poly S refills it
                    sw, sh = int(rng.integers(1, 3)), int(rng.integers(1, 3))
                    arr = [[1] * sh for _ in range(sw)]
                    if sw * sh > 2 and rng.random() < 0.5:
                        arr[sw - 1][sh - 1] = 0
                    sid = shape_id(arr)
                    shapes[str(sid)] = arr
                    for kind in ("poly", "ylop"):
                        cx, cy = take()
                        col = COLOR_NAMES[int(rng.integers(ncol))]
                        used_colors.add(col)
                        cells_out.append((cx, cy, {"type": kind, "color": col, "polyshape": sid}))
        # triangles: count = touched walk edges (must be 1..4)
        if rules.get("triangles"):
            for (cx, cy) in list(free):
                touches = sum(((cx + dx, cy + dy) in pset) for dx, dy in _DIRS)
                if touches and rng.random() < 0.25 * fill:
                    free.remove((cx, cy))
                    col = COLOR_NAMES[int(rng.integers(ncol))]
                    used_colors.add(col)
                    cells_out.append((cx, cy, {"type": "triangle", "color": col, "count": touches}))
        # squares: a single colour per region
        sq_col = None
        if rules.get("squares") and free and rng.random() < fill:
            sq_col = COLOR_NAMES[int(rng.integers(ncol))]
            for _ in range(int(rng.integers(1, 3))):
                c = take()
                if c is None:
                    break
                cells_out.append((c[0], c[1], {"type": "square", "color": sq_col}))
            used_colors.add(sq_col)
        # stars: pairs of a colour not otherwise present in the region
        if rules.get("stars") and rng.random() < fill:
            avail = [c for c in COLOR_NAMES[:ncol] if c not in used_colors]
            if avail and len(free) >= 2:
                col = avail[int(rng.integers(len(avail)))]
                for _ in range(2):
                    cx, cy = take()
                    cells_out.append((cx, cy, {"type": "star", "color": col}))

    if any(p["type"] == "ylop" for _, _, p in cells_out) and not any_poly:  # pragma: no cover
        raise AssertionError("ylop without poly would crash the reference (SPaRC_Gym.py:734)")

    # cells list: typed cells first, then dots, then gaps -- never start with a gap-only
    # cell (the reference's `symbol` variable would be unbound / stale, SPaRC_Gym.py:305-306)
    cells_out.sort(key=lambda c: (c[0], c[1]))
    ycells = [{"position": {"x": int(x), "y": int(y)}, "properties": props} for x, y, props in cells_out]
    ycells += [{"position": {"x": int(x), "y": int(y)}, "properties": {"dot": True}} for x, y in dots]
    if not ycells and gaps:
        gaps = []  # a gap-only puzzle cannot be ingested by the reference
    ycells += [{"position": {"x": int(x), "y": int(y)}, "properties": {"gap": True}} for x, y in gaps]

    sols = [path]
    for k in range(int(extra_solutions)):
        alt = _node_path(_vertex_walk(w, h, start, end, rng))
        if not (set(alt) & set(gaps)) and alt not in sols:
            sols.append(alt)

    text = {"puzzle": {"start": {"x": int(start[0]), "y": int(start[1])},
                       "end": {"x": int(end[0]), "y": int(end[1])},
                       "cells": ycells}}
    row = {
        "id": puzzle_id or f"syn-{w}x{h}-{int(seed):08d}",
        "difficulty_level": int(1 + (w + h) // 3),
        "grid_size": {"width": int(w), "height": int(h)},
        "solution_count": len(sols),
        "solutions": [{"path": [{"x": int(x), "y": int(y)} for x, y in s]} for s in sols],
        "polyshapes": json.dumps(shapes),
        "text_visualization": (json.dumps(text) if text_format == "json"
                               else yaml.dump(text, Dumper=_Dumper, sort_keys=False)),
    }
    row["puzzle_array"] = puzzle_array(row, text)
    return row


_SYM = {"star": "*", "square": "o", "triangle": "A", "poly": "P", "ylop": "Y"}
_COL = {"red": "R", "blue": "B", "yellow": "Y", "green": "G", "black": "K", "purple": "P",
        "orange": "O", "white": "W"}


def puzzle_array(row, text=None):
    """[y][x] symbol grid in the SPaRC text convention (legend: reference
    llm_testing/llm_host.py:102-120): '+' walkable node, 'N' empty rule cell, 'S'/'E' start /
    end, '.' dot, 'G' gap, '<sym>-<colour>' typed cells."""
    X = 2 * row["grid_size"]["width"] + 1
    Y = 2 * row["grid_size"]["height"] + 1
    tv = (text or yaml.load(row["text_visualization"], Loader=_Loader))["puzzle"]
    g = [["N" if (x % 2 and y % 2) else "+" for x in range(X)] for y in range(Y)]
    for c in tv["cells"]:
        x, y = c["position"]["x"], c["position"]["y"]
        p = c["properties"]
        if "type" in p:
            s = _SYM.get(p["type"], "?")
            if p["type"] == "triangle":
                s += str(p.get("count", ""))
            if p["type"] in ("poly", "ylop"):
                s += str(p.get("polyshape", ""))
            g[y][x] = f"{s}-{_COL.get(p.get('color', ''), 'X')}"
        elif "dot" in p:
            g[y][x] = "."
        elif "gap" in p:
            g[y][x] = "G"
    g[tv["start"]["y"]][tv["start"]["x"]] = "S"
    g[tv["end"]["y"]][tv["end"]["x"]] = "E"
    return g


SMALL_PIECES = [[[1, 1]], [[1], [1]], [[1, 1, 1]], [[1], [1], [1]], [[1, 1], [1, 0]], [[1, 1], [0, 1]],
                [[1, 0], [1, 1]], [[0, 1], [1, 1]], [[1, 1], [1, 1]], [[1, 1, 1], [0, 1, 0]], [[1, 1, 1, 1]],
                [[1, 0], [1, 0], [1, 1]], [[1, 1, 0], [0, 1, 1]], [[0, 1, 1], [1, 1, 0]], [[1], [1], [1], [1]],
                [[1, 1, 1], [1, 0, 0]]]


def gen_dense_poly_puzzle(w, h, seed, puzzle_id=None):
    """A stress puzzle for the exact-fit search (SPaRC_Gym.py:738-895): the listed solution runs along the
    border (start bottom-left, up, then right to the top-right end), so the whole w x h board is ONE region,
    and it holds random small polyominoes of total area w*h -- about a dozen distinct pieces on 6x6.  Whether
    they tile the board is what the search has to find out (1e3..1e5 nodes; most draws have no tiling).
    Denser than anything `gen_puzzle` emits, on purpose."""
    rng = np.random.default_rng([int(seed), 0xDE5E])
    polys, tot = [], 0
    while tot < w * h:
        q = SMALL_PIECES[int(rng.integers(len(SMALL_PIECES)))]
        a = int(np.array(q).sum())
        if tot + a <= w * h and len(q) <= w and len(q[0]) <= h:
            polys.append(q)
            tot += a
        elif w * h - tot == 1:
            polys.append([[1]])
            tot += 1
    ids, cells = {}, []
    for k, q in enumerate(polys):
        sid = ids.setdefault(json.dumps(q), 20 + len(ids))
        cx, cy = divmod(k, h)
        cells.append({"position": {"x": 2 * cx + 1, "y": 2 * cy + 1},
                      "properties": {"type": "poly", "color": COLOR_NAMES[k % 8], "polyshape": sid}})
    path = [(0, y) for y in range(2 * h, -1, -1)] + [(x, 0) for x in range(1, 2 * w + 1)]
    text = {"puzzle": {"start": {"x": 0, "y": 2 * h}, "end": {"x": 2 * w, "y": 0}, "cells": cells}}
    row = {"id": puzzle_id or f"dense-{w}x{h}-{int(seed):06d}", "difficulty_level": 5,
           "grid_size": {"width": int(w), "height": int(h)}, "solution_count": 1,
           "solutions": [{"path": [{"x": x, "y": y} for x, y in path]}],
           "polyshapes": json.dumps({str(v): json.loads(k) for k, v in ids.items()}),
           "text_visualization": json.dumps(text)}
    row["puzzle_array"] = puzzle_array(row, text)
    return row


def gen_many_ylops_puzzle(seed, solvable=True, puzzle_id=None):
    """A 3x3 board that is ONE region (the listed solution runs along the border) holding a full 3x3 poly, a second poly
    of seven cells and SEVEN one-cell ylops: more ylop instances than the fast exact-fit search's 4-bit counters serve
    (store flag bit 2 -> 6-bit counters).  It tiles: the ylops sit exactly under the second poly.  `solvable=False`
    gives the second poly an eighth cell, so the areas no longer balance (9 + 8 - 7 != 9) and the rule fails."""
    rng = np.random.default_rng([int(seed), 0x7109])
    full = [[1, 1, 1], [1, 1, 1], [1, 1, 1]]
    holes = [(0, 0), (0, 2), (2, 0), (2, 2)]
    hx, hy = holes[int(rng.integers(len(holes)))]
    second = [[0 if (i, j) == (hx, hy) else 1 for j in range(3)] for i in range(3)]          # 8 cells
    if solvable:
        second[1][1] = 0                                                                     # 7 cells (still 4-connected)
    ylops = [[[1]]] * 7
    shapes = {"31": full, "32": second, "33": [[1]], "34": [[1, 1]]}
    cells, k = [], 0
    inst = [("poly", 31), ("poly", 32)] + [("ylop", 33 if y == [[1]] else 34) for y in ylops]
    for kind, sid in inst:
        cx, cy = divmod(k, 3)
        k += 1
        cells.append({"position": {"x": 2 * cx + 1, "y": 2 * cy + 1},
                      "properties": {"type": kind, "color": COLOR_NAMES[k % 8], "polyshape": sid}})
    w = h = 3
    path = [(0, y) for y in range(2 * h, -1, -1)] + [(x, 0) for x in range(1, 2 * w + 1)]
    text = {"puzzle": {"start": {"x": 0, "y": 2 * h}, "end": {"x": 2 * w, "y": 0}, "cells": cells}}
    row = {"id": puzzle_id or f"ylops-{int(seed):04d}-{'ok' if solvable else 'no'}", "difficulty_level": 5,
           "grid_size": {"width": w, "height": h}, "solution_count": 1,
           "solutions": [{"path": [{"x": x, "y": y} for x, y in path]}],
           "polyshapes": json.dumps(shapes), "text_visualization": json.dumps(text)}
    row["puzzle_array"] = puzzle_array(row, text)
    return row


def gen_pool(n, sizes, rules, seed0=0, **kw):
    """`n` rows; sizes is a list of (w, h) cycled over; row i uses seed seed0+i."""
    return [gen_puzzle(*sizes[i % len(sizes)], seed0 + i, rules, **kw) for i in range(n)]


def config_pool(config, n_puzzles=None, text_format="json"):
    """Puzzle pools of the five BASELINE.json configs."""
    kw = dict(text_format=text_format)
    if config in (1, "1"):
        return [gen_puzzle(4, 4, 0, RULES_GAPS_DOTS, **kw)]
    if config in (2, "2"):
        return gen_pool(n_puzzles or 64, [(4, 4)], RULES_GAPS_DOTS, **kw)
    if config in (3, "3", 5, "5"):
        return gen_pool(n_puzzles or 1024, [(3, 3), (4, 4), (5, 5), (6, 6)], RULES_FULL, seed0=1000, **kw)
    if config in (4, "4"):
        return gen_pool(n_puzzles or 4096, [(6, 6)], RULES_FULL, seed0=100000, **kw)
    raise ValueError(f"unknown config {config!r}")


def load_rows(path):
    """Rows from a locally cached copy of the SPaRC dataset: .json (list of rows),
    .jsonl, or .parquet -- same columns as the HF dataset."""
    path = str(path)
    if path.endswith(".parquet"):
        import pandas as pd
        return pd.read_parquet(path).to_dict("records")
    with open(path) as f:
        if path.endswith(".jsonl"):
            return [json.loads(line) for line in f if line.strip()]
        return json.load(f)
