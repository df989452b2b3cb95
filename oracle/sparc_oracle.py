"""CPU ORACLE -- test infrastructure, not product code.

A plain Python/numpy restatement of the reference SPaRC environment's hot path
(reference: SPaRC_Gym/SPaRC_Gym.py, cited per function as SG.py:<lines>).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this module; the product (`sparc_gym_b200`) never does and fails loudly without its
CUDA library.

Parity status: PINNED.  The reference ships no tests or golden vectors,
so this restatement is pinned against the *unmodified reference executed in the build
container* (tests/golden/make_golden.py -> tests/golden/*.json.gz, replayed by
tests/test_oracle_golden.py on every run, and differential-fuzzed live against the
reference whenever /root/reference is present, tests/test_oracle_vs_reference.py).

Deliberate difference: `clean_reset=True` (default) zeroes the `visited` and
`agent_location` planes when a puzzle is (re)loaded; the reference aliases its cached planes
and leaks them across episodes (SG.py:149-151).  `clean_reset=False` reproduces the leak.
"""
from __future__ import annotations

import json
from collections import deque

import numpy as np
import yaml

try:  # libyaml loader: same documents, ~7x faster than the pure-Python loader
    _YAML_LOADER = yaml.CSafeLoader
except AttributeError:  # pragma: no cover
    _YAML_LOADER = yaml.SafeLoader

COLOR_TO_NUMBER = {"red": 1, "blue": 2, "yellow": 3, "green": 4, "black": 5, "purple": 6,
                   "orange": 7, "white": 8}                      # SG.py:310
DIRECTIONS = ((1, 0), (0, -1), (-1, 0), (0, 1))                  # SG.py:212-217 (right, up, left, down)
FIXED_PLANES = ("visited", "gaps", "agent_location", "target_location")  # SG.py:272-277
RULE_NAMES = ("reached_target", "path_not_crossing", "no_gap_violations", "all_dots_collected",
              "square_color_separation", "star_pairing_exact", "triangles_edge_count",
              "poly_ylop_area", "all_rules_satisfied")           # SG.py:912-935

_UNSET = object()


def _load_yaml(text):
    return yaml.load(text, Loader=_YAML_LOADER)


def process_rows(rows, observation="new"):
    """Dataset rows -> list of puzzle dicts (SG.py:219-368).

    Reproduces the ingest quirks: planes are created in first-appearance order after the four
    fixed ones (Q13); the loop variable `symbol` is never reset, so keys other than
    `type`/`dot` reuse the previous value -- across cells and across puzzles (Q11, Q12).
    """
    puzzles = []
    symbol = _UNSET
    for row in rows:
        X = 2 * int(row["grid_size"]["width"]) + 1            # SG.py:243-247
        Y = 2 * int(row["grid_size"]["height"]) + 1
        sols = [[[pt["x"], pt["y"]] for pt in item["path"]] for item in row["solutions"]]  # 252-256
        shapes = _load_yaml(row["polyshapes"])                  # SG.py:260-261
        tv = _load_yaml(row["text_visualization"])["puzzle"]    # SG.py:265-268
        planes = {k: np.zeros((X, Y), dtype=np.int32) for k in FIXED_PLANES}
        color = np.zeros((X, Y), dtype=np.int32)
        info = np.zeros((X, Y), dtype=np.int64)
        cells = tv["cells"] or []
        for cell in cells:                                      # first pass, SG.py:283-325
            props = cell.get("properties", {})
            pos = cell.get("position", {})
            count = shape = col = None
            for key, value in props.items():
                if key == "type":
                    symbol = f"{value}"
                    col = props.get("color", "")
                    if value == "triangle":
                        count = props.get("count", "")
                    elif value not in ("star", "square"):
                        shape = props.get("polyshape", "")
                elif key == "dot":
                    symbol = "dot"
                if symbol is _UNSET:
                    raise UnboundLocalError("cannot access local variable 'symbol'")
                if symbol not in planes:                        # SG.py:305-306
                    planes[symbol] = np.zeros((X, Y), dtype=np.int32)
                if col and col in COLOR_TO_NUMBER:              # SG.py:309-315
                    color[pos.get("x")][pos.get("y")] = COLOR_TO_NUMBER[col]
                if count:                                       # SG.py:318-325
                    info[pos.get("x")][pos.get("y")] = count
                elif shape:
                    info[pos.get("x")][pos.get("y")] = shape
        for cell in cells:                                      # second pass, SG.py:329-343
            props = cell.get("properties", {})
            pos = cell.get("position", {})
            x, y = pos.get("x"), pos.get("y")
            for key, value in props.items():
                if key == "type":
                    symbol = f"{value}"
                elif key == "dot":
                    symbol = "dot"
                elif key == "gap":
                    symbol = "gaps"
                if symbol is _UNSET:
                    raise UnboundLocalError("cannot access local variable 'symbol'")
                if symbol in planes:
                    planes[symbol][x, y] = 1
        planes["gaps"][1:X - 1:2, 1:Y - 1:2] = 1               # SG.py:345-351 (all rule cells)
        pz = {
            "difficulty": row["difficulty_level"], "x_size": X, "y_size": Y,
            "solution_count": row["solution_count"], "solution_paths": sols,
            "polyshapes": shapes,
            "start_location": (tv["start"]["x"], tv["start"]["y"]),
            "target_location": (tv["end"]["x"], tv["end"]["y"]),
            "obs_array": planes, "color_array": color, "additional_info": info,
            "id": row["id"],
        }
        if observation == "SPaRC":
            pz["observ"] = row["puzzle_array"]                 # SG.py:358-360
        puzzles.append(pz)
    return puzzles


# --------------------------------------------------------------------------------------
# regions + rules (SG.py:374-950)
# --------------------------------------------------------------------------------------

def compute_regions(planes, path):
    """SG.py:422-454.  BFS over the whole node lattice: a node is entered if it is a rule
    cell, or if it is neither a gap nor on the path (Q6).  Only rule cells are labelled; ids
    follow raster order (x outer, y inner).  Returns ([cells per region], region_map)."""
    gaps = planes["gaps"]
    X, Y = gaps.shape
    blocked = gaps == 1
    for (px, py) in path:
        blocked[px, py] = True
    region_map = np.full((X, Y), -1, dtype=np.int32)
    regions = []
    for x in range(1, X, 2):
        for y in range(1, Y, 2):
            if region_map[x, y] != -1:
                continue
            rid = len(regions)
            region_map[x, y] = rid
            seen_other = np.zeros((X, Y), dtype=bool)
            cells = []
            q = deque([(x, y)])
            while q:
                cx, cy = q.popleft()
                if cx & 1 and cy & 1:
                    cells.append((cx, cy))
                for dx, dy in ((0, 1), (0, -1), (1, 0), (-1, 0)):
                    nx, ny = cx + dx, cy + dy
                    if not (0 <= nx < X and 0 <= ny < Y):
                        continue
                    if nx & 1 and ny & 1:
                        if region_map[nx, ny] == -1:
                            region_map[nx, ny] = rid
                            q.append((nx, ny))
                    # note: a rule cell is always a gap (SG.py:345-351) hence always blocked
                    if not blocked[nx, ny] and not seen_other[nx, ny]:
                        seen_other[nx, ny] = True
                        q.append((nx, ny))
            regions.append(cells)
    return regions, region_map


def shape_offsets(arr):
    """SG.py:856-871: node-lattice offsets of a shape relative to its anchor = the
    lexicographically first filled entry; array index 0 maps to grid x (Q9)."""
    a = np.array(arr, dtype=np.int32)
    if a.ndim != 2:
        a = a.reshape(a.shape[0], -1) if a.ndim > 2 else np.atleast_2d(a)
    xs, ys = np.where(a == 1)
    if len(xs) == 0:
        return []
    ax = xs.min()
    ay = ys[xs == ax].min()
    return [(2 * int(x - ax), 2 * int(y - ay)) for x, y in zip(xs, ys)]


def exact_fit(X, Y, region_cells, polys, ylops):
    """SG.py:738-895.  `polys` / `ylops` are lists of (name, shape array).  Existence of a
    placement: every ylop anywhere on the rule-cell lattice (all its cells in the grid; adds
    one more layer to be covered), then the polys, translated only, cover each cell exactly
    as many times as it is demanded."""
    area = lambda arr: int(np.array(arr, dtype=np.int32).sum())
    net = sum(area(a) for _, a in polys) - sum(area(a) for _, a in ylops)
    if net == 0 and sorted(n for n, _ in polys) == sorted(n for n, _ in ylops):
        return True                                             # SG.py:766-776
    polys = [(n, shape_offsets(a)) for n, a in polys]
    ylops = [(n, shape_offsets(a)) for n, a in ylops]
    need = np.zeros((X, Y), dtype=np.int32)
    if net > 0:                                                 # SG.py:779-780
        for (cx, cy) in region_cells:
            need[cx, cy] = 1
    anchors = [(x, y) for x in range(1, X, 2) for y in range(1, Y, 2)]

    def in_grid(offs, ax, ay):
        return all(0 <= ax + dx < X and 0 <= ay + dy < Y for dx, dy in offs)

    def place_polys(rest):
        if (need < 0).any():                                    # SG.py:822 (grid > 0)
            return False
        if not rest:
            return not (need > 0).any()                         # SG.py:825-826
        open_cells = np.argwhere(need > 0)
        if open_cells.size == 0:                                # SG.py:829-830
            return True
        tx, ty = (int(v) for v in open_cells[0])                # argwhere is already (x, y)-sorted
        tried = set()
        for i, (name, offs) in enumerate(rest):
            if name in tried:
                continue
            tried.add(name)
            if not in_grid(offs, tx, ty):
                continue
            for dx, dy in offs:
                need[tx + dx, ty + dy] -= 1
            if place_polys(rest[:i] + rest[i + 1:]):
                return True
            for dx, dy in offs:
                need[tx + dx, ty + dy] += 1
        return False

    def place_ylops(k):
        if k == len(ylops):
            return place_polys(polys)
        offs = ylops[k][1]
        for (ax, ay) in anchors:
            if not in_grid(offs, ax, ay):
                continue
            for dx, dy in offs:
                need[ax + dx, ay + dy] += 1
            if place_ylops(k + 1):
                return True
            for dx, dy in offs:
                need[ax + dx, ay + dy] -= 1
        return False

    return place_ylops(0)


def validate(pz, path, agent, target, *, reference_cost=False):
    """SG.py:901-950: the nine rule flags + region summaries for the current state."""
    planes, color, info = pz["obs_array"], pz["color_array"], pz["additional_info"]
    X, Y = pz["x_size"], pz["y_size"]
    regions, region_map = compute_regions(planes, path)
    # symbols per region (SG.py:456-482)
    sym = [dict() for _ in regions]
    colors = [dict() for _ in regions]
    for layer, arr in planes.items():
        if layer in FIXED_PLANES:
            continue
        for x, y in zip(*np.where(arr == 1)):
            rid = region_map[x, y]
            if rid == -1:
                continue
            sym[rid].setdefault(layer, []).append((int(x), int(y)))
            c = int(color[x, y])
            if c:
                colors[rid][c] = colors[rid].get(c, 0) + 1
    flags, det = {}, {}
    flags["reached_target"] = bool(agent[0] == target[0] and agent[1] == target[1])    # SG.py:488
    det["reached_target"] = {"agent_loc": [int(v) for v in agent], "target_loc": [int(v) for v in target]}
    pt = [tuple(p) for p in path]
    dup = {}                                                                           # SG.py:498
    for q in pt:
        dup[q] = dup.get(q, 0) + 1
    dup = {k: v for k, v in dup.items() if v > 1}
    flags["path_not_crossing"] = len(dup) == 0
    det["path_not_crossing"] = {"duplicates": dup}
    viol = [(x, y) for x, y in pt if planes["gaps"][x, y] == 1]                        # SG.py:508
    flags["no_gap_violations"] = len(viol) == 0
    det["no_gap_violations"] = {"violations": viol}
    if "dot" in planes:                                                                # SG.py:521
        d = planes["dot"] == 1
        total, got = int(d.sum()), int((d & (planes["visited"] == 1)).sum())
        flags["all_dots_collected"] = got == total
        det["all_dots_collected"] = {"total": total, "collected": got}
    else:
        flags["all_dots_collected"] = True
        det["all_dots_collected"] = {"total": 0, "collected": 0}
    ok = True                                                                          # SG.py:534
    if "square" in planes:
        bad, rows = [], []
        for r in range(len(regions)):
            squares = sym[r].get("square", [])
            if not squares:
                continue
            cs = set(color[x, y] for x, y in squares if color[x, y] != 0)   # numpy scalars, like SG.py:547
            if len(cs) > 1:
                bad.append(r)
            rows.append({"region": r, "square_count": len(squares), "colors": list(cs)})
        ok = not bad
        det["square_color_separation"] = {"violating_regions": bad, "region_square_details": rows}
    else:
        det["square_color_separation"] = {"regions": []}
    flags["square_color_separation"] = ok
    ok = True                                                                          # SG.py:554
    if "star" in planes:
        violations, per_region = [], []
        for r in range(len(regions)):
            stars = sym[r].get("star", [])
            if not stars:
                continue
            total = {}
            for coords in sym[r].values():
                for x, y in coords:
                    c = color[x, y]
                    if c != 0:
                        total[c] = total.get(c, 0) + 1
            star_colors = {}
            for x, y in stars:
                c = color[x, y]
                if c == 0:
                    violations.append({"region": r, "color": 0, "found_total": 1})
                    continue
                star_colors[c] = star_colors.get(c, 0) + 1
            region_ok, rows = True, []
            for c, n_star in star_colors.items():
                tc = total.get(c, 0)
                if tc != 2:
                    region_ok = False
                    violations.append({"region": r, "color": c, "found_total": tc, "star_cells": n_star})
                rows.append({"color": c, "total_symbols_of_color": tc, "star_cells": n_star, "ok": tc == 2})
            per_region.append({"region": r, "details": rows, "all_ok": region_ok})
        ok = not violations
        det["star_pairing_exact"] = {"violations": violations, "per_region": per_region}
    else:
        det["star_pairing_exact"] = {"regions": []}
    flags["star_pairing_exact"] = ok
    mism = []                                                                          # SG.py:617
    if "triangle" in planes:
        tri = planes["triangle"]
        pset = set(pt)
        for x in range(1, X - 1):
            for y in range(1, Y - 1):
                if tri[x, y] == 1 and int(info[x, y]) > 0:
                    t = sum(((x + dx, y + dy) in pset) for dx, dy in DIRECTIONS)
                    if t != int(info[x, y]):
                        mism.append({"x": x, "y": y, "required": int(info[x, y]), "touches": t})
    flags["triangles_edge_count"] = not mism
    det["triangles_edge_count"] = {"mismatches": mism}
    # poly / ylop (SG.py:648-736)
    ok = True
    shapes = pz["polyshapes"]
    det["poly_ylop_area"] = {"regions": []}
    if isinstance(shapes, dict):
        by_region, any_inst = {}, False
        for x in range(X):
            for y in range(Y):
                v = info[x, y]
                if v != 0 and f"{v}" in shapes:
                    kind = "poly" if planes["poly"][x, y] == 1 else "ylop"  # KeyError as in SG.py:734
                    any_inst = True
                    rid = region_map[x, y]
                    if rid != -1:
                        by_region.setdefault(int(rid), []).append((f"{v}", kind))
        if any_inst and reference_cost:
            compute_regions(planes, path)   # the reference recomputes the regions here (SG.py:660)
        rows = []
        for rid, insts in by_region.items():
            area = {n: int(np.array(shapes[n]).sum()) for n, _ in insts}
            pa = sum(area[n] for n, k in insts if k == "poly")
            ya = sum(area[n] for n, k in insts if k == "ylop")
            net, ra = pa - ya, len(regions[rid])
            row = {"region": rid, "area_check": {"region_area": ra, "poly_area": pa, "ylop_area": ya,
                                                 "net": net, "ok": ra == net}}
            if net != ra:
                fit = False
                row["exact_fit"] = {"ok": False, "skipped": True}
            else:
                polys = [(n, shapes[n]) for n, k in insts if k == "poly"]
                ylops = [(n, shapes[n]) for n, k in insts if k == "ylop"]
                fit = exact_fit(X, Y, regions[rid], polys, ylops)
                row["exact_fit"] = {"ok": fit, "region_id": rid, "region_area": ra, "poly_area": pa,
                                    "ylop_area": ya, "net": net}
            row["ok"] = fit
            ok &= fit
            rows.append(row)
        if any_inst:
            det["poly_ylop_area"] = {"violations": [r["region"] for r in rows if not r["ok"]],
                                     "region_details": rows}
    flags["poly_ylop_area"] = ok
    flags["all_rules_satisfied"] = all(flags[k] for k in RULE_NAMES[:8])               # SG.py:932-935
    det["all_rules_satisfied"] = {"rules_checked": list(RULE_NAMES[:8])}
    summaries = {r: {"id": r, "area": len(regions[r]),
                     "symbol_counts": {k: len(v) for k, v in sym[r].items()},
                     "colors": colors[r]} for r in range(len(regions))}
    return flags, summaries, region_map, det


# --------------------------------------------------------------------------------------
# the environment (SG.py:44-217, 956-1265)
# --------------------------------------------------------------------------------------

class OracleEnv:
    """Single-env restatement with the reference's call surface:
    `reset(seed=None, options=None) -> (obs, info)`, `step(a) -> (obs, r, term, trunc, info)`.

    `info["rule_status"]` holds `{name: {"passed": bool}}` for the nine rules plus
    `_terminated` / `_truncated` / `_regions`, i.e. the reference's structure without the
    per-rule `detail` payloads.  `reference_cost=True` additionally repeats the work the
    reference repeats (two validations per step, SG.py:1227 + 1011; three legality scans;
    a second region pass inside the poly rule, SG.py:660) so that timing this class is a fair
    stand-in for timing the reference.
    """

    def __init__(self, rows, observation="new", traceback=False, max_steps=2000, *,
                 clean_reset=True, reference_cost=False):
        if observation not in ("new", "SPaRC"):
            raise ValueError("Invalid observation type. Choose 'new' or 'SPaRC'.")
        self.observation = observation
        self.traceback = traceback
        self.max_steps = max_steps
        self.clean_reset = clean_reset
        self.reference_cost = reference_cost
        self.puzzles = process_rows(rows, observation)
        self.current_puzzle_index = 0
        self.current_step = 0
        self.np_random = None
        self._load(0)

    # -- episode init (SG.py:95-217)
    def _load(self, index):
        pz = self.puzzles[index]
        self.pz = pz
        if self.clean_reset:
            pz["obs_array"]["visited"][...] = 0
            pz["obs_array"]["agent_location"][...] = 0
        self.obs_array = pz["obs_array"]
        self.color_array = pz["color_array"]
        self.additional_info = pz["additional_info"]
        self.x_size, self.y_size = pz["x_size"], pz["y_size"]
        self.solution_paths = pz["solution_paths"]
        self.solution_count = pz["solution_count"]
        if self.observation == "SPaRC":
            self.observ = [[str(c) for c in r] for r in pz["observ"]]   # SG.py:153-164
        sx, sy = pz["start_location"]
        self.path = [[sx, sy]]
        self.normal_reward = 0
        self.outcome_reward = 0
        self.agent = np.array([sx, sy], dtype=np.int32)
        self.target = np.array(pz["target_location"], dtype=np.int32)
        if self.reference_cost:
            validate(pz, self.path, self.agent, self.target, reference_cost=True)  # SG.py:182
        self.obs_array["visited"][sx, sy] = 1
        self.obs_array["agent_location"][sx, sy] = 1
        self.obs_array["target_location"][self.target[0], self.target[1]] = 1

    # -- SG.py:1024-1051
    def legal_actions(self):
        out = []
        gaps, visited = self.obs_array["gaps"], self.obs_array["visited"]
        for a, (dx, dy) in enumerate(DIRECTIONS):
            nx, ny = int(self.agent[0]) + dx, int(self.agent[1]) + dy
            cx = min(max(nx, 0), self.x_size - 1)
            cy = min(max(ny, 0), self.y_size - 1)
            if gaps[cx, cy] != 0:
                continue
            inside = (nx, ny) == (cx, cy)
            if visited[cx, cy] == 1:
                if (self.traceback and len(self.path) >= 2 and inside
                        and self.path[-2][0] == cx and self.path[-2][1] == cy):
                    out.append(a)
            elif inside:
                out.append(a)
        return out

    def _obs(self):
        if self.observation == "new":                                       # SG.py:979
            return {"base": self.obs_array, "color": self.color_array,
                    "additional_info": self.additional_info}
        return json.dumps(self.observ, separators=(",", ":"))               # SG.py:992

    def _info(self):
        flags, regions, _, det = validate(self.pz, self.path, self.agent, self.target,
                                          reference_cost=self.reference_cost)    # SG.py:1011
        status = {k: {"passed": bool(v), "detail": det[k]} for k, v in flags.items()}
        status["_terminated"] = {"passed": True, "detail": False}           # Q16
        status["_truncated"] = {"passed": True, "detail": False}
        status["_regions"] = regions
        return {"solution_count": self.solution_count, "difficulty": self.pz["difficulty"],
                "grid_x_size": self.x_size, "grid_y_size": self.y_size,
                "legal_actions": self.legal_actions(), "current_step": self.current_step,
                "agent_location": self.agent, "rule_status": status,
                "Rewards": {"normal_reward": self.normal_reward,
                            "outcome_reward": self.outcome_reward}}

    # -- SG.py:1057-1108
    def reset(self, seed=None, options=None):
        if seed is not None or self.np_random is None:
            self.np_random = np.random.default_rng(seed)
        if options is not None:
            pid = options.get("puzzle_id", None)
            for i, pz in enumerate(self.puzzles):
                if pz["id"] == pid:
                    self.current_puzzle_index = i
                    break
        elif seed is not None:
            self.current_puzzle_index = int(self.np_random.integers(len(self.puzzles)))
        else:
            self.current_puzzle_index = (self.current_puzzle_index + 1) % len(self.puzzles)
        self.current_step = 0
        self._load(self.current_puzzle_index)
        return self._obs(), self._info()

    # -- SG.py:1111-1238
    def step(self, action):
        orig = (int(self.agent[0]), int(self.agent[1]))
        self.current_step += 1
        self.normal_reward = 0
        truncated = self.current_step >= self.max_steps
        if action in self.legal_actions():
            dx, dy = DIRECTIONS[action]
            nx, ny = orig[0] + dx, orig[1] + dy
            if self.obs_array["visited"][nx, ny] == 1:        # traceback undo (legality
                self.obs_array["agent_location"][orig] = 0    # guarantees n == path[-2])
                self.obs_array["visited"][orig] = 0
                if self.observation == "SPaRC":               # SG.py:1150-1152
                    self.observ[orig[1]][orig[0]] = "." if self.obs_array["gaps"][orig] == 1 else "+"
                del self.path[-1]
            else:
                self.obs_array["agent_location"][orig] = 0
                if self.observation == "SPaRC":
                    self.observ[orig[1]][orig[0]] = "V"
                self.path.append([nx, ny])
            self.agent = np.array([nx, ny], dtype=np.int32)
            self.obs_array["visited"][nx, ny] = 1
            self.obs_array["agent_location"][nx, ny] = 1
            if self.observation == "SPaRC":
                self.observ[ny][nx] = "L"
        terminated = bool(self.agent[0] == self.target[0] and self.agent[1] == self.target[1])
        legal_after = self.legal_actions()
        if not legal_after:
            truncated = True
        if terminated:
            truncated = False
        if terminated or truncated:                            # SG.py:1204-1213
            for i in range(self.solution_count):
                if np.array_equal(self.path, self.solution_paths[i]):
                    self.outcome_reward = 1
                    self.normal_reward = 1
                    break
            if self.outcome_reward != 1:
                self.outcome_reward = -1
                self.normal_reward = -1
        else:                                                  # SG.py:1214-1223
            self.outcome_reward = 0
            if (int(self.agent[0]), int(self.agent[1])) != orig:
                n = len(self.path)
                for i in range(self.solution_count):
                    s = self.solution_paths[i]
                    if n <= len(s) and all(self.path[k][0] == s[k][0] and self.path[k][1] == s[k][1]
                                           for k in range(n)):
                        self.normal_reward = 0.01
                        break
                    self.normal_reward = -0.01
        if self.reference_cost:                                # SG.py:1227 (result overwritten, Q16)
            validate(self.pz, self.path, self.agent, self.target, reference_cost=True)
            self.legal_actions()                               # third legality scan is in _info
        return self._obs(), self.normal_reward, terminated, truncated, self._info()
