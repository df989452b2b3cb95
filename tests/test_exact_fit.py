"""The kernels' exact-fit search (one sweep over the cells, sparc_core.cuh `exact_fit`) against
(a) the reference's own answers (golden polyfit vectors), (b) the oracle's restatement of the
reference's search and (c) the reference-order search compiled from the same header -- on
stamped regions with stacked / off-region ylops, duplicates, zero-area and oversized shapes."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import pytest

import sparc_oracle as so
from emu.emu_env import build_emu
from fuzzrows import SMALL_SHAPES
from sparc_gym_b200.store import SHAPE_DTYPE, shape_slot

LIB = C.CDLL(build_emu())
LIB.emu_exact_fit.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                              C.POINTER(C.c_longlong)]


LIB.emu_exact_fit_split.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.POINTER(C.c_longlong)]


def device_fit(which, X, Y, region, polys, ylops, split=None):
    """polys / ylops: lists of shape arrays.  Returns (bool, search iterations); with split=(code, nsplit) the answer
    and effort of that one copy of the split search."""
    names = {}
    for a in polys + ylops:
        names.setdefault(json.dumps(a), len(names))
    slots = np.zeros(max(len(names), 1), SHAPE_DTYPE)
    for k, i in names.items():
        slots[i] = shape_slot(json.loads(k), (X - 1) // 2, (Y - 1) // 2)
    npoly = np.zeros(32, np.uint8)
    nylop = np.zeros(32, np.uint8)
    for a in polys:
        npoly[names[json.dumps(a)]] += 1
    for a in ylops:
        nylop[names[json.dumps(a)]] += 1
    reg = 0
    for (x, y) in region:
        reg |= 1 << ((x // 2) * 8 + (y // 2))
    it = C.c_longlong()
    if split is not None:
        ok = LIB.emu_exact_fit_split(C.c_uint64(reg), (X - 1) // 2, (Y - 1) // 2, slots.ctypes.data, len(names),
                                     npoly.ctypes.data, nylop.ctypes.data, split[0], split[1], C.byref(it))
        return bool(ok), it.value
    ok = LIB.emu_exact_fit(which, C.c_uint64(reg), (X - 1) // 2, (Y - 1) // 2, slots.ctypes.data, len(names),
                           npoly.ctypes.data, nylop.ctypes.data, C.byref(it))
    return bool(ok), it.value


def area(a):
    return int(np.array(a).sum())


def test_golden_polyfit_cases():
    with gzip.open(os.path.join(os.path.dirname(__file__), "golden", "polyfit.json.gz")) as f:
        cases = json.loads(f.read())
    n = 0
    for i, c in enumerate(cases):
        if not c["area_matched"]:
            continue        # the rule never calls the search then (SG.py:696-699)
        polys = [c["shapes"][x["name"]] for x in c["instances"] if x["kind"] == "poly"]
        ylops = [c["shapes"][x["name"]] for x in c["instances"] if x["kind"] == "ylop"]
        for which in (0, 1):
            got, _ = device_fit(which, c["X"], c["Y"], c["region"], polys, ylops)
            assert got == c["expected"], (i, which)
        n += 1
    assert n > 100


def random_case(rng, max_ylops=6):
    w, h = int(rng.integers(1, 6)), int(rng.integers(1, 6))
    need = np.zeros((w, h), int)
    polys, ylops = [], []
    shapes = SMALL_SHAPES + [[[1, 1, 1, 1]], [[1, 0], [1, 0], [1, 1]], [[1] * 8]]
    for _ in range(int(rng.integers(1, 6))):
        a = np.array(shapes[int(rng.integers(len(shapes)))])
        if a.shape[0] > w or a.shape[1] > h:
            continue
        ox, oy = int(rng.integers(w - a.shape[0] + 1)), int(rng.integers(h - a.shape[1] + 1))
        need[ox:ox + a.shape[0], oy:oy + a.shape[1]] += a
        polys.append(a.tolist())
    for _ in range(int(rng.integers(0, 4))):
        a = np.array(shapes[int(rng.integers(5))])
        if a.shape[0] > w or a.shape[1] > h:
            continue
        ox, oy = int(rng.integers(w - a.shape[0] + 1)), int(rng.integers(h - a.shape[1] + 1))
        sub = need[ox:ox + a.shape[0], oy:oy + a.shape[1]]
        if (sub >= a).all() and rng.random() < 0.8:
            sub -= a
            ylops.append(a.tolist())
    if need.max(initial=0) > 1:
        # the region is a set: extra layers must come from ylops, so dig the surplus out with 1x1 ylops
        for i, j in zip(*np.nonzero(need > 1)):
            for _ in range(need[i, j] - 1):
                ylops.append([[1]])
            need[i, j] = 1
    region = [(2 * i + 1, 2 * j + 1) for i in range(w) for j in range(h) if need[i, j] > 0]
    r = rng.random()
    if r < 0.3 and polys:                         # area-preserving perturbations: mostly unsolvable
        k = int(rng.integers(3))
        if k == 0:                                # move one region cell somewhere else
            free = [(2 * i + 1, 2 * j + 1) for i in range(w) for j in range(h) if (2 * i + 1, 2 * j + 1) not in region]
            if free and region:
                region[int(rng.integers(len(region)))] = free[int(rng.integers(len(free)))]
                region.sort()
        elif k == 1:                              # swap a poly for another shape of the same area
            i = int(rng.integers(len(polys)))
            same = [q for q in shapes if area(q) == area(polys[i]) and q != polys[i]]
            if same:
                polys[i] = same[int(rng.integers(len(same)))]
        else:                                     # swap a ylop likewise / add a cancelling pair of different shapes
            two = [q for q in shapes if area(q) == 2]
            polys.append(two[0]); ylops.append(two[1])
    elif r < 0.4 and polys:                       # area-breaking (filtered out unless it happens to match)
        k = int(rng.integers(3))
        if k == 0 and len(region) > 1:
            region.pop(int(rng.integers(len(region))))
        elif k == 1:
            polys[int(rng.integers(len(polys)))] = shapes[int(rng.integers(len(shapes)))]
        else:
            polys.append(shapes[int(rng.integers(3))])
    return 2 * w + 1, 2 * h + 1, region, polys, ylops[:max_ylops]


def test_sweep_equals_reference_order_and_oracle():
    rng = np.random.default_rng(11)
    n = n_fit = n_y = 0
    worst = 0
    while n < 1500:
        X, Y, region, polys, ylops = random_case(rng)
        if not region or not polys:
            continue
        if sum(map(area, polys)) - sum(map(area, ylops)) != len(region):
            continue       # only area-matched cases reach the search
        a, it_sweep = device_fit(0, X, Y, region, polys, ylops)
        b, it_ref = device_fit(1, X, Y, region, polys, ylops)
        assert a == b, (X, Y, region, polys, ylops)
        if n % 6 == 0 and len(ylops) <= 2:      # the Python search is slow: sample it
            want = so.exact_fit(X, Y, region, [(json.dumps(p), p) for p in polys], [(json.dumps(p), p) for p in ylops])
            assert a == want, (X, Y, region, polys, ylops)
        n += 1
        n_fit += a
        n_y += bool(ylops)
        worst = max(worst, it_sweep)
    assert n_fit > 300 and n - n_fit > 100 and n_y > 300, (n_fit, n_y, worst)


def test_dense_boards_sweep_equals_reference_order():
    """6x6 board, ~11 random small pieces of total area 36 (dataset.gen_dense_poly_puzzle's recipe): the sweep search
    of the kernels and the reference-order search compiled from the same header agree, and the sweep stays within
    a few 1e4 nodes on the hardest draws (the reference-order search is capped here by skipping draws it would
    need > 2e6 nodes for)."""
    from sparc_gym_b200.dataset import SMALL_PIECES
    rng = np.random.default_rng(0)
    region = [(2 * i + 1, 2 * j + 1) for i in range(6) for j in range(6)]
    n = n_fit = worst = 0
    for _ in range(60):
        polys, tot = [], 0
        while tot < 36:
            q = SMALL_PIECES[int(rng.integers(len(SMALL_PIECES)))]
            if tot + area(q) <= 36:
                polys.append(q)
                tot += area(q)
            elif 36 - tot == 1:
                break
        if tot != 36:
            continue
        a, it = device_fit(0, 13, 13, region, polys, [])
        worst = max(worst, it)
        if it < 4000:                     # the reference-order search has no floor pruning: keep it to the easy draws
            b, _ = device_fit(1, 13, 13, region, polys, [])
            assert a == b, polys
            n += 1
            n_fit += a
    assert n > 20 and 0 < n_fit and worst < 200000, (n, n_fit, worst)


def split_fit(X, Y, region, polys, ylops, nsplit=5):
    """The 2^nsplit copies of the split search as k_val's lanes run them: (OR of the answers, node counts per copy).
    A real warp stops all copies as soon as one finds a fit; here every copy runs to its own end."""
    res = [device_fit(0, X, Y, region, polys, ylops, split=(code, nsplit)) for code in range(1 << nsplit)]
    return any(r[0] for r in res), [r[1] for r in res]


def test_split_search_partitions_the_tree():
    """exact_fit(code, nsplit): the copies' answers OR to the plain search's answer (random cases with ylops, perturbed
    unsolvable cases), and on unsolvable cases -- where every copy must exhaust its share -- no copy visits more nodes
    than the plain search does, and together they visit it about once (shared trunk above the splits aside)."""
    rng = np.random.default_rng(5)
    n = n_fit = 0
    while n < 400:
        X, Y, region, polys, ylops = random_case(rng)
        if not region or not polys or sum(map(area, polys)) - sum(map(area, ylops)) != len(region):
            continue
        want, it_plain = device_fit(0, X, Y, region, polys, ylops)
        got, its = split_fit(X, Y, region, polys, ylops)
        assert got == want, (X, Y, region, polys, ylops)
        if not want:
            assert max(its) <= it_plain, (its, it_plain)
        n += 1
        n_fit += want
    assert n_fit > 80 and n - n_fit > 40


def test_split_search_on_dense_boards():
    """Dense 6x6 boards (the hard tail of DESIGN.md): answers agree and the slowest copy of an unsolvable or hard board
    does a small fraction of the plain search's work."""
    from sparc_gym_b200.dataset import SMALL_PIECES
    rng = np.random.default_rng(1)
    region = [(2 * i + 1, 2 * j + 1) for i in range(6) for j in range(6)]
    ratios = []
    for _ in range(25):
        polys, tot = [], 0
        while tot < 36:
            q = SMALL_PIECES[int(rng.integers(len(SMALL_PIECES)))]
            if tot + area(q) <= 36:
                polys.append(q)
                tot += area(q)
            elif 36 - tot == 1:
                break
        if tot != 36:
            continue
        want, it_plain = device_fit(0, 13, 13, region, polys, [])
        res = [device_fit(0, 13, 13, region, polys, [], split=(code, 5)) for code in range(32)]
        assert any(r[0] for r in res) == want, polys
        if it_plain > 2000:
            # a warp ends when the first copy finds a fit (solvable) or when the slowest copy gives up (unsolvable)
            warp_nodes = min(r[1] for r in res if r[0]) if want else max(r[1] for r in res)
            ratios.append(warp_nodes / it_plain)
    assert ratios and float(np.median(ratios)) < 0.5, ratios


def brute_fit(X, Y, region, polys, ylops):
    """Independent exhaustive answer for small boards: every multiset of ylop anchors (instances of one shape in
    non-decreasing anchor order), then an exact cover of the resulting demand by the polys (translation only, anchored at
    the shape's first cell in (x, y) order -- SG.py:856-871)."""
    w, h = (X - 1) // 2, (Y - 1) // 2

    def cells_at(a, cx, cy):
        a = np.array(a)
        pts = [(i, j) for i in range(a.shape[0]) for j in range(a.shape[1]) if a[i, j]]
        if not pts:
            return None
        ai, aj = pts[0]
        out = [(cx + i - ai, cy + j - aj) for i, j in pts]
        return out if all(0 <= x < w and 0 <= y < h for x, y in out) else None

    need0 = np.zeros((w, h), int)
    for (x, y) in region:
        need0[x // 2, y // 2] = 1
    polys = [p for p in polys if np.array(p).sum() > 0]
    ylops = sorted((json.dumps(y) for y in ylops if np.array(y).sum() > 0))

    def cover(need, left):
        if (need < 0).any():
            return False
        if not left:
            return not need.any()
        nz = np.argwhere(need > 0)
        if len(nz) == 0:
            return False
        cx, cy = map(int, nz[0])
        tried = set()
        for k, p in enumerate(left):
            key = json.dumps(p)
            if key in tried:
                continue
            tried.add(key)
            c = cells_at(p, cx, cy)
            if c is None:
                continue
            n2 = need.copy()
            for x, y in c:
                n2[x, y] -= 1
            if cover(n2, left[:k] + left[k + 1:]):
                return True
        return False

    def place(k, need, last):
        if k == len(ylops):
            return cover(need, polys)
        shape = json.loads(ylops[k])
        start = last if k and ylops[k] == ylops[k - 1] else 0
        for a in range(start, w * h):
            c = cells_at(shape, a // h, a % h)
            if c is None:
                continue
            n2 = need.copy()
            for x, y in c:
                n2[x, y] += 1
            if place(k + 1, n2, a):
                return True
        return False

    return place(0, need0, 0)


def test_wide_counters_agree_with_the_fast_search_and_with_brute_force():
    """Puzzles with more than six ylop instances run the same sweep on 6-bit demand counters (header flag bit 2): on cases
    the 4-bit search can do it gives the same answers node for node, and with 7-12 ylops it agrees with an independent
    exhaustive search on small boards."""
    rng = np.random.default_rng(21)
    n = 0
    while n < 300:
        X, Y, region, polys, ylops = random_case(rng)
        if not region or not polys or sum(map(area, polys)) - sum(map(area, ylops)) != len(region):
            continue
        a, ia = device_fit(0, X, Y, region, polys, ylops)
        b, ib = device_fit(2, X, Y, region, polys, ylops)
        assert a == b and ia == ib, (X, Y, region, polys, ylops)
        n += 1
    n = n_fit = n_many = 0
    while n < 60:
        X, Y, region, polys, ylops = random_case(rng, max_ylops=12)
        w, h = (X - 1) // 2, (Y - 1) // 2
        if w * h > 9 or len(ylops) < 7 or not region or not polys:
            continue
        # rebalance with monominoes so that the areas match and the search actually runs
        gap = sum(map(area, polys)) - sum(map(area, ylops)) - len(region)
        if gap > 0:
            ylops = ylops + [[[1]]] * gap
        elif gap < 0:
            polys = polys + [[[1]]] * (-gap)
        if len(ylops) > 12:
            continue
        got, _ = device_fit(2, X, Y, region, polys, ylops)
        assert got == brute_fit(X, Y, region, polys, ylops), (X, Y, region, polys, ylops)
        n += 1
        n_fit += got
        n_many += len(ylops) > 6
    assert n_many == 60 and 5 < n_fit < 60
