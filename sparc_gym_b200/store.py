"""Pack ingested puzzles into the struct-of-arrays puzzle store the CUDA kernels read
(`SparcStoreDesc`, include/sparc_b200.h).

Everything the kernels need is precomputed here, once, on the host (the reference does its
equivalent work in its constructor, SG.py:77-90):

* dense bit-packed one-hot planes in *observation output order* (the encoder only has to
  OR in the two dynamic planes and expand bits to int32);
* node bitboards for the dynamics (walkable nodes) and for the rule validator (16-stride);
* rule-cell bitboards (uint64) for the per-region symbol / colour counting;
* polyshape slots (anchor-normalised masks exactly as `_get_offsets`, SG.py:856-871);
* a trie over the first `solution_count` listed solutions so that the "is the walked path a
  prefix of / equal to a listed solution" reward test (SG.py:1204-1223) is O(1) per step.

Bit conventions: dense board bit = x*slab_y + y; 16-stride board bit = x*16 + y;
cell board bit = cx*8 + cy for the rule cell (2cx+1, 2cy+1).
"""
from __future__ import annotations

import numpy as np

from .ingest import FIXED_KEYS, SYMBOL_KEYS, Puzzle

MAX_DIM = 15
BOARD_WORDS = 8
MAX_SHAPES = 32
MAX_YLOPS = 6            # served by the 4-bit demand counters of the fast exact-fit search
MAX_YLOPS_WIDE = 30      # ... by the 6-bit counters (header flag bit 2)
TRIE_NODE_WORDS = 8
PLANE_IDS = {k: i for i, k in enumerate(FIXED_KEYS + SYMBOL_KEYS)}
DIRS = ((1, 0), (0, -1), (-1, 0), (0, 1))   # action id -> (dx, dy)  (SG.py:212-217)

# node / cell board slots (mirror the enums in include/sparc_b200.h)
NB_WALK, NB_GAPS16, NB_CELL16, NB_VALID16, NB_DOT16, NB_TRI1, NB_TRI2, NB_TRI3, NB_TRI4, NB_TRIBAD = range(10)
N_NODE_BOARDS = 10
CB_CELLS, CB_STAR, CB_SQUARE, CB_SYMB0, CB_SYMB1, CB_COLOR1 = 0, 1, 2, 3, 4, 5
N_CELL_BOARDS = 16

HEADER_DTYPE = np.dtype([("X", "u1"), ("Y", "u1"), ("start_x", "u1"), ("start_y", "u1"),
                         ("target_x", "u1"), ("target_y", "u1"), ("n_shapes", "u1"), ("flags", "u1"),
                         ("shape_off", "<u4"), ("trie_root", "<u4"), ("solution_count", "<i4"),
                         ("difficulty", "<i4"), ("reserved", "<u4", (2,))])
SHAPE_DTYPE = np.dtype([("mask", "<u8"), ("anchors", "<u8"), ("poly_cells", "<u8"), ("ylop_cells", "<u8")])
assert HEADER_DTYPE.itemsize == 32 and SHAPE_DTYPE.itemsize == 32


def _board(bits_xy, stride):
    """bool[X, Y] -> uint32[8] with bit = x*stride + y."""
    out = np.zeros(BOARD_WORDS, np.uint32)
    xs, ys = np.nonzero(bits_xy)
    idx = xs.astype(np.int64) * stride + ys
    np.bitwise_or.at(out, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    return out


def _cell_board(bits_xy):
    """bool[X, Y] -> uint64 over rule cells, bit = cx*8 + cy."""
    v = 0
    xs, ys = np.nonzero(bits_xy[1::2, 1::2])
    for cx, cy in zip(xs.tolist(), ys.tolist()):
        v |= 1 << (cx * 8 + cy)
    return v


def shape_slot(arr, w=7, h=7):
    """Shape record (SparcShape) of a polyshape array on a w x h board of rule cells: the anchor-normalised mask
    exactly as `_get_offsets` (SG.py:856-871) -- index 0 of the array runs along grid x, anchor = first filled
    entry in (x, y) order -- and the cells on which that anchor may sit with everything in bounds (SG.py:880-883).
    mask = cells | anchor_dy << 56 | flags << 59; anchors = cells | area << 56."""
    a = np.array(arr)
    if a.ndim != 2:
        raise ValueError(f"polyshape array must be 2-D, got shape {a.shape}")
    if not np.isin(a, (0, 1)).all():
        raise ValueError("polyshape arrays must be 0/1")
    xs, ys = np.nonzero(a == 1)
    rec = np.zeros((), SHAPE_DTYPE)
    area = int(a.sum())
    if area > 255:
        raise ValueError("polyshape with more than 255 filled entries")
    rec["anchors"] = area << 56
    if len(xs) == 0:
        rec["mask"] = 2 << 59
        return rec
    ax = xs.min()
    ay = ys[xs == ax].min()
    dx, dy = xs - ax, ys - ay
    min_dy = int(dy.min())
    max_dx, span_y = int(dx.max()), int(dy.max()) - min_dy
    if max_dx > 6 or span_y > 6:
        rec["mask"] = 1 << 59          # larger than any supported board: never placeable
        return rec
    m = 0
    for i, j in zip(dx.tolist(), (dy - min_dy).tolist()):
        m |= 1 << (i * 8 + j)
    # bit2: the filled entries are 4-connected (lets the exact-fit search prune isolated cells)
    cells = set(zip(dx.tolist(), (dy - min_dy).tolist()))
    seen, todo = set(), [next(iter(cells))]
    while todo:
        c = todo.pop()
        if c in seen:
            continue
        seen.add(c)
        todo += [n for n in ((c[0] + 1, c[1]), (c[0] - 1, c[1]), (c[0], c[1] + 1), (c[0], c[1] - 1)) if n in cells]
    flags = 4 if seen == cells else 0
    ady = -min_dy
    rec["mask"] = m | (ady << 56) | (flags << 59)
    anchors = 0
    for cx in range(w):
        for cy in range(h):
            if cy - ady >= 0 and cx + max_dx < w and cy - ady + span_y < h:
                anchors |= 1 << (cx * 8 + cy)
    rec["anchors"] = anchors | (area << 56)
    return rec


class _Trie:
    def __init__(self):
        self.nodes = [[0] * TRIE_NODE_WORDS]            # node 0 = null

    def new(self, parent, depth):
        self.nodes.append([0, 0, 0, 0, parent, 0, depth, 0])
        return len(self.nodes) - 1

    def add_puzzle(self, pz: Puzzle):
        """Insert the first `solution_count` listed paths (SG.py:1205-1206, 1217-1218)."""
        if pz.solution_count > len(pz.solution_paths):
            raise ValueError(f"puzzle {pz.id}: solution_count={pz.solution_count} exceeds the "
                             f"{len(pz.solution_paths)} listed solutions (the reference raises IndexError)")
        root = 0
        for sol in pz.solution_paths[:max(pz.solution_count, 0)]:
            if not sol or tuple(sol[0]) != tuple(pz.start):
                continue                                # can never be a prefix of a walked path
            if not root:
                root = self.new(0, 1)
            node, complete = root, True
            for a, b in zip(sol, sol[1:]):
                step = (b[0] - a[0], b[1] - a[1])
                if step not in DIRS:                    # not a unit move: unreachable beyond here
                    complete = False
                    break
                d = DIRS.index(step)
                nxt = self.nodes[node][d]
                if not nxt:
                    nxt = self.new(node, self.nodes[node][6] + 1)
                    self.nodes[node][d] = nxt
                node = nxt
            if complete:
                self.nodes[node][5] = 1
        return root


class PuzzleStore:
    """Host copy of the packed pool + the per-puzzle metadata the Python drop-in needs."""

    def __init__(self, puzzles, slab=None, planes="auto"):
        if not puzzles:
            raise ValueError("empty puzzle pool")
        self.puzzles = list(puzzles)
        n = len(self.puzzles)
        for pz in self.puzzles:
            if pz.X > MAX_DIM or pz.Y > MAX_DIM:
                raise ValueError(f"puzzle {pz.id}: lattice {pz.X}x{pz.Y} exceeds {MAX_DIM}x{MAX_DIM}")
            unknown = [k for k in pz.keys if k not in PLANE_IDS]
            if unknown:
                raise ValueError(f"puzzle {pz.id}: unsupported symbol kinds {unknown}")
        self.slab_x = max(p.X for p in self.puzzles) if slab is None else int(slab[0])
        self.slab_y = max(p.Y for p in self.puzzles) if slab is None else int(slab[1])
        if any(p.X > self.slab_x or p.Y > self.slab_y for p in self.puzzles) or \
                self.slab_x > MAX_DIM or self.slab_y > MAX_DIM:
            raise ValueError("slab too small / too large")
        if planes == "auto":
            present = {k for p in self.puzzles for k in p.keys}
            self.plane_keys = [k for k in FIXED_KEYS + SYMBOL_KEYS if k in present]
        elif planes == "all":
            self.plane_keys = list(FIXED_KEYS + SYMBOL_KEYS)
        else:
            self.plane_keys = list(planes)
            assert self.plane_keys[:4] == list(FIXED_KEYS)
        self.n_planes = len(self.plane_keys)
        SX, SY = self.slab_x, self.slab_y
        S = SX * SY
        self.words_per_puzzle = (self.n_planes * S + 31) // 32

        self.headers = np.zeros(n, HEADER_DTYPE)
        self.node_boards = np.zeros((n, N_NODE_BOARDS, BOARD_WORDS), np.uint32)
        self.cell_boards = np.zeros((n, N_CELL_BOARDS), np.uint64)
        self.static_bits = np.zeros((n, self.words_per_puzzle), np.uint32)
        self.color = np.zeros((n, SX, SY), np.uint8)
        self.info = np.zeros((n, SX, SY), np.int32)
        self.color4 = np.zeros((n, (4 * S + 31) // 32), np.uint32)   # 4-bit colour ids, dense slab order
        self.info_nz = np.zeros((n, (S + 31) // 32), np.uint32)      # additional_info != 0
        shapes = []
        trie = _Trie()
        self.id_to_index = {}

        for i, pz in enumerate(self.puzzles):
            self.id_to_index.setdefault(pz.id, i)         # first match wins (SG.py:1079-1082)
            X, Y = pz.X, pz.Y
            hdr = self.headers[i]
            hdr["X"], hdr["Y"] = X, Y
            sx, sy = pz.start
            tx, ty = pz.target
            if not (0 <= sx < X and 0 <= sy < Y and 0 <= tx < X and 0 <= ty < Y):
                raise ValueError(f"puzzle {pz.id}: start/end outside the lattice")
            hdr["start_x"], hdr["start_y"], hdr["target_x"], hdr["target_y"] = sx, sy, tx, ty
            hdr["solution_count"], hdr["difficulty"] = pz.solution_count, int(pz.difficulty)
            hdr["flags"] = 1 if pz.solution_count > 0 else 0
            hdr["trie_root"] = trie.add_puzzle(pz)

            pad = lambda a: np.pad(np.asarray(a), ((0, SX - X), (0, SY - Y)))
            gaps = pz.planes["gaps"] == 1
            valid = np.ones((X, Y), bool)
            cell = np.zeros((X, Y), bool)
            cell[1::2, 1::2] = True
            nb = self.node_boards[i]
            nb[NB_WALK] = _board(pad(~gaps), SY)
            nb[NB_GAPS16] = _board(gaps, 16)
            nb[NB_CELL16] = _board(cell, 16)
            nb[NB_VALID16] = _board(valid, 16)
            if "dot" in pz.planes:
                nb[NB_DOT16] = _board(pz.planes["dot"] == 1, 16)
            if "triangle" in pz.planes:                                      # SG.py:617-636
                interior = np.zeros((X, Y), bool)
                interior[1:X - 1, 1:Y - 1] = True
                tri = (pz.planes["triangle"] == 1) & interior & (pz.info > 0)
                for k in (1, 2, 3, 4):
                    nb[NB_TRI1 + k - 1] = _board(tri & (pz.info == k), 16)
                nb[NB_TRIBAD] = _board(tri & (pz.info > 4), 16)

            # dense one-hot planes in output order; visited / agent_location stay zero,
            # target_location is static per puzzle (SG.py:187)
            bits = np.zeros((self.n_planes, SX, SY), bool)
            for pi, key in enumerate(self.plane_keys):
                if key in ("visited", "agent_location"):
                    continue
                if key == "target_location":
                    bits[pi, tx, ty] = True
                elif key in pz.planes:
                    bits[pi, :X, :Y] = pz.planes[key] == 1
            flat = np.nonzero(bits.ravel())[0]
            np.bitwise_or.at(self.static_bits[i], flat >> 5, np.uint32(1) << (flat & 31).astype(np.uint32))
            if pz.color.min() < 0 or pz.color.max() > 8:
                raise ValueError("colour id out of range")
            self.color[i, :X, :Y] = pz.color
            if pz.info.min() < -2 ** 31 or pz.info.max() >= 2 ** 31:
                raise ValueError("additional_info does not fit int32")
            self.info[i, :X, :Y] = pz.info
            idx = np.nonzero(self.color[i].ravel())[0]
            np.bitwise_or.at(self.color4[i], idx >> 3, self.color[i].ravel()[idx].astype(np.uint32) << ((idx & 7) * 4).astype(np.uint32))
            idx = np.nonzero(self.info[i].ravel())[0]
            np.bitwise_or.at(self.info_nz[i], idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))

            # rule-cell boards (SG.py:456-482, 534-614)
            cb = self.cell_boards[i]
            cb[CB_CELLS] = _cell_board(cell)
            if "star" in pz.planes:
                cb[CB_STAR] = _cell_board(pz.planes["star"] == 1)
            if "square" in pz.planes:
                cb[CB_SQUARE] = _cell_board(pz.planes["square"] == 1)
            layers = np.zeros((X, Y), np.int32)
            for key, arr in pz.planes.items():
                if key not in FIXED_KEYS:
                    layers += arr == 1
            if layers[1::2, 1::2].max(initial=0) > 3:
                raise ValueError("a rule cell sits in more than 3 symbol layers")
            cb[CB_SYMB0] = _cell_board((layers & 1) == 1)
            cb[CB_SYMB1] = _cell_board((layers & 2) == 2)
            for c in range(1, 9):
                cb[CB_COLOR1 + c - 1] = _cell_board(pz.color == c)

            # polyshape instances (SG.py:716-736): any nonzero additional_info whose str() names
            # a listed shape; 'poly' iff the poly plane is set there, else 'ylop' (Q10)
            hdr["shape_off"] = len(shapes)
            slots = {}
            n_ylops = 0
            if isinstance(pz.polyshapes, dict):
                xs, ys = np.nonzero(pz.info)
                for x, y in zip(xs.tolist(), ys.tolist()):
                    name = f"{pz.info[x, y]}"
                    if name not in pz.polyshapes:
                        continue
                    if "poly" not in pz.planes:
                        raise ValueError(f"puzzle {pz.id}: shape instance at ({x},{y}) but no 'poly' "
                                         "plane -- the reference crashes on this puzzle (SG.py:734)")
                    if not (x % 2 and y % 2):
                        continue                                  # off-cell: region id -1, ignored (SG.py:664-667)
                    if name not in slots:
                        slots[name] = shape_slot(pz.polyshapes[name], (pz.X - 1) // 2, (pz.Y - 1) // 2)
                    bit = np.uint64(1 << ((x // 2) * 8 + (y // 2)))
                    if pz.planes["poly"][x, y] == 1:
                        slots[name]["poly_cells"] |= bit
                    else:
                        slots[name]["ylop_cells"] |= bit
                        n_ylops += 1
            if len(slots) > MAX_SHAPES:
                raise ValueError(f"puzzle {pz.id}: more than {MAX_SHAPES} distinct polyshapes")
            if n_ylops > MAX_YLOPS_WIDE:
                raise ValueError(f"puzzle {pz.id}: more than {MAX_YLOPS_WIDE} ylop instances (the reference accepts them, but "
                                 "its exact-fit search enumerates 49^k anchor combinations, SG.py:794-816)")
            if n_ylops > MAX_YLOPS:
                hdr["flags"] |= 4           # wide demand counters in the exact-fit search
            hdr["n_shapes"] = len(slots)
            if slots or cb[CB_STAR] or cb[CB_SQUARE]:
                hdr["flags"] |= 2           # some region rule (squares / stars / shapes) can fail
            # biggest shapes first: the exact-fit search tries slots in this order, and placing
            # the rigid pieces early finds tilings (or dead ends) much sooner.  The answer does
            # not depend on the order (the reference's own order is by dataset position).
            shapes.extend(sorted(slots.values(), key=lambda r: -(int(r["anchors"]) >> 56)))

        self.shapes = np.array(shapes, SHAPE_DTYPE) if shapes else np.zeros(1, SHAPE_DTYPE)
        self.n_shapes_total = len(shapes)
        self.trie = np.array(trie.nodes, np.uint32)
        self.n_puzzles = n

    # -- convenience
    @property
    def obs_elems_per_env(self):
        return self.n_planes * self.slab_x * self.slab_y

    def algorithmic_bytes_per_step(self, obs_format="reference"):
        """B_alg of SURVEY.md section 8d: full observation rewrite (reference dtypes int32 / int32 / int64, or the
        compact uint8 / uint8 / int32 format, which is reported separately) + the state record read and written
        + action/reward/flags."""
        S = self.slab_x * self.slab_y
        per_node = (4 * self.n_planes + 4 + 8) if obs_format == "reference" else (self.n_planes + 1 + 4)
        return S * per_node + 192 + 16
