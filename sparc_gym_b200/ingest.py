"""Host-side puzzle ingest: dataset rows -> per-puzzle planes.

Mirrors what the reference builds once in its constructor (`_process_puzzles`,
reference SPaRC_Gym/SPaRC_Gym.py:219-368 -- "SG.py" below): for every row a dict of int32
one-hot planes `[X, Y]` (`visited, gaps, agent_location, target_location` first, then one plane
per symbol kind in order of first appearance), an int32 colour plane, an int64
`additional_info` plane, start / target, the listed solution paths and the polyshape table.

The reference's ingest has order-dependent behaviour that shows up in its observations, so it
is reproduced here rather than "fixed" (SURVEY quirk ledger):

  Q11  inside a cell's `properties` every key -- not only `type` / `dot` -- touches the plane
       of the *current symbol*, which is whatever the last `type` / `dot` key set, possibly in
       an earlier cell;
  Q12  that current symbol even survives from one puzzle to the next, so a puzzle that opens
       with a gap-only cell gets an (all-zero) plane of its predecessor's last symbol, and a
       dataset that opens that way cannot be ingested at all (the reference raises
       UnboundLocalError; so do we);
  Q13  the plane set and its key order are per puzzle.

This is plain host Python like the reference's own ingest; the device never sees rows.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np
import yaml

_Loader = getattr(yaml, "CSafeLoader", yaml.SafeLoader)

FIXED_KEYS = ("visited", "gaps", "agent_location", "target_location")
SYMBOL_KEYS = ("star", "square", "triangle", "poly", "ylop", "dot")
COLOR_IDS = {"red": 1, "blue": 2, "yellow": 3, "green": 4, "black": 5, "purple": 6, "orange": 7,
             "white": 8}  # SG.py:310


def _parse(text):
    """The dataset stores YAML strings; JSON is a YAML subset and parses ~20x faster."""
    if isinstance(text, (dict, list)) or text is None:
        return text
    s = text.lstrip()
    if s[:1] in "{[":
        try:
            return json.loads(s)
        except ValueError:
            pass
    return yaml.load(text, Loader=_Loader)


@dataclass
class Puzzle:
    """One ingested puzzle (the reference's per-puzzle dict, SG.py:236-366)."""
    id: str
    difficulty: int
    X: int
    Y: int
    start: tuple
    target: tuple
    solution_count: int
    solution_paths: list
    polyshapes: object
    keys: list                      # plane key order of this puzzle (Q13)
    planes: dict                    # key -> int32[X, Y] (visited / agent_location / target all zero)
    color: np.ndarray               # int32[X, Y]
    info: np.ndarray                # int64[X, Y]
    observ: object = None           # puzzle_array, only for observation='SPaRC'
    extra: dict = field(default_factory=dict)


class _CurrentSymbol:
    """The reference's function-scoped `symbol` variable (SG.py:288-343): it is assigned by
    `type` / `dot` (and `gap` in the second pass) and read by every key."""
    __slots__ = ("name",)

    def __init__(self):
        self.name = None

    def get(self):
        if self.name is None:
            # what Python raises in the reference when the first key ever seen is not type/dot
            raise UnboundLocalError("cannot access local variable 'symbol' where it is not "
                                    "associated with a value")
        return self.name


def ingest_rows(rows, observation="new"):
    """rows (dataset-schema dicts, or a pandas DataFrame) -> list[Puzzle]."""
    if hasattr(rows, "to_dict") and not isinstance(rows, (list, tuple)):
        rows = rows.to_dict("records")
    cur = _CurrentSymbol()
    out = []
    for row in rows:
        w, h = int(row["grid_size"]["width"]), int(row["grid_size"]["height"])
        X, Y = 2 * w + 1, 2 * h + 1                                    # SG.py:243-247
        tv = _parse(row["text_visualization"])["puzzle"]
        cells = tv.get("cells") or []
        planes = {k: np.zeros((X, Y), np.int32) for k in FIXED_KEYS}
        color = np.zeros((X, Y), np.int32)
        info = np.zeros((X, Y), np.int64)

        # pass 1 (SG.py:283-325): create planes, colours, additional info
        for cell in cells:
            props = cell.get("properties", {}) or {}
            pos = cell.get("position", {}) or {}
            px, py = pos.get("x"), pos.get("y")
            col = cnt = shp = None
            for key, value in props.items():
                if key == "type":
                    cur.name = f"{value}"
                    col = props.get("color", "")
                    if value == "triangle":
                        cnt = props.get("count", "")
                    elif value not in ("star", "square"):
                        shp = props.get("polyshape", "")
                elif key == "dot":
                    cur.name = "dot"
                name = cur.get()
                if name not in planes:
                    planes[name] = np.zeros((X, Y), np.int32)
                if col and col in COLOR_IDS:
                    color[px][py] = COLOR_IDS[col]
                if cnt:
                    info[px][py] = cnt
                elif shp:
                    info[px][py] = shp
        # pass 2 (SG.py:329-343): set the one-hot bits
        for cell in cells:
            props = cell.get("properties", {}) or {}
            pos = cell.get("position", {}) or {}
            px, py = pos.get("x"), pos.get("y")
            for key, value in props.items():
                if key == "type":
                    cur.name = f"{value}"
                elif key == "dot":
                    cur.name = "dot"
                elif key == "gap":
                    cur.name = "gaps"
                name = cur.get()
                if name in planes:
                    planes[name][px, py] = 1
        planes["gaps"][1:X - 1:2, 1:Y - 1:2] = 1                       # SG.py:345-351

        out.append(Puzzle(
            id=row["id"], difficulty=row["difficulty_level"], X=X, Y=Y,
            start=(tv["start"]["x"], tv["start"]["y"]), target=(tv["end"]["x"], tv["end"]["y"]),
            solution_count=int(row["solution_count"]),
            solution_paths=[[[pt["x"], pt["y"]] for pt in item["path"]] for item in row["solutions"]],
            polyshapes=_parse(row["polyshapes"]),
            keys=list(planes.keys()), planes=planes, color=color, info=info,
            observ=row.get("puzzle_array") if observation == "SPaRC" else None,
        ))
    return out
