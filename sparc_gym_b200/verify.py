"""Batched path verification: the env-side work of the reference's LLM evaluation host
(llm_testing/llm_host.py:63-242 drives one env per puzzle, one action per LLM turn, and reads
`reward`, `terminated`, `truncated` and `info['rule_status']`) for thousands of (puzzle, proposed
path) pairs at once: every pair gets its own env, the paths are replayed in lockstep without
writing observations, and the verdicts come back as arrays.
"""
from __future__ import annotations

import numpy as np
import torch

from .store import DIRS
from .vec_env import SparcVecEnv

RULE_NAMES = ("reached_target", "path_not_crossing", "no_gap_violations", "all_dots_collected",
              "square_color_separation", "star_pairing_exact", "triangles_edge_count",
              "poly_ylop_area", "all_rules_satisfied")


def path_to_actions(path):
    """[(x, y), ...] node path -> action ids (right, up, left, down; SG.py:12-25); a step that is not a
    unit move becomes the invalid action 4 (the env ignores it, as `action in legal` would, SG.py:1137)."""
    out = []
    for a, b in zip(path, path[1:]):
        step = (int(b[0]) - int(a[0]), int(b[1]) - int(a[1]))
        out.append(DIRS.index(step) if step in DIRS else 4)
    return out


def verify_paths(rows, puzzle_index, paths, traceback=False, device="cuda:0"):
    """Replay `paths[k]` (node paths starting at the puzzle's start, or action lists) on puzzle
    `puzzle_index[k]`.  Returns a dict of numpy arrays, one entry per pair:
      solved        the path is one of the listed solutions (reward +1, SG.py:1205-1209)
      terminated / truncated, steps (actions consumed until the episode ended, or all of them),
      return_cents  sum of rewards x100, rule flags by name at the state where the replay stopped."""
    n = len(paths)
    acts = [p if (len(p) == 0 or np.isscalar(p[0])) else path_to_actions(p) for p in paths]
    T = max((len(a) for a in acts), default=0)
    table = np.full((max(T, 1), n), 4, np.uint8)               # padding = invalid action: a no-op step
    for k, a in enumerate(acts):
        table[:len(a), k] = a
    env = SparcVecEnv(rows, n, traceback=traceback, max_steps=max(T, 1) + 1, autoreset_mode="disabled",
                      validate="never", device=device, write_obs=False)
    env.reset(puzzle_idx=torch.as_tensor(np.asarray(puzzle_index, np.int32)))
    dev_table = torch.from_numpy(table).to(env.device)
    length = torch.from_numpy(np.array([len(a) for a in acts], np.int32)).to(env.device)
    ret = torch.zeros(n, dtype=torch.int32, device=env.device)
    steps = torch.zeros(n, dtype=torch.int32, device=env.device)
    solved = torch.zeros(n, dtype=torch.bool, device=env.device)
    cents = torch.tensor([0, 1, -1, 100, -100], dtype=torch.int32, device=env.device)
    for t in range(T):
        _, _, te, tr, info = env.step(dev_table[t])
        live = (length > t) & (steps == t)                      # still replaying, episode not over
        code = info["reward_code"].long()
        ret += torch.where(live, cents[code], 0)
        solved |= live & (code == 3)
        done = te.bool() | tr.bool()
        steps = torch.where(live, torch.where(done, -(t + 1), t + 1), steps)     # negative = episode over: stop counting
    flags = env.validate().to(torch.int32) & 0xFFFF
    out = {"solved": solved.cpu().numpy(), "terminated": env.terminated.bool().cpu().numpy(),
           "truncated": env.truncated.bool().cpu().numpy(), "steps": steps.abs().cpu().numpy(),
           "return_cents": ret.cpu().numpy()}
    f = flags.cpu().numpy()
    for k, name in enumerate(RULE_NAMES):
        out[name] = ((f >> k) & 1).astype(bool)
    env.close()
    return out
