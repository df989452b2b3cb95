"""CPU (no GPU needed): how big the exact-fit searches of terminal validation are under random play on the config-4 pool --
the device code compiled for the host (tests/emu) counts the nodes of every search; histogram by floor(log2(nodes)),
scaled to one step of 131,072 envs.  python tools/fit_hist.py   (about a minute)"""
import sys, ctypes as C, numpy as np, time
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
from emu.emu_env import EmuVecEnv, build_emu
from sparc_gym_b200.dataset import config_pool
lib = C.CDLL(build_emu())
rows = config_pool(4, n_puzzles=4096)
N, T = 8192, 1200
env = EmuVecEnv(rows, N, tile=8, traceback=False, max_steps=1000, autoreset_mode="same_step", validate="terminal", defer_validation=True)
env.reset()
rng = np.random.default_rng(0)
out = (C.c_longlong * 24)()
lib.emu_fit_hist(out, 1)
t0 = time.time()
for t in range(T):
    env.step(rng.integers(0, 4, N).astype(np.uint8))
lib.emu_fit_hist(out, 1)
st = env.stats()
print("episodes", int(st[0]), "env-steps", N * T, "time %.1fs" % (time.time() - t0))
h = list(out)
tot = sum(h)
print("searches", tot, "= %.3f per episode" % (tot / max(int(st[0]), 1)))
for b, c in enumerate(h):
    if c: print(f"  nodes in [{1<<b}, {(2<<b)-1}]: {c}  ({c / (N*T) * 131072:.2f} per 131,072-env step)")
