import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_gym_b200.dataset import config_pool
from sparc_gym_b200.vec_env import SparcVecEnv
rows = config_pool(4, n_puzzles=1024)
env = SparcVecEnv(rows, 131072, traceback=False, max_steps=1000, autoreset_mode="same_step", validate="terminal")
env.reset()
acts = torch.randint(0, 4, (30, 131072), dtype=torch.uint8, device="cuda")
for t in range(10): env.step(acts[t])
out = (C.c_uint64 * 5)()
env.lib.sparc_debug_counters.argtypes = [C.c_void_p, C.c_void_p]
env.lib.sparc_debug_counters(env.handle, out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for t in range(10, 30): env.step(acts[t])
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
env.lib.sparc_debug_counters(env.handle, out)
names = ["prod wait-empty", "prod dynamics", "prod build", "cons wait-full", "cons expand"]
nprod = int(os.environ.get("NPROD", 8)); ncons = int(os.environ.get("NCONS", 8))
print(f"ms/step {ms:.4f}  (kernel cycles at 1.965 GHz: {ms*1e-3*1.965e9:.0f})")
for i, nme in enumerate(names):
    per = out[i] / 20 / 148 / (nprod if i < 3 else ncons)
    print(f"{nme:18s} {per:12.0f} cycles per warp per step")
