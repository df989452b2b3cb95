"""Wall-clock cost of one host-buffer call (sparc_step_host through SparcVecEnv.step_host) on launch-bound batches,
beside the device-resident step (CUDA graph) of the same env: config 2 and neighbours.

    gpurun -- python tools/hostcall_probe.py [envs ...]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparc_gym_b200.dataset import config_pool  # noqa: E402
from sparc_gym_b200.vec_env import DeviceStore, SparcVecEnv, _as_store  # noqa: E402

store = DeviceStore(_as_store(config_pool(2, n_puzzles=64)), "cuda:0")
K = 2000
for n in [int(a) for a in sys.argv[1:]] or [4096]:
    for spin in ("1", "0"):          # 0 = the plain copy sequence around the launch
        os.environ["SPARC_HOST_ZERO_COPY"] = spin
        env = SparcVecEnv(store, n, traceback=False, max_steps=1000, autoreset_mode="same_step", validate="terminal")
        env.reset()
        A = torch.randint(0, 4, (K, n), dtype=torch.uint8).pin_memory()      # fresh actions every call (one vector repeated walks every env into a wall)
        a = A[0]
        r = torch.empty(n, dtype=torch.float32).pin_memory()
        te = torch.empty(n, dtype=torch.uint8).pin_memory()
        tr = torch.empty(n, dtype=torch.uint8).pin_memory()
        for _ in range(50):
            env.step_host(a, r, te, tr)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            for t in range(K):
                env.step_host(A[t], r, te, tr)
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / K)
        # the bare ctypes call with prepared arguments: what the Python wrapper adds
        args = [(env.handle, A[t].data_ptr(), r.data_ptr(), te.data_ptr(), tr.data_ptr(), env._obs_p, env._stream()) for t in range(K)]
        f = env.lib.sparc_step_host
        t0 = time.perf_counter()
        for t in range(K):
            f(*args[t])
        torch.cuda.synchronize()
        bare = (time.perf_counter() - t0) / K
        if hasattr(env.lib, "sparc_debug_hosttiming"):          # -DHOST_TIMING builds: where the C call spends its time
            import ctypes as C
            o = (C.c_double * 5)()
            env.lib.sparc_debug_hosttiming(o)
            if o[4]:
                print("    inside the C call, us: pack %.2f  launch %.2f  spin %.2f  copy out %.2f" % tuple(o[k] / o[4] for k in range(4)))
        print(f"envs {n:6d}  SPARC_HOST_ZERO_COPY={spin}: step_host {best * 1e6:6.2f} us per call = {n / best:.3e} env-steps/s; "
              f"bare C call {bare * 1e6:6.2f} us", flush=True)
        env.close()
