#!/usr/bin/env python
"""Run bench.py over kernel build variants x tile sizes (one process each) and tabulate."""
import glob, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload = sys.argv[1] if len(sys.argv) > 1 else "config4"
tiles = [int(t) for t in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["8", "16", "32", "64"])]
libs = sorted(glob.glob(os.path.join(root, "build/variants/*.so")))
if len(sys.argv) > 3:
    libs = [l for l in libs if any(k in l for k in sys.argv[3].split(","))]
rows = []
for lib in libs:
    for tile in tiles:
        env = dict(os.environ, SPARC_B200_LIB=lib, SPARC_TILE=str(tile))
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", workload, "--steps", "30", "--warmup", "5",
                            "--no-e2e", "--no-cpu-baseline", "--pool", "512"], env=env, capture_output=True, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            rows.append((os.path.basename(lib), tile, j["value"], j["roofline"]["frac"], j["ms_per_step"]))
        except Exception:
            rows.append((os.path.basename(lib), tile, None, None, r.stderr[-200:]))
        print(rows[-1], flush=True)
