#!/usr/bin/env python
"""Write-only HBM ceiling on this GPU: how fast can *anything* fill a 1.15 GB buffer (the size of
one config-4 observation rewrite)?  Context for roofline.frac, whose denominator is the
driver's read+write copy bandwidth (MEASURED_PEAKS.json)."""
import json

import torch

n = 131072 * 8788 // 4
x = torch.empty(n, dtype=torch.int32, device="cuda")
y = torch.empty(n, dtype=torch.int32, device="cuda")


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


res = {}
for name, f, nbytes in (("zero_ (memset)", lambda: x.zero_(), 4 * n), ("fill_(1)", lambda: x.fill_(1), 4 * n),
                        ("copy_ (read+write)", lambda: y.copy_(x), 8 * n)):
    ms = timeit(f)
    res[name] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1)}
print(json.dumps(res))
