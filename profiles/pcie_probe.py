"""Measures pinned host<->device copy bandwidth on the box (context for the e2e number)."""
import json
import torch

out = {}
for mb in (10, 160, 640):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, (src, dst) in {"h2d": (h, d), "d2h": (d, h)}.items():
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out[f"{name}_{mb}MB_GBs"] = round(5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps(out))
