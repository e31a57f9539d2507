"""Measures pinned host<->device copy bandwidth on this box (context for bench.py's e2e numbers)."""
import json
import torch

def bw(nbytes, direction, reps=5):
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    best = 0.0
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if direction == "h2d":
            d.copy_(h, non_blocking=True)
        else:
            h.copy_(d, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
    return best

if __name__ == "__main__":
    out = {"h2d_GBps": bw(1 << 30, "h2d"), "d2h_GBps": bw(1 << 30, "d2h"), "d2h_3MB_GBps": bw(3 << 20, "d2h", 20)}
    # both directions at once
    h1 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); d1 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    out["duplex_each_GBps"] = (1 << 30) / (a.elapsed_time(b) * 1e-3) / 1e9
    print(json.dumps(out))
