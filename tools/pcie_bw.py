"""Pinned host <-> device copy bandwidth on this box (explains the e2e bound of bench.py)."""
import time, torch
n = 64 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=20):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return reps * n / (time.perf_counter() - t) / 1e9
print("H2D GB/s", round(run(lambda: d.copy_(h, non_blocking=True)), 1))
print("D2H GB/s", round(run(lambda: h.copy_(d, non_blocking=True)), 1))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("concurrent H2D+D2H GB/s each", round(run(both), 1))
