import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
d.copy_(h, non_blocking=True); h.copy_(d, non_blocking=True)
print("H2D GB/s", n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9)
print("D2H GB/s", n / t(lambda: h.copy_(d, non_blocking=True)) / 1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("bidir GB/s each", n / t(both) / 1e9)
t0 = time.perf_counter(); z = torch.empty(1 << 31, dtype=torch.uint8, pin_memory=True); print("pin alloc 2GiB s", time.perf_counter() - t0)
del z; t0 = time.perf_counter(); z = torch.empty(1 << 31, dtype=torch.uint8, pin_memory=True); print("pin re-alloc 2GiB s", time.perf_counter() - t0)
