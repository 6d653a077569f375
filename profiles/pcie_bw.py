import torch, time
x = torch.empty(200_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty(200_000_000, dtype=torch.uint8, device="cuda")
h = torch.empty(136_000_000, dtype=torch.uint8).pin_memory()
dd = torch.empty(136_000_000, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("h2d", lambda: d.copy_(x, non_blocking=True)), ("d2h", lambda: h.copy_(dd, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.time()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.time() - t) / 5
    print(name, "GB/s", (200e6 if name == "h2d" else 136e6) / dt / 1e9)
torch.cuda.synchronize(); t = time.time()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): h.copy_(dd, non_blocking=True)
torch.cuda.synchronize()
print("both ms per pair", (time.time() - t) / 5 * 1e3)
