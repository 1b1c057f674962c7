import torch, time
dev=torch.device('cuda',0)
def bw(nbytes, d2h=True, both=False, iters=5):
    g=torch.empty(nbytes, dtype=torch.uint8, device=dev); h=torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    g2=torch.empty(nbytes//2, dtype=torch.uint8, device=dev); h2=torch.empty(nbytes//2, dtype=torch.uint8, pin_memory=True)
    s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
    torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(iters):
        with torch.cuda.stream(s1):
            if d2h: h.copy_(g, non_blocking=True)
            else: g.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): g2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    dt=(time.perf_counter()-t)/iters
    return nbytes/dt/1e9, dt*1e3
for n in (283_000_000, 64_000_000, 13_000_000):
    print(n, "D2H GB/s, ms:", bw(n), " H2D:", bw(n, d2h=False), " D2H with concurrent H2D of half:", bw(n, both=True))
# many small copies
g=torch.empty(283_000_000, dtype=torch.uint8, device=dev); h=torch.empty(283_000_000, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); t=time.perf_counter()
for it in range(5):
    for k in range(16):
        o=k*17_000_000
        h[o:o+12_900_000].copy_(g[o:o+12_900_000], non_blocking=True)
        h[o+12_900_000:o+17_000_000].copy_(g[o+12_900_000:o+17_000_000], non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("32 copies totalling", 16*17, "MB:", 16*17e6/dt/1e9, "GB/s", dt*1e3, "ms")
