"""What the PCIe link gives for the copy pattern of one cuCUDecide_frames step (16 pictures of 1080p), without any kernel:
per picture D2H 12.9 MB (packed cost tables) + 4.15 MB (Outlier) + 0.26 MB (OBF) + 8 small arrays; H2D 2 x 4.15 MB.
Variants: D2H alone / with the H2D traffic on another stream / D2H split over two streams / 2 MB-page (THP) pinned buffers."""
import ctypes, mmap, time
import torch
dev = torch.device("cuda", 0)
P = 16
D2H = [12_925_440, 4_147_200, 259_200] + [2_040, 8_160, 32_160, 129_600] * 2 + [2_040]
H2D = [4_147_200, 4_147_200]
tot_d = sum(D2H) * P; tot_h = sum(H2D) * P

def pinned(n, huge=False):
    if not huge:
        return torch.empty(n, dtype=torch.uint8, pin_memory=True)
    sz = (n + (2 << 20) - 1) // (2 << 20) * (2 << 20)
    m = mmap.mmap(-1, sz + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
    al = (addr + (2 << 20) - 1) // (2 << 20) * (2 << 20)
    libc = ctypes.CDLL("libc.so.6", use_errno=True)
    libc.madvise(ctypes.c_void_p(al), ctypes.c_size_t(sz), 14)            # MADV_HUGEPAGE
    ctypes.memset(ctypes.c_void_p(al), 0, sz)                              # touch
    rc = torch.cuda.cudart().cudaHostRegister(al, sz, 0)
    assert int(rc) == 0, rc
    t = torch.frombuffer((ctypes.c_char * n).from_address(al), dtype=torch.uint8)
    t._keep = m
    return t

def run(huge, with_h2d, two_streams, iters=8):
    gd = [torch.empty(s, dtype=torch.uint8, device=dev) for s in D2H]
    hd = [[pinned(s, huge) for s in D2H] for _ in range(P)]
    gh = [torch.empty(s, dtype=torch.uint8, device=dev) for s in H2D]
    hh = [[pinned(s, huge) for s in H2D] for _ in range(P)]
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    def step():
        for p in range(P):
            if with_h2d:
                with torch.cuda.stream(s3):
                    for g, h in zip(gh, hh[p]): g.copy_(h, non_blocking=True)
            for i, (g, h) in enumerate(zip(gd, hd[p])):
                with torch.cuda.stream(s2 if (two_streams and i > 0) else s1): h.copy_(g, non_blocking=True)
    step(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / iters
    return tot_d / dt / 1e9, dt * 1e3

print("thp:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
for huge in (False, True):
    for with_h2d in (False, True):
        for two in (False, True):
            try:
                gbs, ms = run(huge, with_h2d, two)
                print(f"huge={int(huge)} h2d={int(with_h2d)} two_d2h_streams={int(two)}: D2H {gbs:.1f} GB/s, {ms:.2f} ms per step ({tot_d/1e6:.0f} MB down, {tot_h/1e6 if with_h2d else 0:.0f} MB up)")
            except Exception as e:
                print("huge", huge, "failed:", repr(e)[:200])
