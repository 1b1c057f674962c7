"""What the host links give for the copy pattern of one cuCUDecide_frames_u8 step (16 pictures of 1080p, round-2 formats), without
any kernel - the ceiling the e2e figure of bench.py is measured against.

Per picture: D2H 11.21 MB (packed cost tables, 21 980 B/CTU) + 2.07 MB (Outlier as bytes) + 0.13 MB (OBF as bytes) + the per-depth
CU sums, Yc and the per-CTU source HAD; H2D 2 x 2.07 MB (source and reconstruction as bytes).

Single GPU:   python profiles/ubench/pcie_pattern.py
N GPUs of one box (one rank per GPU, all ranks copy at the same time; the figure is the slowest rank's):
              python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/ubench/pcie_pattern.py
Prints one JSON line (rank 0): per-rank and aggregate GB/s down / up and the step time the copies alone take."""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
P = 16
NCTU = 510
D2H = [NCTU * 21980, 1920 * 1080, 480 * 270] + [480 * 4, 1980 * 4, 8040 * 4, 32400 * 4] * 2 + [NCTU * 4, 128]
H2D = [1920 * 1080, 1920 * 1080]
tot_d, tot_h = sum(D2H) * P, sum(H2D) * P


def run(with_h2d, with_d2h, iters=10):
    gd = [torch.empty(s, dtype=torch.uint8, device=dev) for s in D2H]
    hd = [[torch.empty(s, dtype=torch.uint8, pin_memory=True) for s in D2H] for _ in range(P)]
    gh = [torch.empty(s, dtype=torch.uint8, device=dev) for s in H2D]
    hh = [[torch.empty(s, dtype=torch.uint8, pin_memory=True) for s in H2D] for _ in range(P)]
    s1, s3 = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        for p in range(P):
            if with_h2d:
                with torch.cuda.stream(s3):
                    for g, h in zip(gh, hh[p]):
                        g.copy_(h, non_blocking=True)
            if with_d2h:
                with torch.cuda.stream(s1):
                    for g, h in zip(gd, hd[p]):
                        h.copy_(g, non_blocking=True)
    step(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / iters
    x = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return float(x.item())


res = {"ranks": world, "pictures_per_step": P, "d2h_mb_per_step": tot_d / 1e6, "h2d_mb_per_step": tot_h / 1e6}
for name, h, d in (("down_only", False, True), ("up_only", True, False), ("both", True, True)):
    dt = run(h, d)
    res[name] = {"ms_per_step": dt * 1e3,
                 "per_rank_gbs_down": tot_d / dt / 1e9 if d else 0.0, "per_rank_gbs_up": tot_h / dt / 1e9 if h else 0.0,
                 "aggregate_gbs": world * ((tot_d if d else 0) + (tot_h if h else 0)) / dt / 1e9,
                 "ctus_per_s_if_copies_were_all": world * P * NCTU / dt}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
