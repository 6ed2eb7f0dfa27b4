"""torchrun --nproc-per-node N scripts/pcie_probe.py: raw pinned H2D / D2H rates of all ranks at the same time, with and
without binding each rank to the CPUs NVML reports as local to its GPU (host-side limits of the e2e figure)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 512 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
for bind in (False, True):
    before = os.sched_getaffinity(0)
    prev = bench.gpu_local_affinity(local) if bind else None
    now = os.sched_getaffinity(0)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out.fill_(2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): d.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): h_out.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = 4 * n * (2 if mode == "both" else 1) / float(t.item()) / 1e9
    if prev:
        os.sched_setaffinity(0, prev)
    allc = [None] * world
    dist.all_gather_object(allc, (rank, len(before), len(now), min(now), max(now)))
    if rank == 0:
        print(f"bind={bind}: per-rank GB/s with all {world} ranks copying: " + ", ".join(f"{k} {v:.1f}" for k, v in res.items()), flush=True)
        print("   cpus (rank, before, bound, min, max):", allc, flush=True)
dist.destroy_process_group()
