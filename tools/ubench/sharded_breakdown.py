"""Where a sharded step's time goes at N ranks: product alone, broadcast alone, broadcast + product (per rank).
torchrun --nproc-per-node N tools/ubench/sharded_breakdown.py"""
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import ozimmu_b200 as oz
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 8192
g = torch.Generator(device="cuda").manual_seed(rank)
a = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)
b = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)
c = torch.zeros(n * n, dtype=torch.float64, device="cuda")
h = oz.create()
mode = oz.fp64_int8(9)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
variants = {
    "product alone": lambda: oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, mode),
    "broadcast alone": lambda: dist.broadcast(b, src=0),
    "broadcast+product": lambda: oz.sharded_gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, mode, src=0),
    "streamed-B (NCCL panels)": lambda: oz.sharded_gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, mode, src=0, pipeline=True),
    "peer-pull pipeline": lambda: oz.sharded_gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, mode, src=0, transport="peer"),
}
# the GPUs slow down by ~6 % as they heat up over such a run: interleave the variants instead of timing them in turn
acc = {k: [] for k in variants}
for cycle in range(4):
    for name, fn in variants.items():
        acc[name].append(timed(fn, 5))
mine = {k: (sum(v[1:]) / len(v[1:])) for k, v in acc.items()}   # first cycle = warm-up
out = [None] * world
dist.all_gather_object(out, (rank, mine))
if rank == 0:
    for r, d in out:
        print(f"world={world} rank {r}: " + ", ".join(f"{k} {v:.3f} ms" for k, v in d.items()), flush=True)
oz.destroy(h)
dist.destroy_process_group()
