"""Host <-> device copy rates when several GPUs of one box move pinned host data at once: the floor under the sharded
host-operand entry (every rank uploads its rows of A and downloads its rows of C at the same time).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ubench/pcie_multi.py [--no-bind]
For every concurrency level c in {1, 2, 4, ..., N}: ranks 0..c-1 copy 512 MiB H2D, D2H, and both at once (two streams),
the others idle; prints per-rank and aggregate GB/s (CUDA events per rank, max over the active ranks)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (bind_to_gpu_numa_node)

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bind = "--no-bind" not in sys.argv
if bind:
    bench.bind_to_gpu_numa_node(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

MiB = 1 << 20
size = 512 * MiB
h_in = torch.empty(size, dtype=torch.uint8).pin_memory()
h_out = torch.empty(size, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(2)
d_in = torch.empty(size, dtype=torch.uint8, device="cuda")
d_out = torch.ones(size, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
if rank == 0:
    try:
        import pynvml
        pynvml.nvmlInit()
        for i in range(world):
            hnd = pynvml.nvmlDeviceGetHandleByIndex(i)
            mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
            cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
            print(f"GPU {i}: NVML-local CPUs {cpus[0]}-{cpus[-1]} ({len(cpus)})", flush=True)
    except Exception as e:  # noqa: BLE001
        print("nvml:", e)
    print(f"host: {os.cpu_count()} logical CPUs; NUMA binding of the ranks: {'on' if bind else 'off'}", flush=True)
print(f"rank {rank}: runs on CPUs {sorted(os.sched_getaffinity(0))[:1]}..{sorted(os.sched_getaffinity(0))[-1:]} "
      f"({len(os.sched_getaffinity(0))})", flush=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(kind: str, active: bool, reps: int = 4) -> float:
    """ms per repetition on this rank (0 when idle)"""
    barrier()
    if not active:
        barrier()
        return 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if kind in ("h2d", "duplex"):
            s1.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "duplex"):
            s2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    barrier()
    return ms


levels = sorted({c for c in (1, 2, 4, 8, world) if c <= world})
for kind in ("h2d", "d2h", "duplex"):
    for c in levels:
        run(kind, rank < c, 1)   # warm-up
        ms = run(kind, rank < c)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            per = [float(x.item()) for x in out]
        else:
            per = [ms]
        if rank == 0:
            act = [p for p in per if p > 0]
            gbs = [size / p / 1e6 for p in act]
            dirs = 2 if kind == "duplex" else 1
            print(f"{kind:6s} {c} GPU(s) at once: per rank {', '.join(f'{g:5.1f}' for g in gbs)} GB/s per direction; "
                  f"aggregate {dirs * c * size / max(act) / 1e6:6.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
