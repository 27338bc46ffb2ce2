"""NCCL broadcast of one 8192^2 FP64 matrix (512 MiB) between the ranks of a box: time per broadcast, whole and in
8 panels.  torchrun --nproc-per-node N tools/ubench/nccl_bcast.py   (NCCL_* env variants are set by the caller)"""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 8192
b = torch.rand(n * n, dtype=torch.float64, device="cuda")
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
whole = timed(lambda: dist.broadcast(b, src=0))
def panels():
    for p in range(8):
        dist.broadcast(b[p * n * n // 8:(p + 1) * n * n // 8], src=0)
pan = timed(panels)
if rank == 0:
    tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("NCCL_"))
    print(f"world={world} [{tag}] broadcast 512 MiB: whole {whole:.3f} ms ({n*n*8/whole/1e6:.0f} GB/s), 8 panels {pan:.3f} ms", flush=True)
dist.destroy_process_group()
