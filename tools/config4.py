"""BASELINE config 4: 16384^3 DGEMM fp64_int8_9 row-sharded across the GPUs of one box (strong scaling):
rank g owns rows [g*16384/G, (g+1)*16384/G) of A and C, rank 0 owns B and broadcasts it (NCCL) inside the
timed region.  Launch: python -m torch.distributed.run --nproc-per-node G tools/config4.py [n] [steps]
Prints one JSON line on rank 0 (per-GPU and aggregate FP64-equivalent TFLOP/s, max over ranks)."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ozimmu_b200 as oz  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r0, rows = oz.row_block(n, world, rank)
    g = torch.Generator(device="cuda").manual_seed(rank)
    a = 1.0 - torch.rand(rows * n, dtype=torch.float64, device="cuda", generator=g)      # rows x n col-major, ld = rows
    b = (1.0 - torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)) if rank == 0 else \
        torch.zeros(n * n, dtype=torch.float64, device="cuda")
    c = torch.zeros(rows * n, dtype=torch.float64, device="cuda")
    h = oz.create()

    def step():
        assert oz.sharded_gemm(h, 0, 0, rows, n, n, 1.0, a, rows, b, n, 0.0, c, rows, oz.fp64_int8(9), src=0) == 0

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        flop = 2.0 * n ** 3
        print(json.dumps({"config": f"{n}^3 fp64_int8_9 row-sharded over {world} GPU(s), B broadcast from rank 0 each step",
                          "n_gpus": world, "rows_per_gpu": rows, "ms_per_step": ms, "aggregate_tflops": flop / ms / 1e9,
                          "per_gpu_tflops": flop / ms / 1e9 / world, "steps": steps}), flush=True)
    oz.destroy(h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
