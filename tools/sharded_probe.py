"""Development probe (not the bench): the sharded step on N GPUs for several panel counts, device and host operands.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_probe.py [n] [rows_per_rank] [panels=1,2,4] [nohost]"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ozimmu_b200 as oz  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rows = int(sys.argv[2]) if len(sys.argv) > 2 else n
extra = sys.argv[3:]
panel_list = [int(x) for x in next((e.split("=")[1] for e in extra if e.startswith("panels=")), "1,2,4,8,16").split(",")]
g = torch.Generator(device="cuda").manual_seed(rank)
a = torch.rand(rows * n, dtype=torch.float64, device="cuda", generator=g)
b = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)
c = torch.zeros(rows * n, dtype=torch.float64, device="cuda")
h = oz.create()
comm = oz.comm_create()
mode = oz.fp64_int8(9)


def timed(fn, iters=6, warm=2):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [round(float(x), 2) for x in out]


def report(name, per_rank):
    if rank == 0:
        ms = max(per_rank)
        print(f"{name}: max {ms:.2f} ms  {2.0 * rows * world * n * n / ms / 1e9:.1f} TFLOP/s aggregate  per rank {per_rank}", flush=True)


report("product alone (no broadcast)", timed(lambda: oz.gemm(h, 0, 0, rows, n, n, 1.0, a, rows, b, n, 0.0, c, rows, mode)))
for rep in range(2):
    for panels in panel_list:
        report(f"sharded_gemm panels={panels}",
               timed(lambda: oz.sharded_gemm(h, comm, 0, 0, rows, n, n, 1.0, a, rows, b, n, 0.0, c, rows, mode, src=0,
                                             max_panels=panels)))
os.environ["OZIMMU_B200_STREAMED_ONE_TILE"] = "1"
if "nohost" in extra:
    oz.destroy(h); comm.destroy(); dist.destroy_process_group()
    sys.exit(0)
ha = a.cpu().pin_memory()
hb = b.cpu().pin_memory() if rank == 0 else None
hc = torch.zeros(rows * n, dtype=torch.float64).pin_memory()
for blk in ("768", "1024", "512"):
    os.environ["OZIMMU_B200_E2E_PANEL"] = os.environ["OZIMMU_B200_E2E_ROWBLOCK"] = blk
    report(f"sharded_gemm_host block={blk}",
           timed(lambda: oz.sharded_gemm_host(h, comm, 0, 0, rows, n, n, 1.0, ha, rows, hb, n, 0.0, hc, rows, mode, src=0), 4, 2))
ok = torch.equal(hc.view(torch.int64), c.cpu().view(torch.int64))
print(f"rank {rank}: host-operand result bit-identical to the device path: {ok}", flush=True)
oz.destroy(h); comm.destroy(); dist.destroy_process_group()
