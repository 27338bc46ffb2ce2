"""Development probe: what does cutting ONE product into several launches cost, with nothing else going on?
gemm_streamed_b on one GPU with every panel's event already fired, for several panel counts and launch modes,
against the plain gemm.   python tools/streamed_probe.py [n]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ozimmu_b200 as oz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)
b = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)
c = torch.zeros(n * n, dtype=torch.float64, device="cuda")
want = torch.zeros_like(c)
h = oz.create()
mode = oz.fp64_int8(9)


def timed(fn, iters=6, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ev = torch.cuda.Event()
ev.record()
torch.cuda.synchronize()
oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, want, n, mode)
for rep in range(2):
    print(f"plain gemm: {timed(lambda: oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, mode)):.2f} ms", flush=True)
    for panels in (1, 2, 4, 8, 16):
        w = -(-n // panels // 256) * 256
        edges = list(range(0, n, w)) + [n]
        for one_tile in ("1", "0"):
            os.environ["OZIMMU_B200_STREAMED_ONE_TILE"] = one_tile
            ms = timed(lambda: oz.gemm_streamed_b(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, mode, edges,
                                                  [ev.cuda_event] * (len(edges) - 1)))
            ok = torch.equal(c.view(torch.int64), want.view(torch.int64))
            print(f"streamed_b panels={len(edges) - 1} one_tile={one_tile}: {ms:.2f} ms  bit-identical={ok}", flush=True)
oz.destroy(h)
