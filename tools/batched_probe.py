"""Strided-batched DGEMM: grouped launch vs one call per entry vs the reference (which loops, src/cublas.cu:380-406).
python tools/batched_probe.py"""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ozimmu_b200 as oz
import oracle_lib

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

h = oz.create()
ref = None
if oracle_lib.reference() is not None:
    from gpu_util import Reference
    ref = Reference()
for (n, batch) in [(1024, 8), (1024, 32), (1536, 16), (2048, 16), (4096, 4)]:
    a = torch.rand(batch * n * n, dtype=torch.float64, device="cuda")
    b = torch.rand(batch * n * n, dtype=torch.float64, device="cuda")
    c = torch.zeros(batch * n * n, dtype=torch.float64, device="cuda")
    c2 = torch.zeros_like(c)
    mode = oz.fp64_int8(9)
    def loop():
        for e in range(batch):
            oz.gemm(h, 0, 0, n, n, n, 1.0, a[e * n * n:], n, b[e * n * n:], n, 0.0, c[e * n * n:], n, mode)
    def grouped():
        oz.gemm_strided_batched(h, 0, 0, n, n, n, 1.0, a, n, n * n, b, n, n * n, 0.0, c2, n, n * n, batch, mode)
    t_loop, t_grp = timed(loop), timed(grouped)
    assert torch.equal(c.view(torch.int64), c2.view(torch.int64))
    line = f"n={n} batch={batch}: per-entry calls {t_loop:.3f} ms, grouped {t_grp:.3f} ms ({t_loop / t_grp:.2f}x), " \
           f"{2 * n**3 * batch / t_grp / 1e9:.1f} TFLOP/s-equiv"
    if ref is not None:
        c3 = torch.zeros_like(c)
        def ref_loop():
            for e in range(batch):
                ref.gemm(0, 0, n, n, n, 1.0, a[e * n * n:], n, b[e * n * n:], n, 0.0, c3[e * n * n:], n, mode)
        t_ref = timed(ref_loop, 3)
        line += f"; reference (loops) {t_ref:.3f} ms ({t_ref / t_grp:.2f}x), bit-identical={torch.equal(c3.view(torch.int64), c2.view(torch.int64))}"
    print(line, flush=True)
oz.destroy(h)
