"""Times ozimmu_gemm_host (host operands, pinned) at n^3 for a list of block schedules:
python tools/e2e_probe.py [n] [panel:rowblock[:taper] ...]   (0 = whole operand in one piece; default sweep below)"""
import os, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ozimmu_b200 as oz
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
combos = sys.argv[2:] or ["1024:0", "1024:1024", "512:512", "768:768", "1536:1536", "2048:2048", "512:1024", "1024:512"]
a = torch.rand(n * n, dtype=torch.float64).pin_memory()
b = torch.rand(n * n, dtype=torch.float64).pin_memory()
c = torch.zeros(n * n, dtype=torch.float64).pin_memory()
h = oz.create()
for combo in combos:
    panel, rowblock, taper = (combo.split(":") + ["1"])[:3]
    os.environ["OZIMMU_B200_E2E_PANEL"], os.environ["OZIMMU_B200_E2E_ROWBLOCK"] = panel, rowblock
    os.environ["OZIMMU_B200_E2E_TAPER"] = taper
    for _ in range(2):
        oz.gemm_host(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(9))
    it = 5
    best, t_all = 1e9, time.perf_counter()
    for _ in range(it):
        t0 = time.perf_counter()
        oz.gemm_host(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(9))
        best = min(best, time.perf_counter() - t0)
    dt = (time.perf_counter() - t_all) / it
    print(f"gemm_host n={n} panel={panel} rowblock={rowblock} taper={taper}: mean {dt*1e3:.2f} ms best {best*1e3:.2f} ms  "
          f"{2*n**3/dt/1e12:.2f} TFLOP/s-equiv", flush=True)
# raw PCIe numbers for context
d = torch.empty(n * n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize()
print(f"H2D {n*n*8/2**20:.0f} MiB: {(time.perf_counter()-t0)*1e3:.2f} ms")
t0 = time.perf_counter(); c.copy_(d, non_blocking=True); torch.cuda.synchronize()
print(f"D2H {n*n*8/2**20:.0f} MiB: {(time.perf_counter()-t0)*1e3:.2f} ms")
oz.destroy(h)
