"""Times ozimmu_gemm_host (host operands, pinned) at n^3 for a list of block schedules and checks every result
against the device entry bit for bit:
python tools/e2e_probe.py [n] [panel:rowblock[:taper[:ENV=V,ENV=V...]] ...]   (0 = whole operand in one piece)
e.g.  python tools/e2e_probe.py 8192 768:768 768:768:0:OZIMMU_B200_E2E_QUEUE=1,OZIMMU_B200_E2E_QUEUE_RESERVE_SMS=16"""
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
want = torch.zeros(n * n, dtype=torch.float64, device="cuda")
oz.gemm(h, 0, 0, n, n, n, 1.0, a.cuda(), n, b.cuda(), n, 0.0, want, n, oz.fp64_int8(9))
torch.cuda.synchronize()
want = want.cpu()
touched = set()
for combo in combos:
    parts = combo.split(":", 3)
    panel, rowblock = parts[0], parts[1]
    taper = parts[2] if len(parts) > 2 else "0"
    extra = parts[3] if len(parts) > 3 else ""
    for key in touched:
        os.environ.pop(key, None)
    touched.clear()
    os.environ["OZIMMU_B200_E2E_PANEL"], os.environ["OZIMMU_B200_E2E_ROWBLOCK"] = panel, rowblock
    os.environ["OZIMMU_B200_E2E_TAPER"] = taper
    for kv in filter(None, extra.split(",")):
        key, val = kv.split("=", 1)
        os.environ[key] = val
        touched.add(key)
    for _ in range(2):
        oz.gemm_host(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(9))
    it = 5
    times = []
    for _ in range(it):
        t0 = time.perf_counter()
        oz.gemm_host(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(9))
        times.append(time.perf_counter() - t0)
    ok = torch.equal(c.view(torch.int64), want.view(torch.int64))
    dt = sum(times) / it
    print(f"gemm_host n={n} panel={panel} rowblock={rowblock} taper={taper} {extra}: mean {dt*1e3:.2f} ms best {min(times)*1e3:.2f} "
          f"worst {max(times)*1e3:.2f} ms  {2*n**3/dt/1e12:.2f} TFLOP/s-equiv  bit-identical={ok}", flush=True)
# raw PCIe numbers for context
d = torch.empty(n * n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize()
print(f"H2D {n*n*8/2**20:.0f} MiB: {(time.perf_counter()-t0)*1e3:.2f} ms")
t0 = time.perf_counter(); c.copy_(d, non_blocking=True); torch.cuda.synchronize()
print(f"D2H {n*n*8/2**20:.0f} MiB: {(time.perf_counter()-t0)*1e3:.2f} ms")
oz.destroy(h)
