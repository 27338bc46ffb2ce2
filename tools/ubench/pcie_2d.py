"""PCIe copy rates of the shapes the host-operand pipeline issues: 1-D against 2-D (strided) copies, both directions,
on an idle GPU and while a product kernel runs.   python tools/ubench/pcie_2d.py"""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import ozimmu_b200 as oz  # noqa: E402

rt = C.CDLL("libcudart.so.12")
n = 8192
host = torch.zeros(n * n, dtype=torch.float64).pin_memory()
dev = torch.zeros(n * n, dtype=torch.float64, device="cuda")
s = torch.cuda.Stream()
H2D, D2H = 1, 2


def copy2d(kind, rows, cols, ld, stream):
    dst, src = (dev.data_ptr(), host.data_ptr()) if kind == H2D else (host.data_ptr(), dev.data_ptr())
    rc = rt.cudaMemcpy2DAsync(C.c_void_p(dst), C.c_size_t(ld * 8), C.c_void_p(src), C.c_size_t(ld * 8), C.c_size_t(rows * 8),
                              C.c_size_t(cols), C.c_int(kind), C.c_void_p(stream.cuda_stream))
    assert rc == 0, rc


def rate(kind, rows, cols, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        copy2d(kind, rows, cols, n, s)
    s.synchronize()
    return rows * cols * 8 * reps / (time.perf_counter() - t0) / 1e9


h = oz.create()
a = torch.rand(n * n, dtype=torch.float64, device="cuda")
c = torch.zeros(n * n, dtype=torch.float64, device="cuda")
shapes = [(8192, 512, "1-D 32 MiB (whole columns)"), (7680, 512, "2-D 7680-row segments"), (768, 7680, "2-D 768-row segments"),
          (512, 8192, "2-D 512-row segments"), (256, 8192, "2-D 256-row segments"), (8192, 8192, "1-D 512 MiB")]
for busy in (False, True):
    for rows, cols, name in shapes:
        if busy:
            for _ in range(3):
                oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, a, n, 0.0, c, n, oz.fp64_int8(9))
        r1, r2 = rate(H2D, rows, cols), 0.0
        if busy:
            for _ in range(3):
                oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, a, n, 0.0, c, n, oz.fp64_int8(9))
        r2 = rate(D2H, rows, cols)
        torch.cuda.synchronize()
        print(f"{'GPU busy (product kernel)' if busy else 'GPU idle':26s} {name:28s}: H2D {r1:5.1f} GB/s   D2H {r2:5.1f} GB/s", flush=True)
oz.destroy(h)
# both directions at once (full duplex): 1-D 512 MiB each way, then the pipeline's mix (2-D H2D of A blocks + 2-D D2H)
s2 = torch.cuda.Stream()
host2 = torch.zeros(n * n, dtype=torch.float64).pin_memory()
dev2 = torch.zeros(n * n, dtype=torch.float64, device="cuda")
for name, (r_in, c_in), (r_out, c_out) in (("1-D both ways", (8192, 8192), (8192, 8192)),
                                             ("H2D 768-row segs + D2H 768-row segs", (768, 8192), (768, 8192))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3 if r_in == 8192 else 30
    for _ in range(reps):
        copy2d(H2D, r_in, c_in, n, s)
        rc = rt.cudaMemcpy2DAsync(C.c_void_p(host2.data_ptr()), C.c_size_t(n * 8), C.c_void_p(dev2.data_ptr()), C.c_size_t(n * 8),
                                  C.c_size_t(r_out * 8), C.c_size_t(c_out), C.c_int(D2H), C.c_void_p(s2.cuda_stream))
        assert rc == 0
    s.synchronize(); t_in = time.perf_counter() - t0
    s2.synchronize(); t_out = time.perf_counter() - t0
    print(f"duplex {name}: H2D {r_in * c_in * 8 * reps / t_in / 1e9:5.1f} GB/s  D2H {r_out * c_out * 8 * reps / t_out / 1e9:5.1f} GB/s", flush=True)
