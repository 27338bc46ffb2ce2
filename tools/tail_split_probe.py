"""Development probe: does cutting the LAST partial round of tiles off a product and running it with narrower tiles pay?
One product launch against [columns 0..n1) with 256-wide tiles + [n1..n) with narrower ones, same stream, through the
kernel-level C-ABI (ozk_split_int8 + ozk_gemm_i8_fused / _block); bits compared with the single launch.
usage: python tools/tail_split_probe.py [n] [m]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ozimmu_b200 as oz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = int(sys.argv[2]) if len(sys.argv) > 2 else n
k, s = n, 9
L = oz.lib()
bits = int(L.ozk_bits_per_int8(k))
pitch = int(L.ozk_slice_pitch(k))
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.rand(m * k, dtype=torch.float64, device="cuda", generator=g)   # m x k column-major (op_n): rows strided
b = torch.rand(k * n, dtype=torch.float64, device="cuda", generator=g)   # k x n column-major (op_n): columns contiguous
a_sl = torch.empty(int(L.ozk_slices_bytes(m, k, s)), dtype=torch.int8, device="cuda")
b_sl = torch.empty(int(L.ozk_slices_bytes(n, k, s)), dtype=torch.int8, device="cuda")
amax = torch.empty(m, dtype=torch.float64, device="cuda")
bmax = torch.empty(n, dtype=torch.float64, device="cuda")
scr = torch.zeros(max(m, n), dtype=torch.int32, device="cuda")
st = int(torch.cuda.current_stream().cuda_stream)
assert L.ozk_split_int8(a_sl.data_ptr(), pitch, amax.data_ptr(), scr.data_ptr(), m, k, a.data_ptr(), m, 1, s, bits, st) == 0
assert L.ozk_split_int8(b_sl.data_ptr(), pitch, bmax.data_ptr(), scr.data_ptr(), n, k, b.data_ptr(), k, 0, s, bits, st) == 0
torch.cuda.synchronize()
c0 = torch.zeros(m * n, dtype=torch.float64, device="cuda")
c1 = torch.zeros(m * n, dtype=torch.float64, device="cuda")


def single(width):
    L.ozk_set_cluster_shape(0, width)
    assert L.ozk_gemm_i8_fused(m, n, k, a_sl.data_ptr(), b_sl.data_ptr(), pitch, amax.data_ptr(), bmax.data_ptr(), s, bits,
                               1.0, 0.0, c0.data_ptr(), m, st) == 0


def parts(plan):
    """plan: [(col0, ncols, width)]"""
    for col0, ncols, width in plan:
        L.ozk_set_cluster_shape(0, width)
        assert L.ozk_gemm_i8_fused_block(m, ncols, k, a_sl.data_ptr(), m, 0, b_sl.data_ptr(), n, col0, pitch, amax.data_ptr(),
                                         bmax.data_ptr() + 8 * col0, s, bits, 1.0, 0.0, c1.data_ptr() + 8 * col0 * m, m, 0, st) == 0


def timed(fn, iters=8, warm=3):
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


pairs = 74
tm = (m + 255) // 256
for rep in range(2):
    for w in (256, 192):
        print(f"n={n} m={m}: one launch, {w}-wide tiles: {timed(lambda: single(w)):.3f} ms", flush=True)
    # every cut after c1 columns of 256-wide tiles that ends a whole number of rounds (or nearly), tail in 192 / 128 / 256
    tried = set()
    for rounds1 in range(1, (tm * ((n + 255) // 256) + pairs - 1) // pairs + 1):
        c1cols = min(rounds1 * pairs // tm, n // 256)
        if c1cols == 0 or c1cols * 256 >= n or c1cols in tried:
            continue
        tried.add(c1cols)
        for w2 in (192, 128, 224):
            plan = [(0, c1cols * 256, 256), (c1cols * 256, n - c1cols * 256, w2)]
            ms = timed(lambda: parts(plan))
            single(256)
            torch.cuda.synchronize()
            same = bool(torch.equal(c0.view(torch.int64), c1.view(torch.int64)))
            t1, t2 = tm * c1cols, tm * ((n - c1cols * 256 + w2 - 1) // w2)
            print(f"  {c1cols * 256} columns x256 ({t1} tiles = {t1 / pairs:.2f} rounds) + {n - c1cols * 256} x{w2} "
                  f"({t2} tiles = {t2 / pairs:.2f} rounds): {ms:.3f} ms  bit-identical {same}", flush=True)
L.ozk_set_cluster_shape(0, 0)
