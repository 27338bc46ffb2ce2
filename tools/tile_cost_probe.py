"""Development probe: the time of ONE round of tiles (72 of 74 CTA pairs busy) and of a second round, per tile shape and
k -- the numbers behind the dispatch cost model (csrc/gemm_fused.cu: dispatch_fused).  Product kernel only (operands
pre-split), through the kernel-level C-ABI.   usage: python tools/tile_cost_probe.py [k,k,...]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ozimmu_b200 as oz  # noqa: E402

ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1024, 2048, 4096, 8192]
s = 9
L = oz.lib()
st = int(torch.cuda.current_stream().cuda_stream)


def timed(fn, iters=10, warm=3):
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


shapes = [(0, 256), (0, 240), (0, 224), (0, 208), (0, 192), (0, 128), (64, 128)]
m = 2048
for k in ks:
    bits = int(L.ozk_bits_per_int8(k))
    pitch = int(L.ozk_slice_pitch(k))
    nmax = 18 * 256
    g = torch.Generator(device="cuda").manual_seed(k)
    a = torch.rand(m * k, dtype=torch.float64, device="cuda", generator=g)
    b = torch.rand(k * nmax, dtype=torch.float64, device="cuda", generator=g)
    a_sl = torch.empty(int(L.ozk_slices_bytes(m, k, s)), dtype=torch.int8, device="cuda")
    b_sl = torch.empty(int(L.ozk_slices_bytes(nmax, k, s)), dtype=torch.int8, device="cuda")
    amax = torch.empty(m, dtype=torch.float64, device="cuda")
    bmax = torch.empty(nmax, dtype=torch.float64, device="cuda")
    scr = torch.zeros(nmax, dtype=torch.int32, device="cuda")
    c = torch.zeros(m * nmax, dtype=torch.float64, device="cuda")
    assert L.ozk_split_int8(a_sl.data_ptr(), pitch, amax.data_ptr(), scr.data_ptr(), m, k, a.data_ptr(), m, 1, s, bits, st) == 0
    assert L.ozk_split_int8(b_sl.data_ptr(), pitch, bmax.data_ptr(), scr.data_ptr(), nmax, k, b.data_ptr(), k, 0, s, bits, st) == 0
    torch.cuda.synchronize()
    for rep in range(2):
        for cm, bn in shapes:
            L.ozk_set_cluster_shape(cm, bn)
            # one round: 8 x 9 tiles of 256 x bn (64-row CTAs: 16 x 4 tiles of 128 x 128); two rounds: twice the columns
            cols1 = 9 * bn if cm == 0 else 4 * 128
            res = []
            for cols in (cols1, 2 * cols1):
                # the plane of B has nmax rows; the block launch addresses its first `cols`
                ms = timed(lambda: L.ozk_gemm_i8_fused_block(m, cols, k, a_sl.data_ptr(), m, 0, b_sl.data_ptr(), nmax, 0, pitch,
                                                             amax.data_ptr(), bmax.data_ptr(), s, bits, 1.0, 0.0, c.data_ptr(), m,
                                                             0, st))
                res.append(ms)
            area1 = m * cols1 / (256 * 256)
            print(f"k={k} tile {'256' if cm == 0 else '128'}x{bn}: one round {res[0] * 1e3:7.1f} us, second round "
                  f"{(res[1] - res[0]) * 1e3:7.1f} us; per 256x256 of C: {res[0] * 1e3 / area1 * 72:7.1f} / "
                  f"{(res[1] - res[0]) * 1e3 / area1 * 72:7.1f} us x 72", flush=True)
    L.ozk_set_cluster_shape(0, 0)
    del a, b, a_sl, b_sl, c
