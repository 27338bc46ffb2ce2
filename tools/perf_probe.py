"""Development probe (not the bench): times the product, its stages and the reference on one GPU.
usage: python tools/perf_probe.py [n] [num_split] [--ref] [--shapes 11,21,12,22]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import ozimmu_b200 as oz  # noqa: E402


def timed(fn, iters=5, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n", type=int, nargs="?", default=4096)
    ap.add_argument("s", type=int, nargs="?", default=9)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--shapes", default="00", help="comma list: p256,p224,... = forced tile width; 00 = default")
    ap.add_argument("--zgemm", action="store_true", help="also time a complex GEMM of the same size")
    ap.add_argument("--no-extras", action="store_true", help="skip the stage profile and the cuBLAS comparisons")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--graph", action="store_true",
                    help="also time the call replayed from a CUDA graph (no host launch cost: what the GPU itself needs)")
    args = ap.parse_args()
    n, s = args.n, args.s
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)
    b = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)
    c = torch.zeros(n * n, dtype=torch.float64, device="cuda")
    h = oz.create()
    L = oz.lib()
    flop = 2.0 * n ** 3
    pairs = s * (s + 1) // 2
    for shape in args.shapes.split(","):
        # pNNN = force the tile width of the CTA-pair kernel (128, 192, 208, 224, 240, 256); anything else: default
        # h128 = 128 x 128 tiles (64 rows per CTA)
        cm, cn = (0, int(shape[1:])) if shape.startswith("p") else ((64, 128) if shape == "h128" else (0, 0))
        L.ozk_set_cluster_shape(cm, cn)
        ms = timed(lambda: oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(s)), args.iters)
        print(f"ozimmu_b200 n={n} s={s} cluster={cm}x{cn}: {ms:.3f} ms  {flop / ms / 1e9:.2f} TFLOP/s-equiv  "
              f"int8 {pairs * flop / ms / 1e12:.3f} Pop/s", flush=True)
        if args.graph:
            st = torch.cuda.Stream()
            oz.set_cuda_stream(h, st)
            with torch.cuda.stream(st):
                oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(s))   # sizes the workspace
                st.synchronize()
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=st):
                    for _ in range(10):
                        oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(s))
                ms = timed(g_.replay, args.iters) / 10
            oz.set_cuda_stream(h, None)
            print(f"ozimmu_b200 n={n} s={s} cluster={cm}x{cn} (CUDA graph of 10 calls): {ms:.4f} ms per call  "
                  f"{flop / ms / 1e9:.2f} TFLOP/s-equiv", flush=True)
    L.ozk_set_cluster_shape(0, 0)
    if args.zgemm:
        za = torch.randn(n * n, dtype=torch.complex128, device="cuda", generator=g)
        zb = torch.randn(n * n, dtype=torch.complex128, device="cuda", generator=g)
        zc = torch.zeros(n * n, dtype=torch.complex128, device="cuda")
        ms = timed(lambda: oz.gemm(h, 0, 0, n, n, n, 1.0 + 0j, za, n, zb, n, 0j, zc, n, oz.fp64_int8(s), oz.complx), args.iters)
        print(f"ozimmu_b200 ZGEMM n={n} s={s}: {ms:.3f} ms  {4 * flop / ms / 1e9:.2f} FP64-equiv TFLOP/s (8 n^3 flop)", flush=True)
        ms = timed(lambda: torch.mm(za.view(n, n), zb.view(n, n)), args.iters)
        print(f"cuBLAS ZGEMM n={n}: {ms:.3f} ms {4 * flop / ms / 1e9:.2f} TFLOP/s", flush=True)
        del za, zb, zc
    if args.no_extras:
        oz.destroy(h)
        return
    # stages
    oz.enable_profiling(h)
    for _ in range(3):
        oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(s))
    oz.print_profiler_result(h, f"n{n}_s{s}")
    oz.disable_profiling(h)
    ms = timed(lambda: torch.mm(a.view(n, n), b.view(n, n)), args.iters)
    print(f"cuBLAS DGEMM n={n}: {ms:.3f} ms {flop / ms / 1e9:.2f} TFLOP/s", flush=True)
    ai = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
    bi = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
    try:
        ms = timed(lambda: torch._int_mm(ai, bi.t()), args.iters)
        print(f"cuBLAS int8 GEMM (torch._int_mm, NT) n={n}: {ms:.3f} ms {flop / ms / 1e12:.3f} Pop/s", flush=True)
    except Exception as e:  # noqa: BLE001
        print("torch._int_mm unavailable:", e)
    if args.ref:
        from gpu_util import Reference
        ref = Reference()
        ms = timed(lambda: ref.gemm(0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, s - 1), args.iters)
        print(f"reference ozIMMU n={n} s={s}: {ms:.3f} ms  {flop / ms / 1e9:.2f} TFLOP/s-equiv", flush=True)
        ref.L.ozref_profiling(ref.h, 1)
        for _ in range(3):
            ref.gemm(0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, s - 1)
        ref.L.ozref_print_profile(ref.h, b"ref")
        ref.L.ozref_profiling(ref.h, 0)
        ref.close()
    oz.destroy(h)


if __name__ == "__main__":
    main()
