"""BASELINE configs 3 and 5 on one GPU (results -> profiles/):
  split : 8192^3, fp64_int8_{3..18}: FP64-equiv TFLOP/s and max relative error vs cuBLAS DGEMM, ours and reference
  auto  : 4096^3 fp64_int8_auto on exp_rand-phi inputs: loss counters, selected mode per threshold, error, TFLOP/s
usage: python tools/sweeps.py split|auto [n]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import ozimmu_b200 as oz  # noqa: E402
import oracle_lib  # noqa: E402


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


def gen(kind, count, g):
    if kind == "urand01":
        return 1.0 - torch.rand(count, dtype=torch.float64, device="cuda", generator=g)
    if kind == "normal01":
        return torch.randn(count, dtype=torch.float64, device="cuda", generator=g)
    phi = float(kind.split("-", 1)[1])
    return (torch.rand(count, dtype=torch.float64, device="cuda", generator=g) - 0.5) * torch.exp(
        phi * torch.randn(count, dtype=torch.float64, device="cuda", generator=g))


def errors(c, ref):
    d = (c - ref).abs()
    return (d / ref.abs().clamp_min(1e-300)).max().item(), (torch.linalg.vector_norm(d) / torch.linalg.vector_norm(ref)).item()


def split_sweep(n):
    from gpu_util import Reference
    ref_lib = Reference() if oracle_lib.reference() is not None else None
    h = oz.create()
    flop = 2.0 * n ** 3
    print("input,mode,ours_tflops,ref_tflops,speedup,max_rel_err_vs_dgemm,rel_residual_vs_dgemm,bit_identical_to_reference")
    for kind in ("urand01", "exp_rand-1"):
        g = torch.Generator(device="cuda").manual_seed(0)
        a, b = gen(kind, n * n, g), gen(kind, n * n, g)
        c = torch.zeros(n * n, dtype=torch.float64, device="cuda")
        cr = torch.zeros_like(c)
        dg = (b.view(n, n) @ a.view(n, n)).reshape(-1)  # cuBLAS DGEMM of the same column-major operands
        for s in range(3, 19):
            ms = timed(lambda: oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.fp64_int8(s)), 3, 1)
            mx, rr = errors(c, dg)
            if ref_lib is not None:
                msr = timed(lambda: ref_lib.gemm(0, 0, n, n, n, 1.0, a, n, b, n, 0.0, cr, n, s - 1), 3, 1)
                same = bool(torch.equal(c.view(torch.int64), cr.view(torch.int64)))
                print(f"{kind},fp64_int8_{s},{flop / ms / 1e9:.2f},{flop / msr / 1e9:.2f},{msr / ms:.2f},{mx:.3e},{rr:.3e},{same}", flush=True)
            else:
                print(f"{kind},fp64_int8_{s},{flop / ms / 1e9:.2f},,,{mx:.3e},{rr:.3e},", flush=True)
        ms = timed(lambda: torch.mm(b.view(n, n), a.view(n, n)), 3, 1)
        print(f"{kind},cublas_dgemm,{flop / ms / 1e9:.2f},,,0,0,", flush=True)
    oz.destroy(h)


def auto_sweep(n):
    from gpu_util import Reference
    ref_lib = Reference() if oracle_lib.reference() is not None else None
    h = oz.create()
    flop = 2.0 * n ** 3
    print("phi,threshold,selected_mode,avg_loss_at_selected,max_rel_err_vs_dgemm,rel_residual,tflops,"
          "ref_selected_mode,counters_3..10_equal_reference,selection_equal_reference,counters_3..18")
    for phi in (0.0, 0.5, 1.0, 2.0, 4.0, 8.0):
        g = torch.Generator(device="cuda").manual_seed(1)
        a, b = gen(f"exp_rand-{phi}", n * n, g), gen(f"exp_rand-{phi}", n * n, g)
        c = torch.zeros(n * n, dtype=torch.float64, device="cuda")
        dg = (b.view(n, n) @ a.view(n, n)).reshape(-1)
        for thr in (0.0, 0.5, 1.0, 1.5, 2.0, 4.0, 8.0):
            cnt = []
            mode = oz.auto_mode_select(h, 0, 0, n, n, n, a, n, b, n, oz.real, thr, cnt)
            oz.set_auto_mantissa_loss_threashold(h, thr)
            ms = timed(lambda: oz.gemm(h, 0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, oz.compute_mode_t.fp64_int8_auto), 3, 1)
            mx, rr = errors(c, dg)
            name = oz.get_compute_mode_name_str(mode)
            avg = cnt[int(mode) - 2] / (2.0 * n * n) if name != "dgemm" else float("nan")
            ref_name, cnt_same, sel_same = "", "", ""
            if ref_lib is not None:
                # the reference keeps 8 counters (fp64_int8_3..10): beyond them its selection is undefined
                ref_mode, ref_cnt = ref_lib.auto_mode_select(0, 0, n, n, n, a, n, b, n, thr)
                in_range = int(oz.compute_mode_t.fp64_int8_3) <= ref_mode <= int(oz.compute_mode_t.fp64_int8_10)
                ref_name = oz.get_compute_mode_name_str(oz.compute_mode_t(ref_mode)) if 0 <= ref_mode <= 18 else str(ref_mode)
                cnt_same = cnt[:8] == ref_cnt
                sel_same = (int(mode) == ref_mode) if in_range else "n/a (reference has no counter beyond fp64_int8_10)"
            print(f"{phi},{thr},{name},{avg:.4f},{mx:.3e},{rr:.3e},{flop / ms / 1e9:.2f},{ref_name},{cnt_same},{sel_same},"
                  f"{' '.join(str(v) for v in cnt)}", flush=True)
    oz.destroy(h)
    if ref_lib is not None:
        ref_lib.close()


if __name__ == "__main__":
    which = sys.argv[1]
    if which == "split":
        split_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 8192)
    else:
        auto_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
