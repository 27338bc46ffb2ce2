"""Generates the golden vectors that pin oracle/oz_oracle.c to the UNMODIFIED reference.

Run on a GPU box (the reference has no CPU path):
    python tests/golden/make_golden.py gpurun_out/golden
and copy the resulting *.npz into tests/golden/.  The reference is oracle/_ref/libozref.so, built by
oracle/Makefile from the sources under /root/reference (never copied).  Inputs are small seeded
matrices stored in the fixture itself, so the CPU test needs neither a GPU nor the reference.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import oracle_lib  # noqa: E402
from gpu_util import Reference, to_dev  # noqa: E402

# (op_a, op_b, m, n, k, num_split, kind, alpha, beta, ld_extra)
GEMM_CASES = [
    # m is a multiple of 4 everywhere: cuBLAS 12.9 on sm_100 rejects the reference's int8 GemmEx
    # (src/gemm.cu:322, ldc = m) otherwise -- the reference's own ci_test sizes 1023/1025 fail on B200.
    (0, 0, 24, 20, 40, 9, "urand01", 1.0, 0.0, 0),
    (1, 0, 16, 23, 33, 13, "exp_rand-2", -1.5, 0.75, 3),
    (0, 1, 32, 9, 64, 3, "normal01", 2.0, 0.0, 1),
    (1, 1, 20, 27, 50, 18, "mixed", 1.0, -1.0, 0),
    (0, 0, 4, 7, 131, 6, "exp_rand-1", 0.5, 2.0, 2),
    (0, 0, 16, 15, 300, 10, "exp_rand-4", 1.0, 0.0, 0),
]
# auto mode: no exact zeros and k % 32 == 0, the only regime where the reference's counters are
# reliable; only the first 8 counters exist in the reference (App. B.1)
AUTO_CASES = [(0, 0, 32, 32, 64, 0.0), (1, 0, 32, 48, 96, 1.0), (0, 1, 64, 32, 32, 2.0), (1, 1, 32, 32, 128, 4.0)]
# complex (zgemm) cases: (op_a, op_b, m, n, k, num_split, kind, alpha, beta)
ZGEMM_CASES = [
    (0, 0, 16, 12, 40, 9, "normal01", 1.0 + 0.0j, 0.0 + 0.0j),
    (1, 0, 12, 9, 33, 13, "exp_rand-1", -0.5 + 1.25j, 0.75 - 0.5j),
    (0, 1, 8, 15, 64, 4, "urand01", 2.0 - 1.0j, 0.0 + 1.0j),
    (1, 1, 20, 7, 50, 18, "mixed", 0.0 + 1.0j, 1.0 + 0.0j),
]
AUTO_THRESHOLDS = [0.0, 0.5, 1.0, 1.5, 2.0, 4.0, 8.0, 30.0]


def stored(op, rows, cols, extra):
    r, c = (rows, cols) if op == 0 else (cols, rows)
    return r + extra, c


def main(out_dir: str) -> None:
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    ref = Reference()
    for idx, (op_a, op_b, m, n, k, s, kind, alpha, beta, extra) in enumerate(GEMM_CASES):
        lda, ca = stored(op_a, m, k, extra)
        ldb, cb = stored(op_b, k, n, extra)
        ldc = m + extra
        a = oracle_lib.gen_matrix(kind, lda * ca, 1000 + idx)
        b = oracle_lib.gen_matrix(kind, ldb * cb, 2000 + idx)
        c = oracle_lib.gen_matrix("normal01", ldc * n, 3000 + idx)
        da, db, dc = to_dev(a), to_dev(b), to_dev(c)
        ref.gemm(op_a, op_b, m, n, k, alpha, da, lda, db, ldb, beta, dc, ldc, s - 1)
        torch.cuda.synchronize()
        bits = int(ref.L.ozref_bits_per_int8(k))
        # reference split_int8<double>(out, ldo, max, m, n, in, ld, op, matrix, ...):
        #   A: (m, k, op_a, matrix_A)   B: (k, n, op_b, matrix_B)
        a_sl, amax = ref.split(da, lda, m, k, op_a, 0, s, bits)
        b_sl, bmax = ref.split(db, ldb, k, n, op_b, 1, s, bits)
        np.savez_compressed(out / f"gemm_{idx}.npz", op_a=op_a, op_b=op_b, m=m, n=n, k=k, num_split=s, alpha=alpha,
                            beta=beta, lda=lda, ldb=ldb, ldc=ldc, a=a, b=b, c_in=c, c_out=dc.cpu().numpy(),
                            a_slices=a_sl.cpu().numpy(), amax=amax.cpu().numpy(), b_slices=b_sl.cpu().numpy(),
                            bmax=bmax.cpu().numpy(), bits=bits, kind=kind)
        print("gemm case", idx, "ok")
    for idx, (op_a, op_b, m, n, k, phi) in enumerate(AUTO_CASES):
        lda, ca = stored(op_a, m, k, 0)
        ldb, cb = stored(op_b, k, n, 0)
        a = oracle_lib.gen_matrix(f"exp_rand-{phi}", lda * ca, 4000 + idx)
        b = oracle_lib.gen_matrix(f"exp_rand-{phi}", ldb * cb, 5000 + idx)
        assert (a != 0).all() and (b != 0).all()
        da, db = to_dev(a), to_dev(b)
        modes, counters = [], None
        for thr in AUTO_THRESHOLDS:
            mode, cnt = ref.auto_mode_select(op_a, op_b, m, n, k, da, lda, db, ldb, thr)
            modes.append(mode)
            counters = cnt
        np.savez_compressed(out / f"auto_{idx}.npz", op_a=op_a, op_b=op_b, m=m, n=n, k=k, lda=lda, ldb=ldb, a=a, b=b,
                            thresholds=np.array(AUTO_THRESHOLDS), modes=np.array(modes),
                            counters8=np.array(counters, dtype=np.uint64))
        print("auto case", idx, "ok", modes)
    for idx, (op_a, op_b, m, n, k, s, kind, alpha, beta) in enumerate(ZGEMM_CASES):
        lda, ca = stored(op_a, m, k, 0)
        ldb, cb = stored(op_b, k, n, 0)
        a = oracle_lib.gen_complex(kind, lda * ca, 6000 + idx)
        b = oracle_lib.gen_complex(kind, ldb * cb, 7000 + idx)
        c = oracle_lib.gen_complex("normal01", m * n, 8000 + idx)
        da, db, dc = to_dev(a), to_dev(b), to_dev(c)
        ref.gemm_complex(op_a, op_b, m, n, k, alpha, da, lda, db, ldb, beta, dc, m, s - 1)
        torch.cuda.synchronize()
        np.savez_compressed(out / f"zgemm_{idx}.npz", op_a=op_a, op_b=op_b, m=m, n=n, k=k, num_split=s,
                            alpha=np.complex128(alpha), beta=np.complex128(beta), lda=lda, ldb=ldb, ldc=m, a=a, b=b,
                            c_in=c, c_out=dc.cpu().numpy(), kind=kind)
        print("zgemm case", idx, "ok")
    ref.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
