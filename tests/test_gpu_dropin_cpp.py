"""-m gpu: a plain C++ cuBLAS application (tests/dropin_app.cu, compiled here with nvcc, dynamically linked to
libcublas) under LD_PRELOAD=libozimmu.so: its cublasDgemm must be served by the Ozaki path (bit-identical to a
direct ozimmu_gemm call, [ozIMMU LOG] lines present) and must be untouched without OZIMMU_COMPUTE_MODE."""
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle_lib
import ozimmu_b200 as oz
from gpu_util import bits, to_dev

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent


def test_cpp_app_under_ld_preload(tmp_path, handle):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available on this box")
    app = tmp_path / "dropin_app"
    b = subprocess.run(["nvcc", "-O2", "-o", str(app), str(HERE / "dropin_app.cu"), "-lcublas"], capture_output=True,
                       text=True, timeout=600)
    assert b.returncode == 0, b.stderr[-2000:]
    n = 1280
    a = oracle_lib.gen_matrix("exp_rand-1", n * n, 1)
    bm = oracle_lib.gen_matrix("exp_rand-1", n * n, 2)
    c = oracle_lib.gen_matrix("normal01", n * n, 3)
    np.concatenate([a, bm, c]).tofile(tmp_path / "in.bin")

    def run(extra_env, out, mode="host"):
        env = dict(os.environ, **extra_env)
        p = subprocess.run([str(app), str(n), str(tmp_path / "in.bin"), str(tmp_path / out), mode], env=env,
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, (p.returncode, p.stdout[-1000:], p.stderr[-1000:])
        return p.stdout, np.fromfile(tmp_path / out)

    log, c_oz = run({"LD_PRELOAD": str(oz.LIB_PATH), "OZIMMU_COMPUTE_MODE": "fp64_int8_12", "OZIMMU_INFO": "1"}, "oz.bin")
    assert "[ozIMMU LOG]" in log
    _, c_pass = run({"LD_PRELOAD": str(oz.LIB_PATH)}, "pass.bin")          # mode unset -> passthrough
    _, c_plain = run({}, "plain.bin")                                       # no preload at all
    assert np.array_equal(bits(c_pass), bits(c_plain)), "passthrough must not change cuBLAS results"
    dc = to_dev(c)
    assert oz.gemm(handle, 0, 1, n, n, n, 1.5, to_dev(a), n, to_dev(bm), n, -0.5, dc, n, oz.fp64_int8(12)) == 0
    torch.cuda.synchronize()
    assert np.array_equal(bits(c_oz), bits(dc)), "intercepted cublasDgemm differs from the direct call"
    assert not np.array_equal(bits(c_oz), bits(c_plain))                    # it really took the other path
    assert np.linalg.norm(c_oz - c_plain) / np.linalg.norm(c_plain) < 1e-14
    # cuBLAS device pointer mode: alpha / beta are read on the device by the Ozaki kernels (the reference dereferences the
    # device pointers on the host, src/gemm.cu:405) -- same bits as the host-pointer call
    oz_env = {"LD_PRELOAD": str(oz.LIB_PATH), "OZIMMU_COMPUTE_MODE": "fp64_int8_12", "OZIMMU_INFO": "1"}
    log, c_dev = run(oz_env, "devptr.bin", "devptr")
    assert "[ozIMMU LOG]" in log
    assert np.array_equal(bits(c_dev), bits(dc)), "device-pointer-mode cublasDgemm differs from the direct call"
    # two host threads, each with its own cuBLAS handle and stream, three DGEMMs each, concurrently: per-device state
    # behind one lock, the shared workspace ordered across the two streams
    _, c_thr = run(oz_env, "threads.bin", "threads")
    assert np.array_equal(bits(c_thr[:n * n]), bits(dc)) and np.array_equal(bits(c_thr[n * n:]), bits(dc))
    # one process driving two GPUs, one cuBLAS handle per GPU (needs 2 GPUs): every GPU gets its own ozIMMU handle and
    # workspace (the reference has one process-global handle, src/cublas.cu:58-86)
    if torch.cuda.device_count() >= 2:
        _, c_two = run(oz_env, "multigpu.bin", "multigpu")
        assert np.array_equal(bits(c_two[:n * n]), bits(dc)) and np.array_equal(bits(c_two[n * n:]), bits(dc))
