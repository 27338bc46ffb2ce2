"""-m gpu: the reference's OWN, unmodified test driver (test/main_test.cu, built by oracle/Makefile against this
repo's header and linked against libozimmu.so) runs its `ci_test` against the product: all op combinations x
{1023,1024,1025}^3 x fp64_int8_8..16 x {real, complex} = 1944 GEMMs, each passing iff the relative residual
against a double-double product is < 1e-15 (reference test/main_test.cu:703-746).  The reference library itself
cannot pass this on B200 (its int8 cublasGemmEx call fails for m = 1023/1025)."""
import re
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
DRIVER = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "main.test.ours"


def test_reference_ci_test_against_product():
    if not DRIVER.exists():
        pytest.skip("oracle/_ref/main.test.ours not built (needs the reference tree at build time)")
    p = subprocess.run([str(DRIVER), "ci_test"], capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stderr[-2000:]
    m = re.search(r"(\d+)\s*/\s*(\d+) PASSED", p.stdout)
    assert m, p.stdout[-2000:]
    assert "FAILED" not in p.stdout, [l for l in p.stdout.splitlines() if "FAILED" in l][:5]
    assert m.group(1) == m.group(2) == "1944", p.stdout[-500:]


def _write_matfile(path, mat):
    """mtk::matfile dense fp64 file (reference test/matfile/include/matfile/matfile.hpp:33-46, save_dense :232-260):
    header {u32 version = 0*1000 + 7, i32 data_t (fp64 = 9), i32 matrix_t (dense = 0), pad, u64 m, u64 n, 4 x u64}, then
    the elements column by column"""
    import struct

    import numpy as np
    m, n = mat.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<IiiI6Q", 7, 9, 0, 0, m, n, 0, 0, 0, 0))
        f.write(np.asfortranarray(mat, dtype=np.float64).tobytes(order="F"))


def _rows(stdout, header_prefix):
    """data lines after the CSV header (the driver's matfile / throughput rows carry two op columns its header does not
    name, test/main_test.cu:143-150, so fields are taken by position)"""
    lines = stdout.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(header_prefix))
    return [l.split(",") for l in lines[start + 1:] if l.count(",") >= 8]


def test_reference_driver_matfile_throughput_and_power_modes(tmp_path):
    """The remaining modes of the reference's test driver (test/main_test.cu:334-512), unmodified, against
    the product: `matfile` (operands loaded from mtk::matfile files, residual against a double-double product),
    the random-input throughput mode, and `power` (NVML-sampled GFLOP/s per watt through its gpu_monitor helper)."""
    if not DRIVER.exists():
        pytest.skip("oracle/_ref/main.test.ours not built (needs the reference tree at build time)")
    import numpy as np
    rng = np.random.default_rng(0)
    m, k, n = 1024, 1280, 1152
    _write_matfile(tmp_path / "A.matrix", rng.standard_normal((m, k)))
    _write_matfile(tmp_path / "B.matrix", rng.standard_normal((k, n)))
    p = subprocess.run([str(DRIVER), "matfile", str(tmp_path / "A.matrix"), str(tmp_path / "B.matrix"), "fp64_int8_9",
                        "fp64_int8_12", "dgemm"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    rows = _rows(p.stdout, "gpu,gemm,input,mode")
    # gpu, D|Z, input, mode, opA, opB, m, n, k, residual, max_relative, throughput
    assert [r[3] for r in rows] == ["fp64_int8_9", "fp64_int8_12", "dgemm"], p.stdout[-1500:]
    for r in rows:
        assert (int(r[6]), int(r[7]), int(r[8])) == (m, n, k)
        assert float(r[9]) < {"fp64_int8_9": 1e-12, "fp64_int8_12": 1e-14, "dgemm": 1e-14}[r[3]], r
    # throughput mode: exp_rand-1 inputs, DGEMM and ZGEMM, 2048
    for gemm in ("dgemm", "zgemm"):
        p = subprocess.run([str(DRIVER), "exp_rand-1", gemm, "seq", "2048", "2048", "1", "fp64_int8_13"],
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        rows = _rows(p.stdout, "gpu,gemm,input,mode")
        assert len(rows) == 1 and float(rows[0][9]) < 1e-14 and float(rows[0][11]) > 1.0, p.stdout
    # power mode: gpu, mode, m, n, k, throughput_in_tflops, avg_watt, gflops_per_watt, time, count
    p = subprocess.run([str(DRIVER), "power", "seq", "4096", "4096", "1", "fp64_int8_9"], capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    rows = _rows(p.stdout, "gpu,mode,m,n,k")
    assert len(rows) == 1 and rows[0][1] == "fp64_int8_9", p.stdout[-1500:]
    assert float(rows[0][5]) > 10.0 and float(rows[0][6]) > 50.0, rows[0]
    print("power mode:", rows[0])
