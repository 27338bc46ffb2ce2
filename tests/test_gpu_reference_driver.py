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
