"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/ozimmu_b200.h
declares, plus the cuBLAS entry points it interposes; host-only entry points behave like the reference's."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

import ozimmu_b200 as oz
from ozimmu_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "ozimmu_b200.h"


def declared_functions():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oz(?:k|immu)_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("ozk_split_int8", "ozk_gemm_i8_fused", "ozk_mantissa_loss", "ozimmu_create", "ozimmu_gemm",
                 "ozimmu_gemm_host", "ozimmu_auto_mode_select"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = oz.lib()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ozimmu_b200.h but not exported"
    assert sorted(declared_functions()) == sorted(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"


def test_library_exports_cublas_interposers():
    out = subprocess.run(["nm", "-D", "--defined-only", str(oz.LIB_PATH)], capture_output=True, text=True, check=True)
    defined = {line.split()[-1] for line in out.stdout.splitlines() if " T " in line}
    for name in _lib.INTERPOSED:
        assert name in defined
    # the C++ API of the reference header is exported too (namespace mtk::ozimmu)
    mangled = subprocess.run(["nm", "-DC", "--defined-only", str(oz.LIB_PATH)], capture_output=True, text=True).stdout
    for fn in ("mtk::ozimmu::create", "mtk::ozimmu::destroy", "mtk::ozimmu::gemm", "mtk::ozimmu::auto_mode_select",
               "mtk::ozimmu::reallocate_working_memory", "mtk::ozimmu::get_compute_mode_name_str",
               "mtk::ozimmu::get_auto_mantissa_loss_threashold", "mtk::ozimmu::set_cuda_stream"):
        assert re.search(re.escape(fn) + r"(\[abi:cxx11\])?\(", mangled), fn


def test_kernels_are_blackwell_native():
    """SASS of the shipped library must contain tcgen05 MMA (UTCIMMA), TMEM loads (LDTM) and bulk async copies
    (UBLKCP: the operand tiles are staged with linear cp.async.bulk, not tensor-map TMA)."""
    out = subprocess.run(["cuobjdump", "-sass", str(oz.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    for mnemonic in ("UTCIMMA", "LDTM", "UBLKCP"):
        assert mnemonic in out.stdout, mnemonic


def test_mode_names_and_sizes():
    assert oz.get_compute_mode_name_str(oz.compute_mode_t.dgemm) == "dgemm"
    assert oz.get_compute_mode_name_str(oz.compute_mode_t.sgemm) == "sgemm"
    assert oz.get_compute_mode_name_str(oz.compute_mode_t.fp64_int8_auto) == "fp64_int8_auto"
    for s in range(3, 19):
        assert oz.get_compute_mode_name_str(oz.fp64_int8(s)) == f"fp64_int8_{s}"
        assert oz.num_split_of(oz.fp64_int8(s)) == s
    assert oz.lib().ozimmu_get_compute_mode_name_str(77) is None
    assert int(oz.lib().ozk_slice_pitch(1)) == 128 and int(oz.lib().ozk_slice_pitch(4097)) == 4224
    # blocked slice layout: rows padded to 256, k to 128
    assert int(oz.lib().ozk_slices_bytes(1, 1, 9)) == 9 * 256 * 128
    assert int(oz.lib().ozk_slices_bytes(8192, 8192, 9)) == 9 * 8192 * 8192


def test_invalid_arguments_rejected_without_gpu():
    lib = oz.lib()
    one = C.c_double(1.0)
    # null handle / bad mode -> 1 (invalid argument), never a crash
    assert lib.ozimmu_gemm(None, 0, 0, 8, 8, 8, C.addressof(one), None, 8, None, 8, C.addressof(one), None, 8, 8, 0) == 1
    assert lib.ozimmu_gemm_host(None, 0, 0, 8, 8, 8, C.addressof(one), None, 8, None, 8, C.addressof(one), None, 8, 8) == 1
    # kernel launchers validate before touching the device
    assert lib.ozk_split_int8(None, 17, None, None, 4, 4, None, 4, 0, 9, 7, None) != 0   # pitch % 16
    assert lib.ozk_split_int8(None, 128, None, None, 4, 4, None, 4, 0, 2, 7, None) != 0  # num_split < 3
    assert lib.ozk_gemm_i8_fused(8, 8, 8, None, None, 128, None, None, 19, 7, 1.0, 0.0, None, 8, None) != 0
    assert lib.ozk_gemm_i8_fused(8, 8, 8, None, None, 128, None, None, 9, 7, 1.0, 0.0, None, 4, None) != 0  # ldc < m
    assert lib.ozk_gemm_i8_fused(8, 8, 8, None, None, 16, None, None, 9, 7, 1.0, 0.0, None, 8, None) != 0   # pitch % 128
    assert lib.ozk_gemm_i8_fused(0, 8, 8, None, None, 128, None, None, 9, 7, 1.0, 0.0, None, 1, None) == 0  # empty
    assert lib.ozimmu_launch_count() == 0


@pytest.mark.parametrize("taper", [0, 1])
@pytest.mark.parametrize("want", [0, 1, 256, 300, 512, 768, 1024, 4096, 100000])
@pytest.mark.parametrize("extent", [1, 100, 256, 257, 700, 1300, 4362, 8192, 16384, 40000, 131072])
def test_host_block_edges(extent, want, taper):
    """Block schedule of the host-operand pipeline (csrc/host_e2e.cu): a partition of [0, extent) into at most 16
    blocks whose inner boundaries sit on the kernel's 256-row tiles (the block-wise split requires it)."""
    lib = oz.lib()
    buf = (C.c_size_t * 32)()
    cnt = lib.ozimmu_host_block_edges(extent, want, taper, C.addressof(buf), 32)
    e = list(buf[:cnt])
    assert 2 <= cnt <= 17
    assert e[0] == 0 and e[-1] == extent
    assert all(a < b for a, b in zip(e, e[1:]))
    assert all(x % 256 == 0 for x in e[1:-1])
    if want == 0:
        assert e == [0, extent]
    if taper and want and cnt > 3:
        sizes = [b - a for a, b in zip(e, e[1:])]
        assert sizes[-1] <= sizes[0] and sizes[-2] <= sizes[0]     # the blocks that arrive last are not the big ones




def test_fused_tile_choice():
    """host logic of the product dispatch (no GPU): the measured cost table picks 256 x 256 tiles for the BASELINE sizes,
    128 x 128 tiles (64 rows per CTA) where a problem would otherwise leave SMs idle, 128-wide tiles at small k, and the
    SM-time criterion for the one-tile launches of the block pipelines"""
    import ctypes as C
    L = oz.lib()

    def choice(m, n, k, batch=1, one_tile=0, sms=148):
        rows, width = C.c_int(), C.c_int()
        assert L.ozk_fused_tile_choice(m, n, k, batch, one_tile, sms, C.addressof(rows), C.addressof(width)) == 0
        assert (rows.value, width.value) in {(256, w) for w in (256, 240, 224, 208, 192, 128)} | {(128, 128)}
        return rows.value, width.value

    assert choice(8192, 8192, 8192) == (256, 256)            # headline
    assert choice(2048, 16384, 16384) == (256, 256)          # config 4, one rank of eight
    assert choice(16384, 16384, 16384) == (256, 256)
    assert choice(1024, 1024, 1024) == (128, 128)            # 64 tiles on 128 SMs instead of 32 on 64
    assert choice(300, 200, 520) == (128, 128)
    assert choice(1536, 1536, 1536) == (256, 128)            # one round of 72 tiles; small k favours the narrow tile
    assert choice(2048, 2048, 2048) == (256, 128)
    assert choice(768, 768, 8192, one_tile=1) == (256, 256)  # block of the host-operand pipeline: SM time per area
    assert choice(64, 64, 64, batch=4096)[0] == 128          # a big batch of tiny GEMMs: the cheapest single tile
    assert L.ozk_fused_tile_choice(8, 8, 8, 1, 0, 148, None, None) == 1
