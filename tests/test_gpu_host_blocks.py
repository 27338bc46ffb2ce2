"""-m gpu: the block-wise entries behind the host-operand pipeline (ozimmu_gemm_host cuts C into row blocks of
op(A) x column blocks of op(B) and computes each block as soon as its operands have arrived over PCIe).
Blocks must be bit-identical to the whole-matrix calls: the split scales every row of A / column of B on its
own (reference src/split.cu:193-242,277-282) and an element of C depends only on its row of A and column of B."""
import os

import numpy as np
import pytest
import torch

import oracle_lib
import ozimmu_b200 as oz
from gpu_util import bits, stream_ptr, to_dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    h = oz.create()
    yield h
    oz.destroy(h)


def _split_whole(L, x, ld, rows, length, col_major, s, nbits):
    pitch = int(L.ozk_slice_pitch(length))
    out = torch.full((int(L.ozk_slices_bytes(rows, length, s)),), 77, dtype=torch.int8, device="cuda")
    mx = torch.full((rows,), -1.0, dtype=torch.float64, device="cuda")
    scr = torch.zeros(rows, dtype=torch.int32, device="cuda")
    assert L.ozk_split_int8(out.data_ptr(), pitch, mx.data_ptr(), scr.data_ptr(), rows, length, x.data_ptr(), ld,
                            int(col_major), s, nbits, stream_ptr()) == 0
    torch.cuda.synchronize()
    return out, mx, pitch


@pytest.mark.parametrize("col_major", [False, True])
@pytest.mark.parametrize("rows,length,edges", [
    (1024, 520, [0, 256, 768, 1024]),      # plane ends on a tile boundary
    (1100, 300, [0, 512, 1024, 1100]),     # ragged last block: it clears the plane's padding rows
    (700, 1000, [0, 700]),                 # one block == the whole-matrix call
])
def test_split_blocks_equal_whole(rows, length, edges, col_major):
    L = oz.lib()
    s, nbits = 9, 7
    ld = (rows if col_major else length) + 3
    x = to_dev(oracle_lib.gen_matrix("exp_rand-1", ld * (length if col_major else rows), 5))
    want, want_mx, pitch = _split_whole(L, x, ld, rows, length, col_major, s, nbits)
    got = torch.full_like(want, 55)
    mx = torch.full((rows,), -1.0, dtype=torch.float64, device="cuda")
    scr = torch.zeros(rows, dtype=torch.int32, device="cuda")
    # blocks in reverse order: nothing may depend on the order of arrival
    for r0, r1 in reversed(list(zip(edges[:-1], edges[1:]))):
        src = x.data_ptr() + 8 * (r0 if col_major else r0 * ld)
        rc = L.ozk_split_int8_block(got.data_ptr(), pitch, rows, r0, mx.data_ptr() + 8 * r0, scr.data_ptr() + 4 * r0,
                                    r1 - r0, length, src, ld, int(col_major), s, nbits, 1, stream_ptr())
        assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(mx.view(torch.int64), want_mx.view(torch.int64))
    assert torch.equal(got, want)


def test_split_block_rejects_misaligned_blocks():
    L = oz.lib()
    x = torch.zeros(1024 * 128, dtype=torch.float64, device="cuda")
    out = torch.zeros(int(L.ozk_slices_bytes(1024, 128, 3)), dtype=torch.int8, device="cuda")
    mx = torch.zeros(1024, dtype=torch.float64, device="cuda")
    scr = torch.zeros(1024, dtype=torch.int32, device="cuda")
    args = lambda r0, n: (out.data_ptr(), 128, 1024, r0, mx.data_ptr(), scr.data_ptr(), n, 128, x.data_ptr(), 128, 0, 3,
                          7, 1, stream_ptr())
    assert L.ozk_split_int8_block(*args(128, 256)) != 0      # row0 not a multiple of 256
    assert L.ozk_split_int8_block(*args(0, 300)) != 0        # ends inside a tile, not at the plane's end
    assert L.ozk_split_int8_block(*args(768, 512)) != 0      # past the plane
    assert L.ozk_split_int8_block(*args(768, 256)) == 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("m,n,k,row_edges,col_edges", [
    (1024, 1280, 600, [0, 512, 1024], [0, 256, 1280]),
    (900, 1100, 1030, [0, 768, 900], [0, 512, 1024, 1100]),
])
@pytest.mark.parametrize("beta", [0.0, -0.75])
@pytest.mark.parametrize("width", [0, 240, 224, 208, 192, 128, -128])
def test_fused_blocks_equal_whole(m, n, k, row_edges, col_edges, beta, width):
    """block launches (all flag combinations, every tile width -- widths below 256 read B rows past the block and, at
    the plane's end, must not read past the plane; -128 = the 128 x 128 tile with 64 rows per CTA) == one whole launch"""
    L = oz.lib()
    s, nbits = 9, int(L.ozk_bits_per_int8(k))
    a = to_dev(oracle_lib.gen_matrix("exp_rand-1", m * k, 31))   # op_t A: k x m column-major == rows contiguous
    b = to_dev(oracle_lib.gen_matrix("exp_rand-1", k * n, 32))   # op_n B: k x n column-major
    c0 = oracle_lib.gen_matrix("normal01", (m + 5) * n, 33)
    ldc = m + 5
    a_sl, amax, pitch = _split_whole(L, a, k, m, k, False, s, nbits)
    b_sl, bmax, _ = _split_whole(L, b, k, n, k, False, s, nbits)
    want = to_dev(c0)
    assert L.ozk_gemm_i8_fused(m, n, k, a_sl.data_ptr(), b_sl.data_ptr(), pitch, amax.data_ptr(), bmax.data_ptr(), s,
                               nbits, 1.25, beta, want.data_ptr(), ldc, stream_ptr()) == 0
    got = to_dev(c0)
    streams = [torch.cuda.Stream() for _ in range(3)]
    torch.cuda.synchronize()
    i = 0
    L.ozk_set_cluster_shape(*((64, 128) if width < 0 else (0, width)))
    for r0, r1 in zip(row_edges[:-1], row_edges[1:]):
        for q0, q1 in zip(col_edges[:-1], col_edges[1:]):
            st = streams[i % 3]   # concurrent launches on several streams, as the host pipeline issues them
            i += 1
            rc = L.ozk_gemm_i8_fused_block(r1 - r0, q1 - q0, k, a_sl.data_ptr(), m, r0, b_sl.data_ptr(), n, q0, pitch,
                                           amax.data_ptr() + 8 * r0, bmax.data_ptr() + 8 * q0, s, nbits, 1.25, beta,
                                           got.data_ptr() + 8 * (q0 * ldc + r0), ldc, i & 3, int(st.cuda_stream))
            assert rc == 0
    torch.cuda.synchronize()
    L.ozk_set_cluster_shape(0, 0)
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))


@pytest.mark.parametrize("panel,rowblock,one_tile,rect_tiles,a_first,tail_pieces", [
    ("256", "256", None, None, None, None), ("512", "0", "0", None, "0", "1"), ("0", "512", None, "2", None, "3"),
    ("768", "1024", "0", "3", "0", None), ("512", "512", "1", "1", None, "8"), ("256", "512", None, None, "0", "2"),
    (None, None, None, None, None, None)])
def test_gemm_host_block_schedules(handle, panel, rowblock, one_tile, rect_tiles, a_first, tail_pieces, monkeypatch):
    """Every block schedule of the host-operand entry (square blocks, column panels only, row blocks only, the
    default; persistent or one-tile-per-pair launches; rectangles cut into several launches; A or B travelling first;
    the last rectangles cut into pieces) gives the bits of the device entry; ragged sizes, padded leading dimensions,
    all op combinations."""
    for name, v in (("OZIMMU_B200_E2E_PANEL", panel), ("OZIMMU_B200_E2E_ROWBLOCK", rowblock),
                    ("OZIMMU_B200_E2E_ONE_TILE", one_tile), ("OZIMMU_B200_E2E_RECT_TILES", rect_tiles),
                    ("OZIMMU_B200_E2E_A_FIRST", a_first), ("OZIMMU_B200_E2E_TAIL_PIECES", tail_pieces)):
        if v is None:
            monkeypatch.delenv(name, raising=False)
        else:
            monkeypatch.setenv(name, v)
    m, n, k, pad = 1300, 1500, 700, 2
    for op_a, op_b, beta in [(0, 0, 0.0), (0, 1, -0.5), (1, 0, 2.0), (1, 1, 0.0)]:
        lda = (m if op_a == 0 else k) + pad
        ldb = (k if op_b == 0 else n) + pad
        ldc = m + pad
        a = oracle_lib.gen_matrix("exp_rand-1", lda * (k if op_a == 0 else m), 41)
        b = oracle_lib.gen_matrix("exp_rand-1", ldb * (n if op_b == 0 else k), 42)
        c = oracle_lib.gen_matrix("normal01", ldc * n, 43)
        dc = to_dev(c)
        assert oz.gemm(handle, op_a, op_b, m, n, k, 1.5, to_dev(a), lda, to_dev(b), ldb, beta, dc, ldc,
                       oz.fp64_int8(8)) == 0
        torch.cuda.synchronize()
        hc = torch.from_numpy(c.copy()).pin_memory()
        assert oz.gemm_host(handle, op_a, op_b, m, n, k, 1.5, torch.from_numpy(a).pin_memory(), lda,
                            torch.from_numpy(b).pin_memory(), ldb, beta, hc, ldc, oz.fp64_int8(8)) == 0
        got, want = hc.numpy().reshape(n, ldc)[:, :m], dc.cpu().numpy().reshape(n, ldc)[:, :m]
        assert np.array_equal(bits(got), bits(want)), (op_a, op_b, beta)
        # the padding rows of C (ld > m) are not part of the matrix: the entry must leave the host copy alone
        assert np.array_equal(bits(hc.numpy().reshape(n, ldc)[:, m:]), bits(c.reshape(n, ldc)[:, m:]))


def test_gemm_host_many_blocks_is_capped(handle, monkeypatch):
    """more blocks than the pipeline has events for: the block edge grows instead"""
    monkeypatch.setenv("OZIMMU_B200_E2E_PANEL", "256")
    monkeypatch.setenv("OZIMMU_B200_E2E_ROWBLOCK", "256")
    m, n, k = 256 * 17 + 10, 256 * 18, 256
    a = oracle_lib.gen_matrix("urand01", m * k, 51)
    b = oracle_lib.gen_matrix("urand01", k * n, 52)
    dc = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    assert oz.gemm(handle, 0, 0, m, n, k, 1.0, to_dev(a), m, to_dev(b), k, 0.0, dc, m, oz.fp64_int8(6)) == 0
    torch.cuda.synchronize()
    hc = np.zeros(m * n)
    assert oz.gemm_host(handle, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, hc, m, oz.fp64_int8(6)) == 0
    assert np.array_equal(bits(hc), bits(dc))


@pytest.mark.parametrize("op_a,op_b,beta", [(0, 0, 0.0), (1, 0, -1.5), (0, 1, 0.5)])
def test_gemm_streamed_b_equals_gemm(handle, op_a, op_b, beta):
    """B filled panel by panel on a side stream (as the multi-GPU broadcast does), each panel guarded by an event:
    the result is the plain call's, bit for bit -- also when the panels are completed in reverse order."""
    m, n, k = 700, 1900, 900
    lda = m if op_a == 0 else k
    ldb = k if op_b == 0 else n
    a = to_dev(oracle_lib.gen_matrix("exp_rand-1", m * k, 61))
    b_full = to_dev(oracle_lib.gen_matrix("exp_rand-1", k * n, 62))
    c0 = oracle_lib.gen_matrix("normal01", m * n, 63)
    want = to_dev(c0)
    assert oz.gemm(handle, op_a, op_b, m, n, k, 2.0, a, lda, b_full, ldb, beta, want, m, oz.fp64_int8(10)) == 0
    torch.cuda.synchronize()
    edges = [0, 512, 1280, n]
    b = torch.zeros_like(b_full)
    got = to_dev(c0)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    events = [torch.cuda.Event() for _ in range(3)]
    bv, bfv = (b.view(n, k), b_full.view(n, k)) if op_b == 0 else (b.view(k, n), b_full.view(k, n))
    with torch.cuda.stream(side):
        for p in (2, 1, 0):
            torch.cuda._sleep(2_000_000)          # the panels really are late
            j0, j1 = edges[p], edges[p + 1]
            if op_b == 0:
                bv[j0:j1].copy_(bfv[j0:j1])       # k x n column-major: columns j0..j1 are contiguous rows of the view
            else:
                bv[:, j0:j1].copy_(bfv[:, j0:j1])  # n x k column-major: rows j0..j1 of every column
            events[p].record(side)
    assert oz.gemm_streamed_b(handle, op_a, op_b, m, n, k, 2.0, a, lda, b, ldb, beta, got, m, oz.fp64_int8(10), edges,
                              [e.cuda_event for e in events]) == 0
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    # invalid panel edges are rejected
    assert oz.gemm_streamed_b(handle, op_a, op_b, m, n, k, 2.0, a, lda, b, ldb, beta, got, m, oz.fp64_int8(10),
                              [0, 500, n], [events[0].cuda_event, events[1].cuda_event]) == 1
    # dgemm / auto fall back to the plain call after waiting for every panel
    got_d = to_dev(c0)
    assert oz.gemm_streamed_b(handle, op_a, op_b, m, n, k, 2.0, a, lda, b, ldb, beta, got_d, m,
                              oz.compute_mode_t.dgemm, edges, [e.cuda_event for e in events]) == 0
    torch.cuda.synchronize()
    assert torch.allclose(got_d, want, rtol=1e-11, atol=1e-11)
