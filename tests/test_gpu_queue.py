"""-m gpu, OPT-IN (OZIMMU_B200_TEST_QUEUE=1): the experimental device-side tile queue (ozk_gemm_i8_fused_queue and
OZIMMU_B200_E2E_QUEUE=1 for ozimmu_gemm_host).  Written at the end of round 1; its two hardware runs
(profiles/r1_queue_experiment.txt): all 7 tests below pass -- the queue launch is bit-identical to the static launch
with the flags preset and with flags that arrive while the launch is already spinning, and the host entry's queue mode
is bit-identical to the device entry.  Performance is not there yet (the block splits crawl on the few SMs the
persistent launch leaves free), so the mode stays off by default and these tests stay opt-in; run them under a short
`timeout`."""
import os

import numpy as np
import pytest
import torch

import oracle_lib
import ozimmu_b200 as oz
from gpu_util import bits, stream_ptr, to_dev

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("OZIMMU_B200_TEST_QUEUE") != "1",
                                 reason="experimental tile queue: set OZIMMU_B200_TEST_QUEUE=1")]


def _split(L, x, ld, rows, length, col_major, s, nbits):
    pitch = int(L.ozk_slice_pitch(length))
    out = torch.zeros(int(L.ozk_slices_bytes(rows, length, s)), dtype=torch.int8, device="cuda")
    mx = torch.zeros(rows, dtype=torch.float64, device="cuda")
    scr = torch.zeros(rows, dtype=torch.int32, device="cuda")
    assert L.ozk_split_int8(out.data_ptr(), pitch, mx.data_ptr(), scr.data_ptr(), rows, length, x.data_ptr(), ld,
                            int(col_major), s, nbits, stream_ptr()) == 0
    return out, mx, pitch


def _items(m, n, a_edges, b_edges, order="rows"):
    """one item per 256 x 256 tile: (tile, flag of its A block, flag of its B block, done counter = A block)"""
    tm_n, tn_n = -(-m // 256), -(-n // 256)
    block = lambda edges, r: max(i for i, e in enumerate(edges[:-1]) if e <= r)
    items = []
    rng = [(tm, tn) for tm in range(tm_n) for tn in range(tn_n)]
    if order == "reverse":
        rng = rng[::-1]
    for tm, tn in rng:
        items.append((tm | (tn << 16), block(a_edges, tm * 256), len(a_edges) - 1 + block(b_edges, tn * 256),
                      block(a_edges, tm * 256)))
    return np.array(items, dtype=np.uint32)


@pytest.mark.parametrize("m,n,k,order", [(1024, 1280, 600, "rows"), (900, 1100, 1030, "reverse"), (4096, 4096, 512, "rows")])
@pytest.mark.parametrize("late_flags", [False, True])
def test_queue_launch_equals_static_launch(m, n, k, order, late_flags):
    L = oz.lib()
    s, nbits = 9, int(L.ozk_bits_per_int8(k))
    a = to_dev(oracle_lib.gen_matrix("exp_rand-1", m * k, 71))   # op_t A: rows contiguous
    b = to_dev(oracle_lib.gen_matrix("exp_rand-1", k * n, 72))   # op_n B
    c0 = oracle_lib.gen_matrix("normal01", m * n, 73)
    a_sl, amax, pitch = _split(L, a, k, m, k, False, s, nbits)
    b_sl, bmax, _ = _split(L, b, k, n, k, False, s, nbits)
    want = to_dev(c0)
    assert L.ozk_gemm_i8_fused(m, n, k, a_sl.data_ptr(), b_sl.data_ptr(), pitch, amax.data_ptr(), bmax.data_ptr(), s,
                               nbits, 1.25, -0.5, want.data_ptr(), m, stream_ptr()) == 0
    a_edges = [0, 512, m] if m > 512 else [0, m]
    b_edges = [0, 256, 768, n] if n > 768 else [0, n]
    items = to_dev(_items(m, n, a_edges, b_edges, order))
    nflags = len(a_edges) - 1 + len(b_edges) - 1
    epoch = 7
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    done = torch.zeros(64, dtype=torch.int32, device="cuda")
    reserve = 8
    words = int(L.ozk_queue_scratch_words(items.shape[0], reserve))
    scratch = torch.empty(words, dtype=torch.int32, device="cuda")
    got = to_dev(c0)
    side = torch.cuda.Stream()
    # every kernel used while the persistent launch spins must have been launched once before (lazy module loading
    # can need a synchronisation the spinning kernel would never grant)
    torch.cuda._sleep(1000)
    for i in range(nflags):           # exactly the launches of the late path below: a fill of a 4-byte-offset view is a
        flags[i:i + 1].fill_(0)       # different (non-vectorised) kernel than a fill of an aligned tensor -- run38 hung
    if not late_flags:                # 4 s in precisely that first launch until the readiness wait timed out
        flags[:nflags] = epoch
    torch.cuda.synchronize()
    launch = torch.cuda.Stream()
    rc = L.ozk_gemm_i8_fused_queue(m, n, k, a_sl.data_ptr(), b_sl.data_ptr(), pitch, amax.data_ptr(), bmax.data_ptr(), s,
                                   nbits, 1.25, -0.5, got.data_ptr(), m, items.data_ptr(), items.shape[0],
                                   flags.data_ptr(), epoch, done.data_ptr(), scratch.data_ptr(), words, reserve,
                                   int(launch.cuda_stream))
    assert rc == 0
    if late_flags:
        # the launch is already spinning; the flags arrive one by one, last block first, from another stream
        with torch.cuda.stream(side):
            for i in reversed(range(nflags)):
                torch.cuda._sleep(1_000_000)
                flags[i:i + 1].fill_(epoch)
    torch.cuda.synchronize()
    assert int(scratch[1].item()) == 0, "a readiness wait timed out"
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    # 16 counts per tile, summed per A block
    tiles_per_block = np.bincount(items.cpu().numpy()[:, 3].astype(np.int64), minlength=64)
    assert np.array_equal(done.cpu().numpy().astype(np.int64), 16 * tiles_per_block)


@pytest.mark.parametrize("join", ["0", "1"])
def test_gemm_host_queue_mode(monkeypatch, join):
    """join=1 (a second launch takes the reserved SMs after the last split; 16 SMs reserved) has not run on hardware"""
    monkeypatch.setenv("OZIMMU_B200_E2E_QUEUE_JOIN", join)
    if join == "1":
        monkeypatch.setenv("OZIMMU_B200_E2E_QUEUE_RESERVE_SMS", "16")
    h = oz.create()
    try:
        m, n, k = 2300, 2000, 900
        a = oracle_lib.gen_matrix("exp_rand-1", m * k, 81)
        b = oracle_lib.gen_matrix("exp_rand-1", k * n, 82)
        c = oracle_lib.gen_matrix("normal01", m * n, 83)
        dc = to_dev(c)
        assert oz.gemm(h, 0, 0, m, n, k, 1.5, to_dev(a), m, to_dev(b), k, -0.5, dc, m, oz.fp64_int8(9)) == 0
        torch.cuda.synchronize()
        monkeypatch.setenv("OZIMMU_B200_E2E_QUEUE", "1")
        monkeypatch.setenv("OZIMMU_B200_E2E_PANEL", "512")
        monkeypatch.setenv("OZIMMU_B200_E2E_ROWBLOCK", "512")
        before = oz.launch_count()
        for call in range(3):   # the first call warms the kernels through the multi-launch path
            hc = c.copy()
            assert oz.gemm_host(h, 0, 0, m, n, k, 1.5, a, m, b, k, -0.5, hc, m, oz.fp64_int8(9)) == 0
            assert np.array_equal(bits(hc), bits(dc)), call
            if call == 0:
                first = oz.launch_count() - before
        # a queue-mode call launches ONE product kernel: far fewer launches than the multi-launch call
        assert (oz.launch_count() - before - first) / 2 < first
    finally:
        oz.destroy(h)


def test_gemm_streamed_b_queue_mode(monkeypatch):
    """gemm_streamed_b with OZIMMU_B200_STREAMED_QUEUE=1 (one queue launch instead of one launch per panel); panels
    completed late, in order, by copies on a side stream.  Passed on a B200 in the round's very last GPU call (run40);
    its speed against the multi-launch path has not been measured."""
    monkeypatch.setenv("OZIMMU_B200_STREAMED_QUEUE", "1")
    h = oz.create()
    try:
        m, n, k = 1800, 2300, 900
        a = to_dev(oracle_lib.gen_matrix("exp_rand-1", m * k, 91))
        b_full = to_dev(oracle_lib.gen_matrix("exp_rand-1", k * n, 92))
        c0 = oracle_lib.gen_matrix("normal01", m * n, 93)
        want = to_dev(c0)
        assert oz.gemm(h, 0, 0, m, n, k, 2.0, a, m, b_full, k, 0.5, want, m, oz.fp64_int8(9)) == 0
        torch.cuda.synchronize()
        edges = [0, 768, 1536, n]
        side = torch.cuda.Stream()
        bfv = b_full.view(n, k)
        for call in range(3):   # the first call goes through the multi-launch path and warms every kernel
            b = torch.zeros_like(b_full)
            bv = b.view(n, k)
            got = to_dev(c0)
            events = [torch.cuda.Event() for _ in range(3)]
            torch.cuda.synchronize()
            with torch.cuda.stream(side):
                for p in range(3):
                    torch.cuda._sleep(2_000_000)
                    bv[edges[p]:edges[p + 1]].copy_(bfv[edges[p]:edges[p + 1]])
                    events[p].record(side)
            assert oz.gemm_streamed_b(h, 0, 0, m, n, k, 2.0, a, m, b, k, 0.5, got, m, oz.fp64_int8(9), edges,
                                      [e.cuda_event for e in events]) == 0
            torch.cuda.synchronize()
            assert torch.equal(got.view(torch.int64), want.view(torch.int64)), call
    finally:
        oz.destroy(h)
