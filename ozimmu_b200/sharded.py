"""Row-block sharding of one DGEMM over the GPUs of a box: rank g owns the rows
[g*bm, min(m, (g+1)*bm)) of A and C, B is replicated by ONE broadcast from the owner rank
(NCCL over NVLink 5 / NVSwitch through torch.distributed), and every rank runs the single-GPU path
on its block.  Bit-identical to the one-GPU result: the split scales A per row and B per column and
K is never partitioned, so no cross-shard reduction exists.

`pipeline=True` (op_n B) sends B in contiguous column panels and hands the panels' arrival events to
`gemm_streamed_b`: split(A) runs while the first panel is on the wire, and every panel of C is computed as soon
as its columns of B have landed (the fused launches rotate over several streams, so a launch back-fills the SMs
its predecessor leaves idle in its last round of tiles).  `pipeline=False` (default) is one broadcast followed
by one product launch.  Measured on 2 x B200 at 8192 rows per rank the pipeline is SLOWER (21.1 vs 19.1 ms per step,
profiles/r1_bench_2gpu_streamed_b.txt): NCCL's broadcast kernels need SMs of their own, the persistent product
kernel occupies every SM, so each later panel's broadcast waits for a whole round of tiles to drain.  It pays only
when the transport does not need SMs (copy-engine peer copies, host staging) -- kept for those callers.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

from . import api


def row_block(m: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first row, number of rows) of rank's block; blocks are ceil(m / world) tall, the last one short."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    bm = -(-m // world_size)
    r0 = min(m, rank * bm)
    return r0, max(0, min(bm, m - r0))


def column_panels(n: int, max_panels: int = 4, min_width: int = 1024) -> List[Tuple[int, int]]:
    """(first column, width) of the broadcast panels: equal widths, multiples of 256 (the kernel's tile and the
    block-wise split's granularity), at least min_width."""
    panels = max(1, min(max_panels, n // max(1, min_width)))
    w = -(-n // panels)
    w = -(-w // 256) * 256
    out = []
    j = 0
    while j < n:
        out.append((j, min(w, n - j)))
        j += w
    return out


def sharded_gemm(handle: api.handle_t, op_A: int, op_B: int, m_local: int, n: int, k: int, alpha: float, a_block,
                 lda: int, b, ldb: int, beta: float, c_block, ldc: int, compute_mode, *, src: int = 0, group=None,
                 pipeline: bool = False) -> int:
    """C_block = alpha * op(A_block) * op(B) + beta * C_block on every rank.

    a_block / c_block: this rank's rows (device, column-major).  b: device buffer of the full B on every
    rank; its CONTENT is taken from rank `src` (the broadcast overwrites the other ranks' copies).
    Without an initialised process group (single GPU) this is a plain api.gemm.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return api.gemm(handle, op_A, op_B, m_local, n, k, alpha, a_block, lda, b, ldb, beta, c_block, ldc,
                        compute_mode)
    flat = b.view(-1)
    if not pipeline or int(op_B) != int(api.op_n) or m_local == 0 or n < 2 * _MIN_PANEL:
        dist.broadcast(flat, src=src, group=group)
        if m_local == 0:
            return 0
        return api.gemm(handle, op_A, op_B, m_local, n, k, alpha, a_block, lda, b, ldb, beta, c_block, ldc,
                        compute_mode)
    # op_n B: k x n column-major, panel [j0, j0 + w) is the contiguous range [j0*ldb, (j0+w)*ldb)
    panels = column_panels(n, max_panels=8, min_width=_MIN_PANEL)
    side = _side_stream()
    side.wait_stream(torch.cuda.current_stream())     # whatever produced / last read `b` on this stream
    events = []
    with torch.cuda.stream(side):
        for (j0, w) in panels:
            hi = min(flat.numel(), (j0 + w) * ldb)
            work = dist.broadcast(flat[j0 * ldb:hi], src=src, group=group, async_op=True)
            work.wait()                               # orders `side` after this panel's broadcast
            ev = torch.cuda.Event()
            ev.record(side)
            events.append(ev)
    edges = [j0 for (j0, _) in panels] + [n]
    rc = api.gemm_streamed_b(handle, op_A, op_B, m_local, n, k, alpha, a_block, lda, b, ldb, beta, c_block, ldc,
                             compute_mode, edges, [ev.cuda_event for ev in events])
    _keep_alive(events)
    return rc


_MIN_PANEL = 1024
_side = {}
_live_events: list = []


def _side_stream():
    """one broadcast-ordering stream per device"""
    import torch
    dev = torch.cuda.current_device()
    if dev not in _side:
        _side[dev] = torch.cuda.Stream(device=dev)
    return _side[dev]


def _keep_alive(events) -> None:
    """the library's streams still wait on these events after this call returns: keep the last two calls' events"""
    _live_events.append(events)
    del _live_events[:-2]
