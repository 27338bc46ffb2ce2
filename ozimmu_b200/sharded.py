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
                 pipeline: bool = False, transport: str = "nccl") -> int:
    """C_block = alpha * op(A_block) * op(B) + beta * C_block on every rank.

    a_block / c_block: this rank's rows (device, column-major).  b: device buffer of the full B on every
    rank; its CONTENT is taken from rank `src` (the broadcast overwrites the other ranks' copies).
    Without an initialised process group (single GPU) this is a plain api.gemm.

    transport="nccl": B is replicated by a NCCL broadcast (`pipeline` = in column panels feeding gemm_streamed_b).
    transport="peer" (op_n B): every rank PULLS B's column panels from the owner's buffer over NVLink with the copy
    engines (CUDA IPC mapping of the owner's allocation, established on first use of a buffer) and computes each
    panel of C as it lands -- no SMs are needed for the transfer, so it overlaps the persistent product kernel.
    The owner must not overwrite B before its next collective call on the same group.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return api.gemm(handle, op_A, op_B, m_local, n, k, alpha, a_block, lda, b, ldb, beta, c_block, ldc,
                        compute_mode)
    flat = b.view(-1)
    if transport == "peer" and int(op_B) == int(api.op_n) and n >= 2 * _MIN_PANEL:
        return _peer_pull_gemm(handle, op_A, op_B, m_local, n, k, alpha, a_block, lda, b, ldb, beta, c_block, ldc,
                               compute_mode, src, group)
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


_peer_views: dict = {}


def _peer_view(b, src: int, group):
    """the owner's B buffer mapped into this process (CUDA IPC), established once per buffer; collective"""
    import torch
    import torch.distributed as dist
    from torch.multiprocessing.reductions import reduce_tensor

    key = (b.data_ptr(), b.numel(), src, id(group))
    if key not in _peer_views:
        rank = dist.get_rank(group)
        payload = [reduce_tensor(b.view(-1)) if rank == src else None]
        dist.broadcast_object_list(payload, src=src, group=group)
        if rank == src:
            view = b.view(-1)
        else:
            rebuild, args = payload[0]
            view = rebuild(*args)          # a tensor on the owner's device, readable from here over NVLink
        _peer_views[key] = view
    return _peer_views[key]


def _peer_pull_gemm(handle, op_A, op_B, m_local, n, k, alpha, a_block, lda, b, ldb, beta, c_block, ldc, compute_mode,
                    src, group) -> int:
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    flat = b.view(-1)
    peer = _peer_view(b, src, group)
    # the owner's B is final once every rank has passed this point in stream order (and the previous call's pulls
    # have completed: each rank's stream waited for them before it got here)
    dist.barrier(group=group)
    panels = column_panels(n, max_panels=8, min_width=_MIN_PANEL)
    cur = torch.cuda.current_stream()
    events = []
    if rank == src:
        ev = torch.cuda.Event()
        ev.record(cur)
        events = [ev] * len(panels)
    else:
        side = _side_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for (j0, w) in panels:
                hi = min(flat.numel(), (j0 + w) * ldb)
                flat[j0 * ldb:hi].copy_(peer[j0 * ldb:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
                events.append(ev)
    if m_local == 0:
        rc = 0
        for ev in events:
            cur.wait_event(ev)
    else:
        edges = [j0 for (j0, _) in panels] + [n]
        rc = api.gemm_streamed_b(handle, op_A, op_B, m_local, n, k, alpha, a_block, lda, b, ldb, beta, c_block, ldc,
                                 compute_mode, edges, [ev.cuda_event for ev in events])
        if rank != src:
            cur.wait_event(events[-1])   # the next call's barrier is ordered after this call's last pull
    _keep_alive(events)
    return rc
