"""Row-block sharding of one DGEMM over the GPUs of a box: rank g owns the rows
[g*bm, min(m, (g+1)*bm)) of A and C, B is replicated by ONE broadcast from the owner rank
(NCCL over NVLink 5 / NVSwitch through torch.distributed), and every rank runs the single-GPU path
on its block.  Bit-identical to the one-GPU result: the split scales A per row and B per column and
K is never partitioned, so no cross-shard reduction exists.

By default B travels as ONE broadcast followed by one product over the whole row block (the fused
kernel then runs a single persistent launch, 14 full rounds of tiles at 8192^2).  `pipeline=True` sends B
in column panels (contiguous for op_n B) and runs split + product per panel while the next panel is on the
wire; on NVLink 5 the broadcast (512 MiB in < 1 ms) is too short to repay the ragged last round that each
per-panel launch adds, so it is off unless B is very large relative to the product.
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
    """(first column, width) of the broadcast panels: equal widths, multiples of 128, at least min_width."""
    panels = max(1, min(max_panels, n // max(1, min_width)))
    w = -(-n // panels)
    w = -(-w // 128) * 128
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
    if not pipeline or int(op_B) != int(api.op_n) or m_local == 0:
        dist.broadcast(flat, src=src, group=group)
        if m_local == 0:
            return 0
        return api.gemm(handle, op_A, op_B, m_local, n, k, alpha, a_block, lda, b, ldb, beta, c_block, ldc,
                        compute_mode)
    # op_n B: k x n column-major, panel [j0, j0 + w) is the contiguous range [j0*ldb, (j0+w)*ldb)
    panels = column_panels(n)
    works = []
    for (j0, w) in panels:
        hi = min(flat.numel(), (j0 + w) * ldb)
        works.append(dist.broadcast(flat[j0 * ldb:hi], src=src, group=group, async_op=True))
    cflat = c_block.view(-1)
    rc = 0
    for (j0, w), work in zip(panels, works):
        work.wait()  # orders the current stream after this panel's broadcast
        rc |= api.gemm(handle, op_A, op_B, m_local, w, k, alpha, a_block, lda, flat[j0 * ldb:], ldb, beta,
                       cflat[j0 * ldc:], ldc, compute_mode)
    return rc
