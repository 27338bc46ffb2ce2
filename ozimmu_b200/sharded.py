"""Row-block sharding of one DGEMM over the GPUs of a box (BASELINE config 4) -- a thin mirror of the
multi-GPU section of the C-ABI (include/ozimmu_b200.h, csrc/sharded.cpp, csrc/host_e2e.cu).

Rank g owns the rows [g*bm, min(m, (g+1)*bm)) of A and C, B is replicated from its owner rank by NCCL broadcasts over
NVLink 5 / NVSwitch (in column panels, each panel of C starting as soon as its columns have landed), and every rank
runs the single-GPU path on its block.  Bit-identical to the one-GPU result: the split scales A per row and B per
column and K is never partitioned, so no cross-shard reduction exists.

The communicator is the library's own (ncclCommInitRank inside libozimmu.so, NCCL resolved at run time); Python only
ships the 128-byte NCCL unique id from rank 0 to the other ranks -- here through torch.distributed, which an
application would replace by whatever bootstrap it has (MPI, a file, a socket).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

from . import _lib, api


def row_block(m: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first row, number of rows) of rank's block; blocks are ceil(m / world) tall, the last one short."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    r0, rows = C.c_size_t(), C.c_size_t()
    _lib.lib().ozimmu_row_block(m, world_size, rank, C.byref(r0), C.byref(rows))
    return int(r0.value), int(rows.value)


def column_panels(n: int, max_panels: int = 8) -> List[Tuple[int, int]]:
    """(first column, width) of the panels B is broadcast in: equal widths, multiples of 256 (the kernel's tile and the
    block-wise split's granularity), at least 1024 wide."""
    edges = (C.c_size_t * 32)()
    cnt = int(_lib.lib().ozimmu_sharded_panel_edges(n, max_panels, C.addressof(edges), 32))
    e = [int(edges[i]) for i in range(cnt)]
    return [(a, b - a) for a, b in zip(e, e[1:])]


class Comm:
    """One communicator of the library (ozimmu_comm_t) on the current CUDA device."""

    def __init__(self, raw: int, rank: int, size: int):
        self.raw, self.rank, self.size = raw, rank, size

    def destroy(self) -> None:
        if self.raw:
            _lib.lib().ozimmu_comm_destroy(self.raw)
            self.raw = 0


def exchange_unique_id(make_id, rank: int, group=None, src: int = 0) -> bytes:
    """rank `src` calls make_id() -> 128 bytes; every rank of the torch.distributed group returns those bytes"""
    import torch.distributed as dist
    payload = [make_id() if rank == src else None]
    dist.broadcast_object_list(payload, src=src, group=group)
    uid = bytes(payload[0])
    if len(uid) != 128:
        raise RuntimeError("NCCL unique id must be 128 bytes")
    return uid


def _make_unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    api._check(_lib.lib().ozimmu_comm_unique_id(C.addressof(buf)), "comm_unique_id")
    return bytes(buf)


def comm_create(group=None) -> Optional[Comm]:
    """Collective over a torch.distributed group (one process per GPU, torch.cuda current device = this rank's GPU):
    the library's own NCCL communicator for the group.  None without an initialised group / with one rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    rank, size = dist.get_rank(group), dist.get_world_size(group)
    uid = exchange_unique_id(_make_unique_id, rank, group)
    raw = C.c_void_p()
    buf = (C.c_ubyte * 128).from_buffer_copy(uid)
    api._check(_lib.lib().ozimmu_comm_create(C.byref(raw), size, rank, C.addressof(buf)), "comm_create")
    return Comm(raw.value, rank, size)


def sharded_gemm(handle: api.handle_t, comm: Optional[Comm], op_A: int, op_B: int, m_local: int, n: int, k: int,
                 alpha: float, a_block, lda: int, b, ldb: int, beta: float, c_block, ldc: int, compute_mode, *,
                 src: int = 0, max_panels: int = 8) -> int:
    """C_block = alpha * op(A_block) * op(B) + beta * C_block on every rank (device operands, asynchronous on the
    handle's stream).  b: device buffer of the full B on every rank; its CONTENT is taken from rank `src`.
    comm=None: a plain single-GPU gemm."""
    al, be = C.c_double(alpha), C.c_double(beta)
    rc = _lib.lib().ozimmu_gemm_sharded(handle.raw, comm.raw if comm else None, int(op_A), int(op_B), m_local, n, k,
                                        C.addressof(al), api._ptr(a_block), lda, api._ptr(b), ldb, C.addressof(be),
                                        api._ptr(c_block), ldc, int(compute_mode), src, max_panels)
    return api._check(rc, "gemm_sharded")


def sharded_gemm_host(handle: api.handle_t, comm: Optional[Comm], op_A: int, op_B: int, m_local: int, n: int, k: int,
                      alpha: float, a_block_host, lda: int, b_host, ldb: int, beta: float, c_block_host, ldc: int,
                      compute_mode, *, src: int = 0) -> int:
    """The same with HOST operands (pinned CPU tensors / numpy arrays): b_host is read on rank `src` only (may be None
    elsewhere); returns when this rank's block of C is complete."""
    al, be = C.c_double(alpha), C.c_double(beta)
    rc = _lib.lib().ozimmu_gemm_sharded_host(handle.raw, comm.raw if comm else None, int(op_A), int(op_B), m_local, n, k,
                                             C.addressof(al), api._ptr(a_block_host), lda, api._ptr(b_host), ldb,
                                             C.addressof(be), api._ptr(c_block_host), ldc, int(compute_mode), src)
    return api._check(rc, "gemm_sharded_host")
