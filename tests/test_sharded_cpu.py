"""CPU: host logic of the multi-GPU path -- the row-block partition (ozimmu_row_block) covers every row exactly once,
the broadcast panels (ozimmu_sharded_panel_edges) tile the columns on the kernel's granularity, the unique-id exchange
runs over a world_size-2 gloo group, and a world_size-2 gloo run of the panel broadcast + per-rank product (the
per-rank product stood in for by the CPU oracle: no GPU here) reproduces the single-process result bit for bit."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

from ozimmu_b200.sharded import column_panels, row_block  # noqa: E402


@pytest.mark.parametrize("m,world", [(16384, 8), (8192, 1), (1023, 4), (5, 8), (0, 2), (1025, 2)])
def test_row_blocks_partition(m, world):
    seen = np.zeros(m, dtype=int)
    prev_end = 0
    for r in range(world):
        r0, rows = row_block(m, world, r)
        assert r0 == prev_end or rows == 0
        seen[r0:r0 + rows] += 1
        prev_end = r0 + rows
    assert (seen == 1).all()


@pytest.mark.parametrize("n", [1, 127, 1024, 4096, 8192, 16384, 5000])
def test_column_panels_cover(n):
    p = column_panels(n)
    assert 1 <= len(p) <= 8 and all(w >= min(n, 1024) or i == len(p) - 1 for i, (_, w) in enumerate(p))
    assert p[0][0] == 0 and sum(w for _, w in p) == n
    assert all(p[i][0] + p[i][1] == p[i + 1][0] for i in range(len(p) - 1))
    assert all(j0 % 256 == 0 for j0, _ in p)   # inner edges on the kernel tile / block-split granularity


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from ozimmu_b200.sharded import column_panels, exchange_unique_id, row_block

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the communicator bootstrap: rank 0's 128-byte id reaches every rank (the id itself needs NCCL: a stand-in here)
    uid = exchange_unique_id(lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    m, n, k, s = 70, 1500, 48, 9
    A = oracle_lib.gen_matrix("exp_rand-1", m * k, 11).reshape(k, m)   # column-major m x k
    r0, rows = row_block(m, world, rank)
    a_block = np.ascontiguousarray(A[:, r0:r0 + rows])                 # column-major rows x k, ld = rows
    if rank == 0:
        b = torch.from_numpy(oracle_lib.gen_matrix("exp_rand-1", k * n, 12))
    else:
        b = torch.zeros(k * n, dtype=torch.float64)
    # the same pipelined panel broadcast sharded_gemm issues, over gloo
    works = [dist.broadcast(b[j0 * k:(j0 + w) * k], src=0, async_op=True) for (j0, w) in column_panels(n)]
    for w in works:
        w.wait()
    c = oracle_lib.oracle_gemm(0, 0, rows, n, k, 1.0, a_block.ravel(), max(rows, 1), b.numpy(), k, 0.0,
                               np.zeros(max(rows, 1) * n), max(rows, 1), s)
    np.save(os.path.join(out_dir, f"c_{rank}.npy"), c.reshape(n, max(rows, 1))[:, :rows])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single(tmp_path):
    import torch.multiprocessing as mp
    import oracle_lib

    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    m, n, k, s = 70, 1500, 48, 9
    a = oracle_lib.gen_matrix("exp_rand-1", m * k, 11)
    b = oracle_lib.gen_matrix("exp_rand-1", k * n, 12)
    full = oracle_lib.oracle_gemm(0, 0, m, n, k, 1.0, a, m, b, k, 0.0, np.zeros(m * n), m, s).reshape(n, m)
    got = np.concatenate([np.load(tmp_path / f"c_{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(got.view(np.int64), full.view(np.int64))
