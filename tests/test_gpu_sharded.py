"""-m gpu: the multi-GPU path on real devices (skipped with fewer than 2 GPUs): 2 ranks, the library's own NCCL
communicator, row blocks of A/C, B broadcast whole / in column panels / block by block from the owner's host memory --
every variant's concatenated result equals the single-GPU product bit for bit."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["OZ_ROOT"]); sys.path.insert(0, os.environ["OZ_ROOT"] + "/tests")
import ozimmu_b200 as oz, oracle_lib
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
comm = oz.comm_create()
assert comm is not None and comm.rank == rank and comm.size == world
m, n, k, s = 1100, 2300, 700, 9
A = torch.from_numpy(oracle_lib.gen_matrix("exp_rand-1", m * k, 5)).view(k, m)   # column-major m x k
r0, rows = oz.row_block(m, world, rank)
a_host = A[:, r0:r0 + rows].contiguous().pin_memory()
a_blk = a_host.cuda()
b_host = torch.from_numpy(oracle_lib.gen_matrix("exp_rand-1", k * n, 6)).pin_memory() if rank == 0 else None
b = b_host.cuda() if rank == 0 else torch.zeros(k * n, dtype=torch.float64, device="cuda")
c_blk = torch.zeros(n, rows, dtype=torch.float64, device="cuda")
h = oz.create()
# one broadcast, then one product launch
assert oz.sharded_gemm(h, comm, 0, 0, rows, n, k, 1.0, a_blk, rows, b, k, 0.0, c_blk, rows, oz.fp64_int8(s), src=0, max_panels=1) == 0
torch.cuda.synchronize()
# B in column panels, every panel of C as soon as its columns have landed (twice: events and buffers are reused)
for rep in range(2):
    c_pipe = torch.zeros_like(c_blk)
    if rank != 0:
        b.zero_()
    assert oz.sharded_gemm(h, comm, 0, 0, rows, n, k, 1.0, a_blk, rows, b, k, 0.0, c_pipe, rows, oz.fp64_int8(s), src=0, max_panels=2) == 0
    torch.cuda.synchronize()
    assert torch.equal(c_pipe.view(torch.int64), c_blk.view(torch.int64)), "panel-pipelined broadcast differs"
# host operands: every rank uploads its rows of A, rank 0 uploads and forwards B block by block
for env in ({}, {"OZIMMU_B200_E2E_PANEL": "1024", "OZIMMU_B200_E2E_ROWBLOCK": "256"}):
    os.environ.update(env)
    c_host = torch.zeros(n, rows, dtype=torch.float64).pin_memory()
    assert oz.sharded_gemm_host(h, comm, 0, 0, rows, n, k, 1.0, a_host, rows, b_host, k, 0.0, c_host, rows, oz.fp64_int8(s), src=0) == 0
    assert torch.equal(c_host.view(torch.int64), c_blk.cpu().view(torch.int64)), "host-operand sharded entry differs"
# op_t B (one block, one broadcast) and beta != 0 through the host entry
Bt = torch.from_numpy(oracle_lib.gen_matrix("exp_rand-1", k * n, 6)).view(n, k).T.contiguous()   # n x k column-major
bt_host = Bt.pin_memory() if rank == 0 else None
c0 = torch.from_numpy(oracle_lib.gen_matrix("normal01", n * rows, 7 + rank)).view(n, rows)
c_dev = c0.cuda()
assert oz.gemm(h, 0, 0, rows, n, k, 1.5, a_blk, rows, b, k, -0.5, c_dev, rows, oz.fp64_int8(s)) == 0
c_host = c0.clone().pin_memory()
assert oz.sharded_gemm_host(h, comm, 0, 1, rows, n, k, 1.5, a_host, rows, bt_host, n, -0.5, c_host, rows, oz.fp64_int8(s), src=0) == 0
torch.cuda.synchronize()
assert torch.equal(c_host.view(torch.int64), c_dev.cpu().view(torch.int64)), "op_t B / beta host-operand sharded entry differs"
torch.save(c_blk.cpu(), os.environ["OZ_OUT"] + f"/c_{rank}.pt")
dist.barrier(); oz.destroy(h); comm.destroy(); dist.destroy_process_group()
"""


def test_two_gpu_row_sharding_bit_exact(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import oracle_lib
    import ozimmu_b200 as oz
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OZ_ROOT=str(ROOT), OZ_OUT=str(tmp_path))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    m, n, k, s = 1100, 2300, 700, 9
    a = torch.from_numpy(oracle_lib.gen_matrix("exp_rand-1", m * k, 5)).cuda()
    b = torch.from_numpy(oracle_lib.gen_matrix("exp_rand-1", k * n, 6)).cuda()
    c = torch.zeros(n, m, dtype=torch.float64, device="cuda")
    h = oz.create()
    assert oz.gemm(h, 0, 0, m, n, k, 1.0, a, m, b, k, 0.0, c, m, oz.fp64_int8(s)) == 0
    torch.cuda.synchronize()
    got = torch.cat([torch.load(tmp_path / f"c_{r}.pt") for r in range(2)], dim=1)
    assert torch.equal(got.view(torch.int64), c.cpu().view(torch.int64))
    oz.destroy(h)
