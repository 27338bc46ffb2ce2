"""-m gpu: the multi-GPU path on real devices (skipped with fewer than 2 GPUs): 2 ranks, NCCL broadcast of
B, row blocks of A/C -- the concatenated result equals the single-GPU product bit for bit."""
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
m, n, k, s = 1100, 2300, 700, 9
A = torch.from_numpy(oracle_lib.gen_matrix("exp_rand-1", m * k, 5)).view(k, m)   # column-major m x k
r0, rows = oz.row_block(m, world, rank)
a_blk = A[:, r0:r0 + rows].contiguous().cuda()
b = (torch.from_numpy(oracle_lib.gen_matrix("exp_rand-1", k * n, 6)) if rank == 0 else torch.zeros(k * n, dtype=torch.float64)).cuda()
c_blk = torch.zeros(n, rows, dtype=torch.float64, device="cuda")
h = oz.create()
assert oz.sharded_gemm(h, 0, 0, rows, n, k, 1.0, a_blk, rows, b, k, 0.0, c_blk, rows, oz.fp64_int8(s), src=0) == 0
torch.cuda.synchronize()
c_pipe = torch.zeros_like(c_blk)
if rank != 0:
    b.zero_()
assert oz.sharded_gemm(h, 0, 0, rows, n, k, 1.0, a_blk, rows, b, k, 0.0, c_pipe, rows, oz.fp64_int8(s), src=0, pipeline=True) == 0
torch.cuda.synchronize()
assert torch.equal(c_pipe.view(torch.int64), c_blk.view(torch.int64)), "panel-pipelined broadcast differs"
for rep in range(2):   # second call reuses the cached IPC mapping of B
    c_peer = torch.zeros_like(c_blk)
    if rank != 0:
        b.zero_()
    assert oz.sharded_gemm(h, 0, 0, rows, n, k, 1.0, a_blk, rows, b, k, 0.0, c_peer, rows, oz.fp64_int8(s), src=0, transport="peer") == 0
    torch.cuda.synchronize()
    assert torch.equal(c_peer.view(torch.int64), c_blk.view(torch.int64)), "peer-pull transport differs"
torch.save(c_blk.cpu(), os.environ["OZ_OUT"] + f"/c_{rank}.pt")
dist.barrier(); oz.destroy(h); dist.destroy_process_group()
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
