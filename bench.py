#!/usr/bin/env python
"""bench.py -- FP64-equivalent TFLOP/s of the Ozaki-scheme DGEMM at fp64_int8_9, 8192^3 (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one C = A*B (N/N, alpha=1, beta=0, tight leading dimensions) of the workload on every rank.
N=1: 8192^3.  N>1 (torchrun, one rank per GPU): weak scaling -- every rank owns an 8192-row block of A and C (global
m = 8192*N), rank 0 owns B and broadcasts it inside the timed step (ozimmu_gemm_sharded: the library's own NCCL
communicator, B in column panels, each panel of C as soon as it has landed).  The JSON line carries:
  value    device-resident throughput (CUDA events, max over ranks)
  e2e      the same through the C-ABI with HOST operands (pinned), H2D/D2H inside the timed region
           (ozimmu_gemm_host / ozimmu_gemm_sharded_host)
  roofline the fused tcgen05 kernel: int8 ops per launch / its CUDA-event duration, against
           2 x the measured bf16 peak of MEASURED_PEAKS.json (int8 runs at twice the bf16 rate)
  parity   (N>1) every rank's first 256 rows of C, bit-compared on rank 0 with a single-GPU product of the same rows
  config4  BASELINE config 4 in the same run: 16384^3 strong scaling, 16384/N rows per rank, B broadcast in the step
  per_rank_ms / bcast_ms   each rank's own step time and the cost of the bare broadcast (what limits the scaling)
  cpu_baseline  host OpenBLAS DGEMM (numpy) on the box's cores -- the reference has no CPU path and
           north_star names host OpenBLAS as the CPU comparator -- plus the scalar oracle port on a small sample
--impl reference times the UNMODIFIED reference (oracle/_ref/libozref.so, built from /root/reference by
oracle/Makefile) on the same workload; it is a GPU library, so that arm runs on the GPU as well.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_DEFAULT = int(os.environ.get("OZ_BENCH_N", "8192"))
NUM_SPLIT = int(os.environ.get("OZ_BENCH_SPLIT", "9"))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled
    every 5 ms from a thread (nvidia-smi -lms cannot deliver more than a couple of samples in a 0.2 s
    region); falls back to nvidia-smi when pynvml is unavailable."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []          # (sm_mhz, power_w, reasons bitmask)
        self.max_mhz = None
        self.stop_flag = False
        self.thread = None
        self.proc = None
        self.lines = []
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    idx = int(ids[self.index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                try:
                    rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:  # noqa: BLE001
                    rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, pw, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            sm = sorted(x[0] for x in self.samples)
            bits = 0
            for x in self.samples:
                bits |= x[2]
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz,
                    "power_w_max": max(x[1] for x in self.samples), "samples": len(sm), "source": "nvml 5 ms poll",
                    "reasons": sorted(v for k, v in names.items() if bits & k)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, flag in zip(names, f[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm),
                "source": "nvidia-smi -lms 100", "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index: int) -> None:
    """Pin this rank's process (and with it the first-touch placement of its pinned host buffers) to the CPUs NVML
    reports as local to its GPU: with eight ranks moving host operands at once, buffers on the far socket halve the
    PCIe rate.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                idx = int(ids[index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # noqa: BLE001
        pass


def dist_setup(gpus: int):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier_sync(world: int):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms: float, world: int) -> float:
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed_loop(fn, steps: int, warmup: int, world: int, sampler=None) -> float:
    """ms per step: W untimed steps, then exactly K steps between barrier+sync, CUDA events, max over ranks."""
    import torch
    for _ in range(warmup):
        fn()
    if sampler is not None:
        sampler.start()      # BEFORE the barrier: see timed_loop_ranks
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier_sync(world)
    return max_over_ranks(e0.elapsed_time(e1), world) / steps


def make_inputs(n: int, rank: int):
    """urand01 (0,1] operands generated on the device (reference test/main_test.cu:195-202), fixed seeds"""
    import torch
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    a = 1.0 - torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g)
    gb = torch.Generator(device="cuda").manual_seed(99)
    b = 1.0 - torch.rand(n * n, dtype=torch.float64, device="cuda", generator=gb)
    c = torch.zeros(n * n, dtype=torch.float64, device="cuda")
    return a, b, c


# ------------------------------------------------------------------------------------------------
def cpu_baseline(n_sample: int = N_DEFAULT) -> dict:
    """host OpenBLAS DGEMM through numpy on a bounded sample + the scalar oracle port on a tiny one"""
    import numpy as np
    try:
        from threadpoolctl import threadpool_info
        threads = max([d.get("num_threads", 1) for d in threadpool_info() if d.get("user_api") == "blas"] or [1])
    except Exception:  # noqa: BLE001
        threads = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    a = rng.random((n_sample, n_sample))
    b = rng.random((n_sample, n_sample))
    a @ b  # warm-up
    reps, t0 = 0, time.perf_counter()
    while reps < 2 or (time.perf_counter() - t0 < 8.0 and reps < 10):
        a @ b
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    out = {"value": 2.0 * n_sample ** 3 / dt / 1e12, "unit": "TFLOP/s", "cores": threads, "kind": "port",
           "sample": f"host OpenBLAS DGEMM via numpy, {n_sample}^3 x{reps} on {threads} threads of {os.cpu_count()} "
                     f"logical cores (the reference has no CPU path; north_star names host OpenBLAS as the CPU comparator)"}
    try:
        import oracle_lib
        m = 256
        x = 1.0 - rng.random(m * m)
        y = 1.0 - rng.random(m * m)
        t0 = time.perf_counter()
        oracle_lib.oracle_gemm(0, 0, m, m, m, 1.0, x, m, y, m, 0.0, np.zeros(m * m), m, NUM_SPLIT)
        dt = time.perf_counter() - t0
        out["oracle_port"] = {"value": 2.0 * m ** 3 / dt / 1e12, "unit": "TFLOP/s", "cores": 1,
                              "sample": f"oracle/oz_oracle.c (scalar C restatement of the reference) {m}^3 fp64_int8_{NUM_SPLIT} x1"}
    except Exception as e:  # noqa: BLE001
        out["oracle_port"] = {"unavailable": str(e)[:200]}
    return out


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        # the roofline launches below are timed one at a time (events + synchronize around each): the burst figure
        return {"bf16": d.get("bf16_tflops", d.get("bf16_tflops_sustained")),
                "bf16_sustained": d.get("bf16_tflops_sustained"), "hbm": d.get("hbm_gbs"), "src": "measured"}
    return {"bf16": 1400.0, "bf16_sustained": None, "hbm": 6650.0, "src": "fallback"}


def ncu_traffic_bytes():
    """(dram bytes per launch of the fused kernel, file it comes from) from the newest committed ncu capture of the
    shipped kernel (profiles/), or (None, None).  ncu cannot run inside the timed bench; the capture is re-taken
    whenever the kernel changes and its file name is stamped into the line."""
    for name in ("r2_fused_pair256_8192.json", "r1_fused_pair256_final_8192.json"):
        p = ROOT / "profiles" / name
        if p.exists():
            try:
                d = json.loads(p.read_text())
                return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), "profiles/" + name
            except Exception:  # noqa: BLE001
                continue
    return None, None


# ------------------------------------------------------------------------------------------------
WORKLOAD = "DGEMM N/N {n}x{n}x{n}, fp64_int8_{s}, alpha=1 beta=0, tight ld"   # the same string in both arms


def gather_ms(ms: float, world: int):
    """every rank's own time (ms), on every rank"""
    if world == 1:
        return [ms]
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def timed_loop_ranks(fn, steps: int, warmup: int, world: int, sampler=None):
    """timed_loop that also returns every rank's own ms per step"""
    import torch
    for _ in range(warmup):
        fn()
    # the clock sampler (NVML initialisation + a thread, tens of milliseconds on rank 0 only) starts BEFORE the barrier:
    # started after it, rank 0 entered the timed loop late and every other rank's first broadcast -- inside ITS timed
    # region -- waited for it: 2-5 ms per step on the receiving ranks of a 10-step loop (calls 40 / 42)
    if sampler is not None:
        sampler.start()
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier_sync(world)
    mine = e0.elapsed_time(e1) / steps
    per_rank = gather_ms(mine, world)
    return max(per_rank), per_rank


def accuracy_vs_cublas(a, b, c, n: int) -> dict:
    """The second half of BASELINE's metric: max relative error of C against cuBLAS DGEMM (torch.mm in FP64) on the
    same operands, plus the relative residual in the Frobenius norm.  a, b, c: column-major n x n, flat."""
    import torch
    # column-major C = A B  <=>  row-major view: C^T = B^T A^T, and a flat column-major buffer viewed (n, n) IS the transpose
    ref = torch.mm(b.view(n, n), a.view(n, n)).reshape(-1)
    diff = (c - ref).abs()
    out = {"max_rel_err_vs_cublas_dgemm": float((diff / ref.abs().clamp_min(1e-300)).max().item()),
           "rel_residual_vs_cublas_dgemm": float((torch.linalg.vector_norm(diff) / torch.linalg.vector_norm(ref)).item()),
           "how": "torch.mm (cuBLAS DGEMM) on the same device operands after the timed region; max |C - C_cublas| / |C_cublas| "
                  "over all elements, ||C - C_cublas||_F / ||C_cublas||_F"}
    del ref, diff
    return out


def sharded_parity(oz, h, comm, rank, world, n, k, a, lda, b, c, ldc, rows, mode, sample=256):
    """Bit-compare every rank's first `sample` rows of C (computed by the sharded step, inside its full row block)
    with a SINGLE-GPU product of the same rows of A on rank 0.  Returns {"checked", "max_ulp", "rows_per_rank"}."""
    import torch
    import torch.distributed as dist
    sample = min(sample, rows)
    # column-major (ld x cols): rows 0..sample of every column
    a_s = a.view(-1, lda)[:, :sample].contiguous()       # k x sample  == column-major sample x k, ld = sample
    c_s = c.view(-1, ldc)[:, :sample].contiguous()       # n x sample
    a_all = [torch.empty_like(a_s) for _ in range(world)]
    c_all = [torch.empty_like(c_s) for _ in range(world)]
    dist.all_gather(a_all, a_s)
    dist.all_gather(c_all, c_s)
    max_ulp = 0
    if rank == 0:
        chk = torch.empty_like(c_s)
        for r in range(world):
            assert oz.gemm(h, oz.op_n, oz.op_n, sample, n, k, 1.0, a_all[r], sample, b, k, 0.0, chk, sample, mode) == 0
            torch.cuda.synchronize()
            x, y = chk.view(torch.int64), c_all[r].view(torch.int64)
            if not torch.equal(x, y):
                # ulp distance of finite doubles through the monotone integer mapping
                fx = torch.where(x < 0, torch.iinfo(torch.int64).min - x, x)
                fy = torch.where(y < 0, torch.iinfo(torch.int64).min - y, y)
                max_ulp = max(max_ulp, int((fx - fy).abs().max().item()))
    t = torch.tensor([max_ulp], dtype=torch.int64, device="cuda")
    dist.broadcast(t, src=0)
    return {"checked": True, "max_ulp": int(t.item()), "rows_per_rank": sample,
            "how": "rank 0: single-GPU ozimmu_gemm of each rank's first rows vs the rows the sharded step produced"}


def bcast_alone_ms(comm, oz, b, world: int) -> float:
    """the bare NCCL broadcast of B through the library's communicator (max_panels=1, no rows), max over ranks"""
    import torch
    h = oz.create()
    dummy = torch.zeros(8, dtype=torch.float64, device="cuda")
    n = int(round(b.numel() ** 0.5))

    def fn():
        assert oz.sharded_gemm(h, comm, oz.op_n, oz.op_n, 0, n, n, 1.0, dummy, 1, b, n, 0.0, dummy, 1, oz.fp64_int8(NUM_SPLIT),
                               src=0, max_panels=1) == 0
    ms = timed_loop(fn, 5, 2, world)
    oz.destroy(h)
    return ms


def run_config4(oz, h, comm, rank, world, steps: int, peaks: dict) -> dict:
    """BASELINE config 4: 16384^3 fp64_int8_9 row-sharded over the ranks (strong scaling), B broadcast inside the step"""
    import torch
    n4 = int(os.environ.get("OZ_BENCH_N4", "16384"))
    s = NUM_SPLIT
    r0, rows = oz.row_block(n4, world, rank)
    g = torch.Generator(device="cuda").manual_seed(4321 + rank)
    a = 1.0 - torch.rand(max(rows, 1) * n4, dtype=torch.float64, device="cuda", generator=g)
    gb = torch.Generator(device="cuda").manual_seed(77)
    b = 1.0 - torch.rand(n4 * n4, dtype=torch.float64, device="cuda", generator=gb)   # only rank 0's content is used
    if rank != 0:
        b.zero_()
    c = torch.zeros(max(rows, 1) * n4, dtype=torch.float64, device="cuda")
    mode = oz.fp64_int8(s)
    panels = int(os.environ.get("OZIMMU_B200_BENCH_PANELS", "1"))

    def step():
        assert oz.sharded_gemm(h, comm, oz.op_n, oz.op_n, rows, n4, n4, 1.0, a, max(rows, 1), b, n4, 0.0, c, max(rows, 1),
                               mode, src=0, max_panels=panels) == 0
    ms, per_rank = timed_loop_ranks(step, steps, 1, world)
    parity = sharded_parity(oz, h, comm, rank, world, n4, n4, a, max(rows, 1), b, c, max(rows, 1), rows, mode) if world > 1 else None
    int8_ops_rank = s * (s + 1) / 2 * 2.0 * rows * n4 * n4
    out = {"workload": WORKLOAD.format(n=n4, s=s) + f"; rows {n4}/{world} per rank, B broadcast from rank 0 inside the step",
           "scaling": "strong", "value": 2.0 * n4 ** 3 / ms / 1e9, "unit": "TFLOP/s", "ms_per_step": ms, "steps": steps,
           "per_rank_ms": per_rank, "tc_fraction": int8_ops_rank / ms / 1e9 / (2.0 * peaks["bf16"]), "parity": parity}
    del a, b, c
    torch.cuda.empty_cache()
    return out


def run_ours(args) -> dict:
    import torch
    import ozimmu_b200 as oz

    rank, world, local = dist_setup(args.gpus)
    n, s = N_DEFAULT, NUM_SPLIT
    mode = oz.fp64_int8(s)
    a, b, c = make_inputs(n, rank)      # every rank: its own 8192-row block of A and C; B comes from rank 0
    if rank != 0:
        b.zero_()                        # only rank 0 owns B: everybody else must receive it in every step
    h = oz.create()
    L = oz.lib()
    comm = oz.comm_create()              # the library's own NCCL communicator (None on one GPU)
    # OZIMMU_B200_BENCH_PANELS=p > 1: B travels in p column panels on the communicator's stream and every panel of C
    # starts as soon as its columns have landed.  Default 1 (one broadcast overlapped by split(A), then one product
    # launch): measured faster on 2 x B200 -- 18.4 ms against 19.7 / 20.5 / 21.0 ms for 2 / 4 / 8 panels
    # (profiles/r2_sharded_probe_2gpu.txt): NCCL's broadcast CTAs hold SMs on both GPUs until both sides run.
    panels = int(os.environ.get("OZIMMU_B200_BENCH_PANELS", "1"))

    def step():
        assert oz.sharded_gemm(h, comm, oz.op_n, oz.op_n, n, n, n, 1.0, a, n, b, n, 0.0, c, n, mode, src=0,
                               max_panels=panels) == 0

    sampler = ClockSampler(local)
    launches0 = oz.launch_count()
    ms, per_rank = timed_loop_ranks(step, args.steps, args.warmup, world, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    launches = (oz.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    flop_step = 2.0 * n * n * n * world
    value = flop_step / ms / 1e9
    parity = sharded_parity(oz, h, comm, rank, world, n, n, a, n, b, c, n, n, mode) if world > 1 else None
    bcast_ms = bcast_alone_ms(comm, oz, b, world) if world > 1 else None
    # rank 0's block against cuBLAS DGEMM (at N > 1 the other ranks' B is only a receive buffer: step() has filled it)
    accuracy = accuracy_vs_cublas(a, b, c, n) if rank == 0 else None

    # ---- roofline of the dominant kernel (fused tcgen05 product+accumulate), CUDA events on its stream ----
    # (measured right after the device-resident loop, i.e. in the same power state; the host-operand loop follows)
    roof = None
    if rank == 0:
        pitch = int(L.ozk_slice_pitch(n))
        bits = int(L.ozk_bits_per_int8(n))
        a_sl = torch.empty(int(L.ozk_slices_bytes(n, n, s)), dtype=torch.int8, device="cuda")
        b_sl = torch.empty(int(L.ozk_slices_bytes(n, n, s)), dtype=torch.int8, device="cuda")
        amax = torch.empty(n, dtype=torch.float64, device="cuda")
        bmax = torch.empty(n, dtype=torch.float64, device="cuda")
        scr = torch.zeros(n, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        assert L.ozk_split_int8(a_sl.data_ptr(), pitch, amax.data_ptr(), scr.data_ptr(), n, n, a.data_ptr(), n, 1, s, bits, st) == 0
        assert L.ozk_split_int8(b_sl.data_ptr(), pitch, bmax.data_ptr(), scr.data_ptr(), n, n, b.data_ptr(), n, 0, s, bits, st) == 0
        durs = []
        reps = max(3, min(args.steps, 10))
        for i in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            assert L.ozk_gemm_i8_fused(n, n, n, a_sl.data_ptr(), b_sl.data_ptr(), pitch, amax.data_ptr(), bmax.data_ptr(),
                                       s, bits, 1.0, 0.0, c.data_ptr(), n, st) == 0
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                durs.append(e0.elapsed_time(e1))
        kms = sum(durs) / len(durs)
        # the practical int8 peak at the same clocks -- cuBLAS' own int8 GEMM (torch._int_mm, 8192^3) timed the
        # same way (one launch at a time); it has no FP64 epilogue to run
        cublas_int8 = None
        try:
            ai = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
            bi = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
            cd = []
            for i in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch._int_mm(ai, bi.t())
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    cd.append(e0.elapsed_time(e1))
            cms = sum(cd) / len(cd)
            cublas_int8 = {"tops": 2.0 * n * n * n / cms / 1e9, "ms": cms, "how": "torch._int_mm 8192^3 (cuBLASLt int8), launches timed one at a time"}
            del ai, bi
        except Exception as e:  # noqa: BLE001
            cublas_int8 = {"unavailable": str(e)[:120]}
        int8_ops = s * (s + 1) / 2 * 2.0 * n * n * n
        peaks = measured_peaks()
        peak = 2.0 * peaks["bf16"]
        traffic, traffic_src = ncu_traffic_bytes()
        roof = {"bound": "tensor", "achieved": int8_ops / kms / 1e9, "peak": peak, "unit": "TFLOP/s",
                "frac": int8_ops / kms / 1e9 / peak, "traffic": traffic,
                "frac_of_sustained_peak": (int8_ops / kms / 1e9 / (2.0 * peaks["bf16_sustained"])) if peaks["bf16_sustained"] else None,
                "hbm_gbs": (traffic / (kms * 1e-3) / 1e9) if traffic else None, "hbm_peak_gbs": peaks["hbm"],
                "kernel": "oz_gemm_pair_kernel<256, 128>", "kernel_ms": kms, "launches_timed": len(durs),
                "cublas_int8": cublas_int8,
                "frac_of_cublas_int8": (int8_ops / kms / 1e9 / cublas_int8["tops"]) if cublas_int8 and "tops" in cublas_int8 else None,
                "ops_per_launch": int8_ops,
                "note": f"int8 TOP/s; peak = 2 x bf16_tflops (burst: launches timed one at a time) of MEASURED_PEAKS.json "
                        f"({peaks['src']}); 2 x sustained bf16 = {2.0 * peaks['bf16_sustained'] if peaks['bf16_sustained'] else None}; "
                        f"nominal dense int8 4500; traffic = dram bytes of one launch from the committed ncu capture ({traffic_src})"}
        del a_sl, b_sl
    # ---- end to end: HOST operands through the C-ABI ------------------------------------------------
    ha = torch.empty(n * n, dtype=torch.float64).pin_memory(); ha.copy_(a)
    hb = None
    if rank == 0:
        hb = torch.empty(n * n, dtype=torch.float64).pin_memory(); hb.copy_(b)
    hc = torch.empty(n * n, dtype=torch.float64).pin_memory()
    if world == 1:
        def e2e_step():
            assert oz.gemm_host(h, oz.op_n, oz.op_n, n, n, n, 1.0, ha, n, hb, n, 0.0, hc, n, mode) == 0
        h2d, d2h = 2 * n * n * 8, n * n * 8
    else:
        def e2e_step():   # every rank: its rows of A up, its rows of C down; rank 0 also uploads B and forwards it over NVLink
            assert oz.sharded_gemm_host(h, comm, oz.op_n, oz.op_n, n, n, n, 1.0, ha, n, hb, n, 0.0, hc, n, mode, src=0) == 0
        h2d, d2h = (world + 1) * n * n * 8, world * n * n * 8
    e2e_steps = max(2, args.steps)      # a call is ~25 ms: the same K steps and W warm-ups as the device-resident loop
    e2e_ms, e2e_per_rank = timed_loop_ranks(e2e_step, e2e_steps, max(2, args.warmup), world)
    e2e = {"value": flop_step / e2e_ms / 1e9, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms, "steps": e2e_steps, "per_rank_ms": e2e_per_rank}
    if world > 1:
        # the host-operand result must be the device-resident one (same rows, same B)
        e2e["bit_identical_to_device_path"] = bool(torch.equal(hc.view(torch.int64), c.cpu().view(torch.int64)))
    del ha, hb, hc

    peaks = measured_peaks()
    int8_ops_rank = s * (s + 1) / 2 * 2.0 * n * n * n
    # BASELINE config 4 (16384^3 strong scaling) in the same run; OZ_BENCH_CONFIG4=0 skips it
    config4 = None
    if os.environ.get("OZ_BENCH_CONFIG4", "1") != "0":
        del a, c
        torch.cuda.empty_cache()
        config4 = run_config4(oz, h, comm, rank, world, max(2, min(args.steps, 3)), peaks)
    base = cpu_baseline() if (rank == 0 and world == 1) else None
    oz.destroy(h)
    if comm is not None:
        comm.destroy()
    out = {"metric": "FP64-equiv TFLOP/s at fp64_int8_9, 8192^3 (2*m*n*k/t)", "value": value, "unit": "TFLOP/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "int8 (int32 products, f64 accumulation; f64 in/out)",
           "data": "synthetic urand01 (0,1], seeded torch.rand on device",
           "config": {"workload": WORKLOAD.format(n=n, s=s),
                      "sharding": "none" if world == 1 else f"one {n}-row block of A and C per GPU (global m={n * world}), B broadcast "
                                  f"from rank 0 inside every step (library NCCL communicator, {panels} broadcast panel(s), split(A) overlaps the transfer)",
                      "timing": "CUDA events on the call stream, inputs (1 GiB) larger than L2 (126 MB) so no flush",
                      "parallelism": f"rows{world}"},
           "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": base,
           "per_rank_ms": per_rank, "bcast_ms": bcast_ms, "parity": parity, "accuracy": accuracy,
           "tc_fraction": int8_ops_rank / ms / 1e9 / (2.0 * peaks["bf16"]), "config4": config4}
    return out if rank == 0 else {}


# ------------------------------------------------------------------------------------------------
def run_reference(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import oracle_lib
    n, s = N_DEFAULT, NUM_SPLIT
    if oracle_lib.reference() is None:
        return reference_cpu_port(args, "oracle/_ref/libozref.so not built")
    from gpu_util import Reference
    a, b, c = make_inputs(n, 0)
    try:
        ref = Reference()
    except Exception as e:  # noqa: BLE001
        return reference_cpu_port(args, f"reference failed to initialise: {e}")

    def step():
        ref.gemm(0, 0, n, n, n, 1.0, a, n, b, n, 0.0, c, n, s - 1)

    sampler = ClockSampler(local)
    ms = timed_loop(step, args.steps, args.warmup, 1, sampler)
    clocks = sampler.stop()
    accuracy = accuracy_vs_cublas(a, b, c, n)
    flop = 2.0 * n * n * n
    ha = torch.empty(n * n, dtype=torch.float64).pin_memory(); ha.copy_(a)
    hb = torch.empty(n * n, dtype=torch.float64).pin_memory(); hb.copy_(b)
    hc = torch.empty(n * n, dtype=torch.float64).pin_memory()
    da, db, dc = torch.empty_like(a), torch.empty_like(b), torch.empty_like(c)

    def e2e_step():  # what an application of the reference does: cudaMemcpy in, cublasDgemm (intercepted), cudaMemcpy out
        da.copy_(ha, non_blocking=True)
        db.copy_(hb, non_blocking=True)
        ref.gemm(0, 0, n, n, n, 1.0, da, n, db, n, 0.0, dc, n, s - 1)
        hc.copy_(dc, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps = max(2, args.steps)
    e2e_ms = timed_loop(e2e_step, e2e_steps, max(1, args.warmup), 1)
    ref.close()
    value = flop / ms / 1e9
    return {"impl": "reference", "metric": "FP64-equiv TFLOP/s at fp64_int8_9, 8192^3 (2*m*n*k/t)", "value": value,
            "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 (int32 products, f64 accumulation; f64 in/out)",
            "data": "synthetic urand01 (0,1], seeded torch.rand on device",
            "config": {"workload": WORKLOAD.format(n=n, s=s),
                       "timing": "CUDA events, inputs larger than L2"},
            "clocks": clocks, "accuracy": accuracy,
            "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": 0, "kind": "reference",
                             "sample": "oracle/_ref/libozref.so: the unmodified reference ozIMMU (its own sources built for sm_100), "
                                       "whole workload per step ON THE GPU -- the reference has no CPU implementation of this path"},
            "e2e": {"value": flop / e2e_ms / 1e9, "unit": "TFLOP/s", "h2d_bytes_per_step": 2 * n * n * 8,
                    "d2h_bytes_per_step": n * n * 8, "ms_per_step": e2e_ms, "steps": e2e_steps}}


def reference_cpu_port(args, why: str) -> dict:
    """fallback arm: the scalar CPU restatement on a bounded sample"""
    import numpy as np
    import oracle_lib
    m = 384
    rng = np.random.default_rng(0)
    x, y = 1.0 - rng.random(m * m), 1.0 - rng.random(m * m)
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_lib.oracle_gemm(0, 0, m, m, m, 1.0, x, m, y, m, 0.0, np.zeros(m * m), m, NUM_SPLIT)
    dt = (time.perf_counter() - t0) / steps
    v = 2.0 * m ** 3 / dt / 1e12
    return {"impl": "reference", "metric": "FP64-equiv TFLOP/s at fp64_int8_9, 8192^3 (2*m*n*k/t)", "value": v, "unit": "TFLOP/s",
            "n_gpus": 1, "steps": steps, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8 (int32 products, f64 accumulation; f64 in/out)", "data": "synthetic urand01",
            "config": {"workload": f"bounded sample {m}^3 fp64_int8_{NUM_SPLIT} of the 8192^3 workload ({why})"},
            "cpu_baseline": {"value": v, "unit": "TFLOP/s", "cores": 1, "kind": "port", "sample": f"oracle/oz_oracle.c {m}^3 x{steps}"},
            "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out:
        print(json.dumps(out), flush=True)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    main()
