mkdir -p gpurun_out
: > gpurun_out/nccl_bcast.log
run() { env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 tools/ubench/nccl_bcast.py 2>&1 | grep "^world" >> gpurun_out/nccl_bcast.log; }
run NCCL_DEBUG=WARN
run NCCL_MIN_NCHANNELS=16
run NCCL_MIN_NCHANNELS=32
run NCCL_MIN_NCHANNELS=32 NCCL_NTHREADS=512
run NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=24
run NCCL_ALGO=Ring NCCL_MIN_NCHANNELS=32
run NCCL_P2P_USE_CUDA_MEMCPY=1
cat gpurun_out/nccl_bcast.log
