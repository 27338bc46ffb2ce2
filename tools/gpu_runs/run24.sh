python -c "import torch; torch.zeros(1).cuda()"
for n in 1024 2048; do timeout 200 python tools/perf_probe.py $n 9 --iters 50 --shapes 00,p128,p192,p256,11 2>&1 | head -12; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 python tools/perf_probe.py 1024 9 --iters 1 2>&1 | grep -E "^\s+(void|oz)|gpu__time" | head -40
