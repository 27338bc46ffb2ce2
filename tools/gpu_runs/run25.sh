python -c "import torch; torch.zeros(1).cuda()"
for ls in 2 0; do for n in 1024 2048; do echo -n "LOCKSTEP=$ls "; OZIMMU_B200_LOCKSTEP=$ls timeout 200 python tools/perf_probe.py $n 9 --iters 50 --shapes 00,p128,p192,p256 2>&1 | head -4; done; done
