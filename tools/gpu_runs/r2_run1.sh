# round 2, call 1 (1 GPU): all GPU tests after the kernel generalisation (tile widths, one-tile launches, complex in one
# launch, device scalars, batched auto), tile-width sweep, e2e schedule sweep (one-tile launches vs persistent vs queue)
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -x) > gpurun_out/r2_t1.log 2>&1; echo "pytest gpu rc=$?"; tail -25 gpurun_out/r2_t1.log
for rep in 1 2; do
  timeout 300 python tools/perf_probe.py 8192 9 --iters 8 --shapes p256,p240,p224,p208,p192 --no-extras 2>&1 | grep ozimmu_b200
done | tee gpurun_out/r2_sweep_tile_width.txt
timeout 200 python tools/perf_probe.py 4096 9 --iters 8 --shapes p256,p224,p192,p128,00 --zgemm --no-extras 2>&1 | tee gpurun_out/r2_perf_4096.txt
timeout 600 python tools/e2e_probe.py 8192 \
  768:768:0:OZIMMU_B200_E2E_ONE_TILE=0 768:768:0:OZIMMU_B200_E2E_ONE_TILE=1 512:512:0:OZIMMU_B200_E2E_ONE_TILE=1 \
  512:512:0:OZIMMU_B200_E2E_ONE_TILE=1,OZIMMU_B200_E2E_STREAMS=6 512:512:0:OZIMMU_B200_E2E_ONE_TILE=1,OZIMMU_B200_E2E_RECT_TILES=32 \
  768:768:0:OZIMMU_B200_E2E_ONE_TILE=1,OZIMMU_B200_E2E_RECT_TILES=24 1024:1024:0:OZIMMU_B200_E2E_ONE_TILE=1,OZIMMU_B200_E2E_RECT_TILES=32 \
  768:768:0:OZIMMU_B200_E2E_ONE_TILE=0 768:768:0:OZIMMU_B200_E2E_ONE_TILE=1 \
  768:768:0:OZIMMU_B200_E2E_QUEUE=1,OZIMMU_B200_E2E_QUEUE_RESERVE_SMS=16 \
  768:768:0:OZIMMU_B200_E2E_QUEUE=1,OZIMMU_B200_E2E_QUEUE_RESERVE_SMS=16,OZIMMU_B200_E2E_QUEUE_JOIN=1 \
  768:768:0:OZIMMU_B200_E2E_QUEUE=1,OZIMMU_B200_E2E_QUEUE_RESERVE_SMS=24,OZIMMU_B200_E2E_QUEUE_JOIN=1 \
  2>&1 | tee gpurun_out/r2_e2e_sweep.txt
