# round 2, call 19 (1 GPU): host-operand entry, block edge x tail pieces with A first (two repetitions)
mkdir -p gpurun_out
args=""
for rep in 1 2; do for b in 768 1024 1280 1536; do for p in 1 2 4; do args="$args $b:$b:0:OZIMMU_B200_E2E_TAIL_PIECES=$p"; done; done; done
timeout 900 python tools/e2e_probe.py 8192 $args 2>&1 | tee gpurun_out/r2_e2e_tail_sweep.txt
