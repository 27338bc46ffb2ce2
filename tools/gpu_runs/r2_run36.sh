# round 2, call 36 (1 GPU): is a bulk copy's cost per copy or per byte?  (cheap addresses in the issuing loop; one copy
# of 2 x chunk against two copies of chunk per stage)
mkdir -p gpurun_out
for bin in bulk_pair bulk_pair_one32k bulk_pair_8k12 bulk_pair_one16k bulk_pair_4k12; do
  timeout 20 tools/ubench/$bin 0 4096 2>&1 | grep -v "remote-barrier=1" | sed "s/^/$bin: /" | tee -a gpurun_out/r2_ubench_bulk_per_copy.txt
  echo "$bin rc=$?"
done
