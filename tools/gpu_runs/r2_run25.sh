# round 2, call 25 (1 GPU): SM ingest rate of linear bulk copies by copy size, ring depth, CTA count and working set
# (what bounds the small tiles: 2 x 8 KB per stage on 128 SMs reach 42 B/clk/SM in the kernel)
mkdir -p gpurun_out
for bin in bulk_pair bulk_pair_8k12 bulk_pair_8k5 bulk_pair_4k24; do
  for args in "0 4096" "128 4096" "128 1152" "64 1152" "0 1152"; do
    timeout 60 tools/ubench/$bin $args 2>&1 | grep -v "remote-barrier=1" | tee -a gpurun_out/r2_ubench_bulk_sizes.txt
  done
done
