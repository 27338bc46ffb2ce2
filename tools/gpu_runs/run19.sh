mkdir -p gpurun_out
(time timeout 1500 oracle/_ref/main.test.ours ci_test) > gpurun_out/ref_ci_test_ours.csv 2> gpurun_out/ref_ci_test_ours.err; echo "ci_test rc=$?"; tail -3 gpurun_out/ref_ci_test_ours.csv; grep -c FAILED gpurun_out/ref_ci_test_ours.csv; tail -5 gpurun_out/ref_ci_test_ours.err
(timeout 300 oracle/_ref/main.test.ours urand01 dgemm seq 8192 8192 1 fp64_int8_9 dgemm) 2>&1 | tail -3
(timeout 300 oracle/_ref/main.test.ours urand01 zgemm seq 4096 4096 1 fp64_int8_9 dgemm) 2>&1 | tail -3
