#!/bin/bash
# Last seconds of a round's GPU budget: the reference's own unit-test and benchmark binaries (built unchanged by
# portblas_b200/build_host.py), all at once, each under its own short timeout; logs come back in gpurun_out/ref_quick/.
O=gpurun_out/ref_quick; mkdir -p $O
T=${1:-20}
printf 'n,n,1024,1024,1024,1.5,0.5\nt,n,512,333,257,1,0\n' > $O/gemm.csv
timeout $T build/ref_unittest_blas3_gemm_tall_skinny_test > $O/tall_skinny.log 2>&1 &
timeout $T build/ref_unittest_blas3_gemm_test --gtest_filter='*alloc_usm*' > $O/gemm.log 2>&1 &
timeout $T build/ref_unittest_blas3_symm_test --gtest_filter='*alloc_usm*' > $O/symm.log 2>&1 &
timeout $T build/ref_unittest_blas3_trsm_test --gtest_filter='*alloc_usm*' > $O/trsm.log 2>&1 &
timeout $T build/ref_unittest_blas3_gemm_batched_test --gtest_filter='*FloatFloat.test/alloc_usm*:*HalfFloat.test/alloc_usm*' > $O/batched.log 2>&1 &
timeout $T build/ref_unittest_blas3_gemm_test --gtest_filter='*alloc_buf*Float*' > $O/gemm_buf.log 2>&1 &
timeout $T build/ref_bench_gemm --csv-param $O/gemm.csv --benchmark_min_time=0.2 > $O/bench_gemm.log 2>&1 &
timeout $T build/ref_sample_gemm > $O/sample.log 2>&1 &
wait
for f in $O/*.log; do echo "== $f: $(grep -c '\[       OK \]' $f) ok, $(grep -c '\[  FAILED  \]' $f) failed"; tail -2 $f; done
