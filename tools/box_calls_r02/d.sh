#!/bin/bash
# Round-2 fourth box call: the in-kernel bf16 split, the pusher warp and the multi-GPU group on one GPU, the reference's
# verifying benchmark binaries, A/B of the fp32 modes.
O=gpurun_out/r02d; mkdir -p $O
timeout 600 python -m pytest tests/test_split16_gpu.py tests/test_multicast_gpu.py tests/test_multi_group_gpu.py tests/test_beta_zero_nan_gpu.py -q -x > $O/pytest_new.txt 2>&1; echo "new rc=$?"; tail -25 $O/pytest_new.txt
timeout 900 python -m pytest tests/test_zz_reference_unittests_gpu.py -q -k "benchmark_harness" > $O/pytest_refbench.txt 2>&1; echo "refbench rc=$?"; tail -25 $O/pytest_refbench.txt
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ref_parity_gpu.py tests/test_full_size_gpu.py tests/test_host_path_gpu.py -q > $O/pytest_gemm.txt 2>&1; echo "gemm rc=$?"; tail -15 $O/pytest_gemm.txt
timeout 300 python tools/ab_variants.py --workload sgemm_splitk --variants default,split16_off,cublas --burst-steps 5 --rounds 5 > $O/ab_splitk.jsonl 2> $O/ab_splitk.err; cat $O/ab_splitk.jsonl
timeout 300 python tools/ab_variants.py --workload sgemm1024 --variants default,split16_off,cublas --burst-steps 200 --rounds 5 > $O/ab_sgemm1024.jsonl 2> $O/ab_1024.err; cat $O/ab_sgemm1024.jsonl
timeout 300 python tools/ab_variants.py --workload bf16gemm_batched --variants default,static,cublas > $O/ab_cfg4.jsonl 2> $O/ab_cfg4.err; cat $O/ab_cfg4.jsonl
tail -3 $O/*.err
