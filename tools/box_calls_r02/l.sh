#!/bin/bash
# Round-2: how much of the fp32 mainloop waits for the chunk fold (timing probes with longer chunks), skinny-M regression.
set -o pipefail
O=gpurun_out/r02l; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 300 python tools/ab_variants.py --workload sgemm8192 --variants default,chunk32,chunk64,chunk_inf --burst-steps 5 --rounds 4 > $O/ab_chunk_8192.jsonl 2> $O/ab_chunk.err; cat $O/ab_chunk_8192.jsonl
timeout 300 python tools/ab_variants.py --shape f32,64,147,13225 --variants default,split16_off,static,nopdl --burst-steps 100 --rounds 3 > $O/ab_skinny1.jsonl 2> $O/ab_skinny1.err; cat $O/ab_skinny1.jsonl
timeout 300 python tools/ab_variants.py --shape f32,64,64,12544 --variants default,split16_off,static,nopdl --burst-steps 100 --rounds 3 > $O/ab_skinny2.jsonl 2> $O/ab_skinny2.err; cat $O/ab_skinny2.jsonl
timeout 300 python tools/ab_variants.py --shape f32,128,256,25088 --variants default,split16_off,static,nopdl --burst-steps 100 --rounds 3 > $O/ab_skinny3.jsonl 2> $O/ab_skinny3.err; cat $O/ab_skinny3.jsonl
