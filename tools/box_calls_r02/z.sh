#!/bin/bash
# Round-2: full GPU suite (three xdist workers sharing the GPU) + smoke on the last tree of the round.
set -o pipefail
O=gpurun_out/r02z; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.txt
timeout 560 python -m pytest tests -m gpu -q -n 3 > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.txt
