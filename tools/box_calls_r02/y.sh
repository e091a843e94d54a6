#!/bin/bash
# Round-2 (8 GPUs, ONE process): NVLink counters of the multicast GEMM kernel with seven peers.
set -o pipefail
O=gpurun_out/r02y; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
M="gpu__time_duration.sum,dram__bytes_write.sum,nvltx__bytes.sum,nvlrx__bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 400 ncu --metrics $M --clock-control none -k regex:gemm_tc_kernel -s 8 -c 8 --csv --log-file $O/ncu_multicast_8gpu.csv ./build/gemm_multi_b200 16384 8 0 > $O/ncu_multicast_8gpu.log 2>&1; echo "ncu rc=$?"; tail -4 $O/ncu_multicast_8gpu.log
python - <<PY
import csv
rows = [r for r in csv.reader(open("$O/ncu_multicast_8gpu.csv")) if len(r) > 10]
hdr = rows[0] if rows else []
for r in rows[1:]:
    d = dict(zip(hdr, r))
    print(d.get("ID"), d.get("Device", "")[:12], d.get("Metric Name"), d.get("Metric Unit"), d.get("Metric Value"))
PY
