#!/bin/bash
# Round-2 (2 GPUs, ONE process): NVLink counters of the multicast GEMM kernel under ncu (samples/gemm_multi_b200.cpp drives
# both devices from one thread, so this is a single-process capture).
set -o pipefail
O=gpurun_out/r02x; mkdir -p $O
python -m portblas_b200.build > /dev/null || { echo "BUILD BROKEN"; exit 9; }
ncu --query-metrics 2>/dev/null | grep -i -E "nvl|fabric|peer" | head -60 > $O/nvlink_metrics_available.txt; wc -l $O/nvlink_metrics_available.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,nvltx__bytes.sum,nvlrx__bytes.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_tc_kernel -c 10 --csv --log-file $O/ncu_multicast_2gpu.csv ./build/gemm_multi_b200 8192 2 0 > $O/ncu_multicast_2gpu.log 2>&1; echo "ncu rc=$?"; tail -5 $O/ncu_multicast_2gpu.log
python - <<PY
import csv
rows = [r for r in csv.reader(open("$O/ncu_multicast_2gpu.csv")) if len(r) > 10]
hdr = rows[0] if rows else []
for r in rows[1:]:
    d = dict(zip(hdr, r))
    print(d.get("ID"), d.get("Device", "")[:12], d.get("Metric Name"), d.get("Metric Unit"), d.get("Metric Value"))
PY
