#!/usr/bin/env bash
# gpurun --timeout 1500 -- 'bash profiles/tools/r2_flat_dyn.sh <tag>'
# flat scan with the dynamic tile schedule: parity tests, A/B against the static schedule in one process, the flat bench
# line, one ncu --set full capture of the two tensor passes, and the cold-start timings (arena + graph sidecar -> HBM).
set -u
T=${1:-w1}
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_flat_tc.py tests/test_gpu_parity.py -x -q -k "flat or prefilter or pair or tensor" > $O/${T}_flat_tests.log 2>&1
echo "flat tests rc=$?"; tail -3 $O/${T}_flat_tests.log
timeout 300 python profiles/tools/flat_ab.py > $O/${T}_flat_ab.log 2>&1
echo "ab rc=$?"; tail -4 $O/${T}_flat_ab.log
timeout 400 python bench.py --workload flat > $O/${T}_bench_flat.json 2> $O/${T}_bench_flat.err
echo "bench rc=$?"; cut -c1-600 $O/${T}_bench_flat.json
timeout 500 ncu --set full --clock-control none --import-source on -k regex:flat_tc2_kernel -s 8 -c 2 -f -o /tmp/${T}_flat \
  python bench.py --workload flat --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_flat.log 2>&1
ncu -i /tmp/${T}_flat.ncu-rep --page raw --csv > $O/${T}_flat_tc2_kernel_raw.csv 2>/dev/null
echo "ncu rc=$?"; wc -c $O/${T}_flat_tc2_kernel_raw.csv
timeout 300 python profiles/tools/cold_start.py 1000000 float32 > $O/${T}_cold_start_1m_f32.json 2> $O/${T}_cold_start_1m_f32.err
echo "cold1 rc=$?"; cat $O/${T}_cold_start_1m_f32.json
df -h /tmp | tail -1
FREE=$(df --output=avail -k /tmp | tail -1)
if [ "$FREE" -gt 20000000 ]; then
  timeout 500 python profiles/tools/cold_start.py 10000000 int8 > $O/${T}_cold_start_10m_int8.json 2> $O/${T}_cold_start_10m_int8.err
  echo "cold10 rc=$?"; cat $O/${T}_cold_start_10m_int8.json
fi
