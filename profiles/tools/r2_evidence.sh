#!/usr/bin/env bash
# Round-2 evidence, one gpurun call (1 GPU): tests, both bench arms, the other precisions, shape sweeps, launch list,
# ncu --set full captures of the traversal kernels, compute-sanitizer, the random-gather ceiling.
#   gpurun --timeout 3000 -- 'bash profiles/tools/r2_evidence.sh <tag>'     -> gpurun_out/<tag>_*
set -u
T=${1:-r2}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py > $O/${T}_bench_hnsw.json 2> $O/${T}_bench_hnsw.err
python bench.py --workload quantized --precision int8 --overlap 4 > $O/${T}_bench_int8.json 2> $O/${T}_bench_int8.err
python bench.py --workload quantized --precision float16 --overlap 4 > $O/${T}_bench_f16.json 2> $O/${T}_bench_f16.err
python bench.py --workload flat > $O/${T}_bench_flat.json 2> $O/${T}_bench_flat.err
python profiles/tools/sweep_tuning.py 4,192,0,4 8,192,0,8 4,192,0,8 > $O/${T}_sweep_f32.log 2>&1
PREC=int8 OV=4 python profiles/tools/sweep_tuning.py 4,64,0,4 8,64,0,8 16,64,0,16 8,64,0,16 > $O/${T}_sweep_int8.log 2>&1
PREC=float16 OV=4 python profiles/tools/sweep_tuning.py 4,64,0,4 4,128,0,4 8,64,0,8 4,64,0,8 > $O/${T}_sweep_f16.log 2>&1
cat $O/${T}_sweep_f32.log $O/${T}_sweep_int8.log $O/${T}_sweep_f16.log
# per-launch device times of one short bench run (share of the step per kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-single-call --sustain-seconds 0 > $O/${T}_launches_bench.log 2>&1
# full captures: one launch each (the throughput shape, 1024 queries; int8 also 4096 queries = every SM slot taken)
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 4 -c 1 -f -o $O/${T}_f32 \
  python profiles/tools/sweep_tuning.py 4,192,0,4 > $O/${T}_ncu_f32.log 2>&1
PREC=int8 OV=1 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_fast -s 4 -c 1 -f -o $O/${T}_int8 \
  python profiles/tools/sweep_tuning.py 8,64,0,8 > $O/${T}_ncu_int8.log 2>&1
B=4096 PREC=int8 OV=1 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_fast -s 4 -c 1 -f -o $O/${T}_int8_b4096 \
  python profiles/tools/sweep_tuning.py 8,64,0,8 > $O/${T}_ncu_int8_b4096.log 2>&1
PREC=float16 OV=1 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 4 -c 1 -f -o $O/${T}_f16 \
  python profiles/tools/sweep_tuning.py 4,64,0,4 > $O/${T}_ncu_f16.log 2>&1
# compute-sanitizer on the C1 shape (construction + traversal + flat + shard group)
timeout 900 compute-sanitizer --tool memcheck python profiles/tools/sanitize_c1.py > $O/${T}_sanitizer_memcheck.log 2>&1; tail -3 $O/${T}_sanitizer_memcheck.log
N=1200 timeout 1200 compute-sanitizer --tool racecheck python profiles/tools/sanitize_c1.py > $O/${T}_sanitizer_racecheck.log 2>&1; tail -3 $O/${T}_sanitizer_racecheck.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather_probe profiles/tools/gather_probe.cu && /tmp/gather_probe > $O/${T}_gather_probe.log 2>&1
ls -la $O | grep ${T}_ | awk '{print $5, $9}'
