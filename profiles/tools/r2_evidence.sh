#!/usr/bin/env bash
# Round-2 evidence on one GPU box:  gpurun --timeout 3000 -- 'bash profiles/tools/r2_evidence.sh <tag> <sections>'
# sections (any of): tests bench sweep launches ncu flatncu cold sanitize gather  -> gpurun_out/<tag>_*
# ncu reports are exported to CSV on the box and deleted (gpurun copies back at most 64 MiB).
set -u
T=${1:-r2}; shift || true
S=" ${*:-tests bench sweep launches ncu sanitize gather} "
O=gpurun_out
mkdir -p $O
has() { [[ "$S" == *" $1 "* ]]; }
if has tests; then python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log; fi
if has bench; then
  python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
  python bench.py > $O/${T}_bench_hnsw.json 2> $O/${T}_bench_hnsw.err
  python bench.py --workload quantized --precision int8 --overlap 4 > $O/${T}_bench_int8.json 2> $O/${T}_bench_int8.err
  python bench.py --workload quantized --precision float16 --overlap 4 > $O/${T}_bench_f16.json 2> $O/${T}_bench_f16.err
  python bench.py --workload flat > $O/${T}_bench_flat.json 2> $O/${T}_bench_flat.err
  python bench.py --workload hybrid > $O/${T}_bench_hybrid.json 2> $O/${T}_bench_hybrid.err
fi
if has flatncu; then  # the two tensor passes of one flat step under ncu --set full, and the launch list of a short flat run
  ncu --set full --clock-control none --import-source on -k regex:flat_tc2_kernel -s 8 -c 2 -f -o /tmp/${T}_flat \
    python bench.py --workload flat --steps 2 --warmup 3 --no-cpu-baseline --sustain-seconds 0 > $O/${T}_ncu_flat.log 2>&1
  ncu -i /tmp/${T}_flat.ncu-rep --page raw --csv > $O/${T}_flat_tc2_kernel_raw.csv 2>/dev/null
  rm -f /tmp/${T}_flat.ncu-rep
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_flat_launches.csv \
    python bench.py --workload flat --steps 2 --warmup 3 --no-cpu-baseline --sustain-seconds 0 > $O/${T}_flat_launches_bench.log 2>&1
fi
if has cold; then
  python profiles/tools/cold_start.py 1000000 float32 > $O/${T}_cold_start_1m_f32.json 2> $O/${T}_cold_start_1m_f32.err
  python profiles/tools/cold_start.py 10000000 int8 > $O/${T}_cold_start_10m_int8.json 2> $O/${T}_cold_start_10m_int8.err
  cat $O/${T}_cold_start_1m_f32.json $O/${T}_cold_start_10m_int8.json
fi
if has sweep; then
  python profiles/tools/sweep_tuning.py 4,192,0,4 8,192,0,8 4,192,0,8 > $O/${T}_sweep_f32.log 2>&1
  PREC=int8 OV=4 python profiles/tools/sweep_tuning.py 4,64,0,4 8,64,0,8 16,64,0,16 8,64,0,16 > $O/${T}_sweep_int8.log 2>&1
  PREC=float16 OV=4 python profiles/tools/sweep_tuning.py 4,64,0,4 4,128,0,4 8,64,0,8 4,64,0,8 > $O/${T}_sweep_f16.log 2>&1
  cat $O/${T}_sweep_f32.log $O/${T}_sweep_int8.log $O/${T}_sweep_f16.log
fi
if has launches; then  # per-launch device times of one short bench run (share of the step per kernel)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-single-call --sustain-seconds 0 > $O/${T}_launches_bench.log 2>&1
fi
capture() {  # name, kernel regex, env..., then the sweep arguments after --
  local name=$1 regex=$2; shift 2
  local envs=()
  while [[ "$1" != "--" ]]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" ncu --set full --clock-control none --import-source on -k regex:$regex -s 4 -c 1 -f -o /tmp/${T}_$name \
    python profiles/tools/sweep_tuning.py "$@" > $O/${T}_ncu_$name.log 2>&1
  ncu -i /tmp/${T}_$name.ncu-rep --page raw --csv > $O/${T}_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_$name.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${T}_${name}_src.csv 2>/dev/null
  python profiles/tools/ncu_by_line.py /tmp/${T}_${name}_src.csv 80 > $O/${T}_${name}_by_line.csv 2>/dev/null
  local nq=${NQ:-1024}
  python profiles/tools/ncu_by_function.py /tmp/${T}_${name}_src.csv $O/${T}_${name}_raw.csv $nq 7384 > $O/${T}_${name}_by_function.txt 2>&1
  tail -n +1 $O/${T}_${name}_by_function.txt | head -14
  rm -f /tmp/${T}_$name.ncu-rep
}
if has ncu; then  # one launch each of the throughput shapes; int8 also with every SM slot taken (4096 queries)
  capture f32 hnsw_search_kernel OV=1 -- 4,192,0,4
  capture f32_alone hnsw_search_kernel OV=1 -- 8,192,0,8
  capture int8 hnsw_search_fast PREC=int8 OV=1 -- 8,64,0,8
  NQ=4096 capture int8_b4096 hnsw_search_fast PREC=int8 OV=1 B=4096 -- 8,64,0,8
  capture f16 hnsw_search_kernel PREC=float16 OV=1 -- 4,128,0,4
fi
if has sanitize; then  # compute-sanitizer on the C1 shape (construction + traversal + flat + shard group)
  timeout 900 compute-sanitizer --tool memcheck python profiles/tools/sanitize_c1.py > $O/${T}_sanitizer_memcheck.log 2>&1; tail -3 $O/${T}_sanitizer_memcheck.log
  N=1200 timeout 1200 compute-sanitizer --tool racecheck python profiles/tools/sanitize_c1.py > $O/${T}_sanitizer_racecheck.log 2>&1; tail -3 $O/${T}_sanitizer_racecheck.log
fi
if has gather; then
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather_probe profiles/tools/gather_probe.cu && /tmp/gather_probe > $O/${T}_gather_probe.log 2>&1
fi
du -sh $O; ls -la $O | grep ${T}_ | awk '{print $5, $9}'
