#!/usr/bin/env bash
# gpurun --timeout 1500 -- 'bash profiles/tools/r2_check.sh <tag> [sections]'   sections: tests cold bench ref flat
set -u
T=${1:-x1}; shift || true
S=" ${*:-tests cold bench} "
O=gpurun_out
mkdir -p $O
has() { [[ "$S" == *" $1 "* ]]; }
if has tests; then timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "tests rc=$?"; tail -4 $O/${T}_pytest.log; fi
if has cold; then
  timeout 300 python profiles/tools/cold_start.py 1000000 float32 > $O/${T}_cold_start_1m_f32.json 2> $O/${T}_cold_start_1m_f32.err
  echo "cold1 rc=$?"; cat $O/${T}_cold_start_1m_f32.json
  timeout 500 python profiles/tools/cold_start.py 10000000 int8 > $O/${T}_cold_start_10m_int8.json 2> $O/${T}_cold_start_10m_int8.err
  echo "cold10 rc=$?"; cat $O/${T}_cold_start_10m_int8.json
fi
if has ref; then timeout 900 python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "ref rc=$?"; cut -c1-400 $O/${T}_bench_reference.json; fi
if has bench; then timeout 1200 python bench.py > $O/${T}_bench_hnsw.json 2> $O/${T}_bench_hnsw.err; echo "bench rc=$?"; cut -c1-1200 $O/${T}_bench_hnsw.json; fi
if has flat; then timeout 400 python bench.py --workload flat > $O/${T}_bench_flat.json 2> $O/${T}_bench_flat.err; echo "flat rc=$?"; cut -c1-700 $O/${T}_bench_flat.json; fi
