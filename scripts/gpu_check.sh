#!/bin/bash
# quick GPU pass: parity tests (optionally a subset via $1) + one bench line
mkdir -p gpurun_out
timeout ${2:-400} python -m pytest ${1:-tests} -m gpu -x -q > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_check.log
if [ -z "$1" ]; then
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/bench_check.json')); print(d['value']/1e9, 'G tets/s', d['ms_per_step'], 'ms'); print(d['stage_ms'])"
fi
