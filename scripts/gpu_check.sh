#!/bin/bash
# quick GPU pass: parity tests (optionally a subset via $1) + one bench line
mkdir -p gpurun_out
timeout 400 python -m pytest ${1:-tests} -m gpu -x -q > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_check.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_check.json'))
print(d['value']/1e9, 'G tets/s', d['ms_per_step'], 'ms', d['device_ms_per_step']); print(d['stage_ms'])
PY
timeout 120 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_check_C3.json 2>&1; echo "C3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_check_C3.json'))
print('C3', d['value']/1e9, 'G tets/s', d['ms_per_step'], 'ms'); print(d['stage_ms'])
PY
timeout 120 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_check_C4.json 2>&1; echo "C4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_check_C4.json'))
print('C4', d['value']/1e9, 'G tets/s', d['ms_per_step'], 'ms'); print(d['stage_ms'])
PY
