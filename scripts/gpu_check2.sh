#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_final.log
timeout 40 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.json 2>/dev/null; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value']/1e9, d['ms_per_step'], d['stage_ms']['general'])"
