#!/bin/bash
# round-end measurement pass (one GPU): tests, bench lines, ncu launch list, one --set full capture
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_final.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 120 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; echo "C3 rc=$?"
timeout 120 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err; echo "C4 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"filter_tiles|eval_functions|general_ia_small|emit_ia|hash_insert|write_verts" --launch-skip 18 --launch-count 6 \
    -f -o gpurun_out/ncu_r1_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/bench_final.json gpurun_out/bench_C3.json gpurun_out/bench_C4.json
ls -la gpurun_out
