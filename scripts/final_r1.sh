#!/bin/bash
# round-end measurement pass (one GPU), ordered by priority; every step has its own time limit
mkdir -p gpurun_out
timeout 120 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on \
    -k regex:"filter_tiles|eval_functions|general_ia_small|emit_ia|hash_insert|rank_reps" --launch-skip 18 --launch-count 6 \
    -f -o gpurun_out/ncu_r1_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 40 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; echo "C3 rc=$?"
timeout 40 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err; echo "C4 rc=$?"
timeout 60 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
cat gpurun_out/bench_final.json | cut -c1-1500
ls -la gpurun_out | head -30
