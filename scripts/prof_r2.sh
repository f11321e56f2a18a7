#!/bin/bash
# bench + ncu launch list + ncu --set full of the pipeline kernels (one GPU).  usage: scripts/prof_r2.sh <tag> [config]
tag=${1:-x}; cfg=${2:-C2}
mkdir -p gpurun_out
timeout 200 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err; echo "bench rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_$tag.csv \
    python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on \
    -k regex:"eval_kernel|filter_classify|general_ia_small|general_ia_mid|emit_kernel|insert_kernel|rank_verts|faces_kernel|filter_mi|highest|emit_mi|hash_insert|rank_reps|write_verts|general_mi" --launch-skip 24 --launch-count 8 \
    -f -o gpurun_out/ncu_r2_$tag python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
cut -c1-1800 gpurun_out/bench_r2_$tag.json
