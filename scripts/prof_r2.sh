#!/bin/bash
# bench + ncu launch list + ncu --set full of the pipeline kernels (one GPU).  usage: scripts/prof_r2.sh <tag> [config]
# The .ncu-rep is summarised on the box (scripts/ncu_summary.py, scripts/ncu_traffic.py) and deleted unless KEEP_REP=1:
# gpurun brings back at most 64 MiB.
tag=${1:-x}; cfg=${2:-C2}
mkdir -p gpurun_out
timeout 200 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err; echo "bench rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_$tag.csv \
    python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
skip=24; cnt=8; if [ "$cfg" = "C3" ]; then skip=38; cnt=13; fi
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"eval_kernel|eval_mi|filter_classify|general_ia_small|general_ia_mid|emit_kernel|insert_kernel|rank_verts|faces_kernel|filter_mi|highest|emit_mi|hash_insert|rank_reps|write_verts|general_mi|classify_mi|count_scan|write_faces_mi" --launch-skip $skip --launch-count ${NCU_COUNT:-$cnt} \
    -f -o gpurun_out/ncu_r2_$tag python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
(echo "# ncu --set full --clock-control none, config $cfg, one launch of each pipeline kernel (scripts/prof_r2.sh $tag $cfg)"; echo; python scripts/ncu_summary.py gpurun_out/ncu_r2_$tag.ncu-rep) > gpurun_out/ncu_r2_${tag}_summary.md
python scripts/ncu_traffic.py gpurun_out/ncu_r2_$tag.ncu-rep $cfg "profiles/ncu_r2_${tag}_summary.md (ncu --set full, scripts/prof_r2.sh $tag $cfg)" > gpurun_out/ncu_traffic_$tag.json
[ -z "$KEEP_REP" ] && rm -f gpurun_out/ncu_r2_$tag.ncu-rep
cut -c1-1800 gpurun_out/bench_r2_$tag.json
