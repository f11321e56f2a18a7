#!/bin/bash
mkdir -p gpurun_out
timeout 80 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_final2.json
