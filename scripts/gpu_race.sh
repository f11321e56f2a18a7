#!/bin/bash
mkdir -p gpurun_out
timeout 70 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/racecheck.log
