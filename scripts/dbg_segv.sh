#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1
echo "full rc=$?"; tail -5 gpurun_out/pytest_final.log
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/memcheck.log
timeout 200 compute-sanitizer --tool initcheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/initcheck.log 2>&1; echo "initcheck rc=$?"; tail -8 gpurun_out/initcheck.log
