#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_check.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_check.log
