#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 800 python -m pytest tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -60 > gpurun_out/r2x_fuzz.log
tail -60 gpurun_out/r2x_fuzz.log
