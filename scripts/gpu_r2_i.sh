#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_tail.py -q -m gpu --timeout 600 2>&1 | tail -n 3
bash scripts/gpu_sanitize.sh
