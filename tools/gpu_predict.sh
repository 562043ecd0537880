#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_predict_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/bench_predict.py 2>&1 | tail -2
