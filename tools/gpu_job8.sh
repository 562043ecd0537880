#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_predict.py > gpurun_out/bench_predict.json 2> gpurun_out/bench_predict.err; cat gpurun_out/bench_predict.json; tail -3 gpurun_out/bench_predict.err
