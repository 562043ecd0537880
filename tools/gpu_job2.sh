#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_predict_gpu.py -x -q > gpurun_out/pytest_predict.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_predict.log
tail -30 gpurun_out/pytest_predict.log
timeout 300 python tools/bench_predict.py > gpurun_out/bench_predict.json 2> gpurun_out/bench_predict.err; cat gpurun_out/bench_predict.json; tail -5 gpurun_out/bench_predict.err
