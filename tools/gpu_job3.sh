#!/bin/bash
mkdir -p gpurun_out
T=40 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_predict.csv python tools/bench_predict.py > gpurun_out/bench_predict_ncu.log 2>&1
tail -2 gpurun_out/bench_predict_ncu.log
