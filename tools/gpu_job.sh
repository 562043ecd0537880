#!/bin/bash
# Round GPU job: GPU test suite, train-step bench, (optional) free-running decode bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_FLAGS:---no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline'].get('section_ms'))"
if [ -n "$PREDICT" ]; then timeout 300 python tools/bench_predict.py > gpurun_out/bench_predict.json 2> gpurun_out/bench_predict.err; cat gpurun_out/bench_predict.json; fi
