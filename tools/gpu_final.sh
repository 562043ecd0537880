#!/bin/bash
# Final evidence of a round on ONE GPU: smoke, default bench line (with the CPU baseline), reference arm, the other BASELINE configs.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
for c in vctk location_sensitive transition_agent predict; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_$c.json').readlines()[-1]); print('$c', d['ms_per_step'], d['value'], d['unit'])"
done
timeout 200 python tools/timeline.py graph 2>&1 | tail -1; gzip -f gpurun_out/trace_graph.json
SATK_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches.csv 2>/dev/null | head -8
