#!/bin/bash
# Round 2 evidence: ncu --set full of the recurrent kernels at config 2 (B=32, Tt=148, Tm=800) + the launch list of one step.
mkdir -p gpurun_out
cap() {  # <kernel regex> <fwd|bwd> <tag> <skip>
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $4 -c 1 -f -o gpurun_out/$3 python tools/run_attn_once.py 32 $2 > gpurun_out/ncu_$3.log 2>&1; tail -1 gpurun_out/ncu_$3.log
}
cap attn_rnn2_fwd fwd r02_attn_rnn2_fwd 1
cap attn_rnn2_bwd bwd r02_attn_rnn2_bwd 1
cap attn_energy_grad bwd r02_attn_energy_grad 1
cap lstm5_bwd bwd r02_lstm5_bwd 2
cap lstm5_fwd fwd r02_lstm5_fwd 2
SATK_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches.csv | head -12
