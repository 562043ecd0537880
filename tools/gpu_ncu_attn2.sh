#!/bin/bash
# ncu --set full (with source counters) of the second-generation attention-RNN kernels at config 2 (B=32).
mkdir -p gpurun_out
MODE=${1:-fwd}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_rnn2 -s 1 -c 1 -f -o gpurun_out/attn_rnn2_$MODE python tools/run_attn_once.py 32 $MODE > gpurun_out/ncu_attn2_$MODE.log 2>&1; tail -3 gpurun_out/ncu_attn2_$MODE.log
