#!/bin/bash
# ncu --set full (with source counters) of the second-generation attention-RNN kernels at config 2 (B=32).
# usage: gpu_ncu_attn2.sh <kernel regex> <fwd|bwd> <tag>
mkdir -p gpurun_out
REGEX=${1:-attn_rnn2_fwd}; MODE=${2:-fwd}; TAG=${3:-$REGEX}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 1 -c 1 -f -o gpurun_out/$TAG python tools/run_attn_once.py 32 $MODE > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
