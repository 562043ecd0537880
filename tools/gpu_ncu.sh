#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 4 -o gpurun_out/gemm_tc_full python tools/run_gemm_once.py > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_rnn -c 2 -o gpurun_out/attn_rnn_full_final python tools/run_attn_once.py 32 bwd > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
