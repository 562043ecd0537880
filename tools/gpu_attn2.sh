#!/bin/bash
# A/B of the attention-RNN kernel generations + phase trace of the second-generation kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python tools/check_attn2.py $CHECK_FLAGS > gpurun_out/check_attn2.log 2>&1; echo "rc=$?" >> gpurun_out/check_attn2.log
tail -30 gpurun_out/check_attn2.log
SATK_LIB_PATH=$PWD/self-attention-tacotron_b200/libsatk_pt.so timeout 300 python tools/check_attn2.py big phases $CHECK_FLAGS > gpurun_out/check_attn2_pt.log 2>&1
tail -4 gpurun_out/check_attn2_pt.log
