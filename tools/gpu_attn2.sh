#!/bin/bash
# A/B of the attention-RNN kernel generations + phase trace of the second-generation kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 120 python tools/check_attn2.py small $CHECK_FLAGS > gpurun_out/check_attn2_small.log 2>&1; echo "rc=$?" >> gpurun_out/check_attn2_small.log
grep -v "^    " gpurun_out/check_attn2_small.log | tail -12
if grep -q "rc=0" gpurun_out/check_attn2_small.log; then
timeout 300 python tools/check_attn2.py $CHECK_FLAGS > gpurun_out/check_attn2.log 2>&1; echo "rc=$?" >> gpurun_out/check_attn2.log
grep -v "^    " gpurun_out/check_attn2.log | tail -30
SATK_LIB_PATH=$PWD/self-attention-tacotron_b200/libsatk_pt.so timeout 300 python tools/check_attn2.py big phases $CHECK_FLAGS > gpurun_out/check_attn2_pt.log 2>&1
tail -4 gpurun_out/check_attn2_pt.log
fi
