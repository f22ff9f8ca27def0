#!/bin/bash
# ncu --set full capture of the launches of one kernel (regex) inside one eager MFT train step (tools/ncu_step.py), summarised into
# gpurun_out/<tag>_full.txt (key metrics) and gpurun_out/<tag>_mix.txt (opcode mix + top stall sites).
# Usage (GPU box, after `python tools/ncu_step.py` exited 0): bash tools/ncu_cap_kernel.sh <kernel-regex> <tag> [skip=2] [count=1]
# NCU_BASE=demangled matches the regex against the full demangled name (template arguments: one instantiation of gemm_tc_kernel).
# mt_tune presets travel through MT_B200_TUNE, e.g. MT_B200_TUNE=13=1 bash tools/ncu_cap_kernel.sh ln_bwd_kernel r02l_lnbwd_bf16g
RE=$1; TAG=$2; SKIP=${3:-2}; CNT=${4:-1}; OUT=gpurun_out
ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base ${NCU_BASE:-function} --kernel-name "regex:$RE" -s $SKIP -c $CNT \
    -f -o $OUT/$TAG python tools/ncu_step.py > $OUT/${TAG}_ncu.log 2>&1
ncu -i $OUT/$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
python tools/ncu_keys.py $OUT/${TAG}_raw.csv > $OUT/${TAG}_full.txt 2>&1
ncu -i $OUT/$TAG.ncu-rep --page source --csv > $OUT/${TAG}_source.csv 2>/dev/null
python tools/ncu_sass_mix.py $OUT/${TAG}_source.csv 24 > $OUT/${TAG}_mix.txt 2>&1
rm -f $OUT/$TAG.ncu-rep $OUT/${TAG}_raw.csv $OUT/${TAG}_source.csv
