#!/bin/bash
OUT=gpurun_out
ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name "regex:attn_tc_bwd2_kernel" -s 2 -c 1 \
    -f -o $OUT/r02_d_attnbwd python tools/ncu_step.py > $OUT/r02_d_ncu_attnbwd.log 2>&1
ncu -i $OUT/r02_d_attnbwd.ncu-rep --page raw --csv > $OUT/r02_d_attnbwd_raw.csv 2>/dev/null
python tools/ncu_keys.py $OUT/r02_d_attnbwd_raw.csv > $OUT/r02_d_full_attnbwd.txt 2>&1
ncu -i $OUT/r02_d_attnbwd.ncu-rep --page source --csv > $OUT/r02_d_attnbwd_source.csv 2>/dev/null
rm -f $OUT/r02_d_attnbwd.ncu-rep $OUT/r02_d_attnbwd_raw.csv
ls -la $OUT/r02_d_*
