#!/bin/bash
# Round-2 (final build) ncu captures of one eager MFT train step (tools/ncu_step.py, grouped stacks): launch list with DRAM bytes, then
# --set full of the kernels this round's last commits changed: tcgen05 attention with keep bits (+ the bit-draw kernel), the second-cut
# MFN recurrences, the grouped QKV input gradient.  Usage (on the GPU box): bash tools/ncu_capture_r02c.sh r02_c
TAG=${1:-r02_c}
OUT=gpurun_out
mkdir -p $OUT
python tools/ncu_step.py > $OUT/${TAG}_step_plain.log 2>&1 || { echo "plain step failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --page raw --log-file $OUT/${TAG}_launches.csv python tools/ncu_step.py > $OUT/${TAG}_ncu_list.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_summary.txt 2>&1
cap() {  # name regex skip count
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name "regex:$2" -s $3 -c $4 \
      -f -o $OUT/${TAG}_$1 python tools/ncu_step.py > $OUT/${TAG}_ncu_$1.log 2>&1
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
  python tools/ncu_keys.py $OUT/${TAG}_$1_raw.csv > $OUT/${TAG}_full_$1.txt 2>&1
  rm -f $OUT/${TAG}_$1_raw.csv
}
cap attn 'attn_tc_fwd_kernel|attn_tc_bwd2_kernel|attn_tc_dropbits' 6 3
cap rec 'lstm_fwd_mma2|lstm_bwd_mma2|mem_fwd_mma2|mem_bwd_mma2' 0 4
cap dgrad 'gemm_tc_kernel<256' 0 2
rm -f $OUT/${TAG}_*.ncu-rep
ls -la $OUT | tail -20
