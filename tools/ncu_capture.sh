#!/bin/bash
# Round-tagged ncu captures of one eager MFT train step (tools/ncu_step.py): launch list + --set full of the top kernels.
# Usage (on the GPU box): bash tools/ncu_capture.sh r01_c
TAG=${1:-r01_x}
OUT=gpurun_out
mkdir -p $OUT
python tools/ncu_step.py > $OUT/${TAG}_step_plain.log 2>&1 || { echo "plain step failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --page raw --log-file $OUT/${TAG}_launches.csv python tools/ncu_step.py > $OUT/${TAG}_ncu_list.log 2>&1
cap() {  # name regex skip count
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name "regex:$2" -s $3 -c $4 \
      -f -o $OUT/${TAG}_$1 python tools/ncu_step.py > $OUT/${TAG}_ncu_$1.log 2>&1
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
}
cap attn_bwd attn128_bwd 2 1
cap attn_fwd attn128_fwd 2 1
cap gemm 'gemm_tc_kernel' 7 8
cap gemm_bwd 'gemm_tc_kernel' 130 12
cap rowwise 'ln_bwd|ln_fwd' 40 4
cap mfn 'lstm_fwd_mma|lstm_bwd_mma|mem_fwd_mma|mem_bwd_mma' 0 4
ls -la $OUT
