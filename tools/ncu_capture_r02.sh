#!/bin/bash
# Round-2 ncu captures of one eager MFT train step (tools/ncu_step.py, grouped stacks): launch list with DRAM bytes, then --set full of
# the row-stream GEMMs, the grouped tcgen05 attention and the LayerNorm kernels.  Usage (on the GPU box): bash tools/ncu_capture_r02.sh r02_b
TAG=${1:-r02_x}
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
cap rs_fwd 'gemm_rs_kernel' 0 4
cap rs_bwd 'gemm_rs_kernel' 24 3
cap attn 'attn_tc_fwd_kernel|attn_tc_bwd_kernel|attn_tc_prep' 5 3
cap gemm_old 'gemm_tc_kernel' 20 5
cap ln 'ln_bwd|ln_fwd' 6 3
ls -la $OUT | tail -30
