TAG=r01_k; OUT=gpurun_out
cap() {
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name "regex:$2" -s $3 -c $4 -f -o $OUT/${TAG}_$1 python tools/ncu_step.py > $OUT/${TAG}_ncu_$1.log 2>&1
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_$1.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_$1_sass.csv 2>/dev/null
}
cap attn_bwd attn128_bwd 2 1
cap attn_fwd attn128_fwd 2 1
