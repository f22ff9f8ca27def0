TAG=${1:-r01_n}; OUT=gpurun_out
cap() {
  ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name "regex:$2" -s $3 -c $4 -f -o $OUT/${TAG}_$1 python tools/ncu_step.py > $OUT/${TAG}_ncu_$1.log 2>&1
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_$1.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_$1_sass.csv 2>/dev/null
}
cap lstm_fwd lstm_fwd_mma 0 1
cap mem_fwd mem_fwd_mma 0 1
