#!/bin/bash
# Builds libmt_b200.so for sm_100a (B200) in-tree.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libmt_b200.so
SRCS="mt_api.cu mt_gemm.cu mt_gemm_simt.cu mt_gemm_tc.cu mt_gemm_rs.cu mt_elementwise.cu mt_attention.cu mt_attention_mma.cu mt_attention_t128.cu mt_attention_tc.cu mt_attention_flash.cu mt_encoder.cu mt_mfn.cu mt_mfn_mma.cu mt_lstm_head.cu mt_lstm_head_mma.cu mt_comm.cu mt_frontend.cu mt_metrics.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -O2"
mkdir -p build
pids=()
for f in $SRCS; do
  [ -f "$f" ] || continue
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$o")" ] || [ ../../include/mt_b200.h -nt "$o" ]; then
    nvcc $FLAGS "$@" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for f in $SRCS; do [ -f "$f" ] && OBJS="$OBJS build/${f%.cu}.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJS -ldl
echo "built $OUT"
