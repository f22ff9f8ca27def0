#!/bin/bash
# ncu --set full capture of the front-end's tcgen05 GEMMs (conv over overlapping rows, conv wgrad) for the linguistic modality.
set -e
mkdir -p gpurun_out
cat > /tmp/front_one.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import multimodal_transformer_b200 as mtb
from multimodal_transformer_b200 import functional as K
sys.path.insert(0, 'tools')
from bench_frontend import params
mtb.set_compute_dtype('bf16')
dev = torch.device('cuda:0')
x = torch.randn(256, 128, 33, 300, device=dev)
ps = params(300, 300, 2, dev)
w = torch.randn(256, 128, 300, device=dev)
for _ in range(2):
    y = K.window_cnn(x, *ps, p_drop=0.3)
    y.backward(w)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 5 -c 5 -o gpurun_out/front_gemm python /tmp/front_one.py > gpurun_out/front_ncu3.log 2>&1
ncu -i gpurun_out/front_gemm.ncu-rep --page raw --csv > gpurun_out/front_gemm_raw.csv 2>/dev/null
rm -f gpurun_out/front_gemm.ncu-rep
