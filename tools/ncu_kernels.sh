#!/bin/bash
# ncu --set full captures of the attention and GEMM kernels at the scale-1/4 shape (each after a plain run, exit 0).
mkdir -p gpurun_out
python tools/run_kernel.py attn --stage 0 --iters 1 > gpurun_out/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_bwd -c 1 -f -o gpurun_out/prof_attn_bwd \
    python tools/run_kernel.py attn --stage 0 --iters 1 > gpurun_out/ncu_attn_bwd.log 2>&1
python tools/run_kernel.py attn --stage 0 --iters 1 > gpurun_out/plain_attn2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -f -o gpurun_out/prof_attn_fwd \
    python tools/run_kernel.py attn --stage 0 --iters 1 > gpurun_out/ncu_attn_fwd.log 2>&1
python tools/run_kernel.py gemm --stage 0 --iters 1 > gpurun_out/plain_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 1 -f -o gpurun_out/prof_gemm_fc1 \
    python tools/run_kernel.py gemm --stage 0 --iters 1 > gpurun_out/ncu_gemm.log 2>&1
cat gpurun_out/plain_attn.log gpurun_out/plain_gemm.log
