#!/bin/bash
# ncu captures for profiles/ (one gpurun call; each ncu run follows a plain run of the same command that exited 0).
set -o pipefail
mkdir -p gpurun_out
python tools/run_kernel.py attn --stage 0 --iters 1 > gpurun_out/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_bwd -c 1 -o gpurun_out/prof_attn_bwd \
    python tools/run_kernel.py attn --stage 0 --iters 1 > gpurun_out/ncu_attn_bwd.log 2>&1
python tools/run_kernel.py attn --stage 0 --iters 1 > gpurun_out/plain_attn2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/prof_attn_fwd \
    python tools/run_kernel.py attn --stage 0 --iters 1 > gpurun_out/ncu_attn_fwd.log 2>&1
python tools/run_kernel.py gemm --stage 0 --iters 1 > gpurun_out/plain_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 1 -o gpurun_out/prof_gemm_fc1 \
    python tools/run_kernel.py gemm --stage 0 --iters 1 > gpurun_out/ncu_gemm.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 6500 -c 1800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -n 20
tail -n 3 gpurun_out/ncu_attn_bwd.log gpurun_out/ncu_gemm.log gpurun_out/ncu_bench.log
