#!/bin/bash
# Runs the stage-level GPU tests one group per process (a device-side trap poisons the CUDA context of its process).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in test_gemm_fprop test_gemm_dgrad test_gemm_wgrad test_gemm_epilogues test_ln_fwd test_ln_bwd \
           test_colsum_cast_convert test_window_index_maps test_attn_fwd test_attn_bwd; do
  echo "=== $grp ===" 
  timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "$grp" -p no:cacheprovider 2>&1 | tail -n 60
done
