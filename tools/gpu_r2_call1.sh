#!/usr/bin/env bash
# Round-2 first GPU call: the whole -m gpu suite (no -x: every failure with its message), then the reported baselines.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -q -m gpu -rfEs -p no:cacheprovider > gpurun_out/r2c1_tests.log 2>&1
echo "gpu suite rc=$?" | tee gpurun_out/r2c1_summary.txt
tail -5 gpurun_out/r2c1_tests.log | tee -a gpurun_out/r2c1_summary.txt
timeout 500 python bench.py --steps 20 --warmup 5 --gpu-eager-baseline > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench rc=$?" | tee -a gpurun_out/r2c1_summary.txt
timeout 300 python tools/bench_config5.py > gpurun_out/r2c1_config5.json 2> gpurun_out/r2c1_config5.err
echo "config5 rc=$?" | tee -a gpurun_out/r2c1_summary.txt
timeout 400 python tools/sweep_config3.py > gpurun_out/r2c1_config3_sweep.md 2> gpurun_out/r2c1_config3_sweep.err
echo "sweep rc=$?" | tee -a gpurun_out/r2c1_summary.txt
