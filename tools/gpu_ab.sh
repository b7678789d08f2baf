#!/bin/bash
# A/B timing of the attention generations and the GEMM shapes + parity tests (one gpurun call).
TAG=${1:-ab}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -n 5 $O/tests_$TAG.log
for st in 0 1 2 3; do
  echo "== attn stage $st (async)"; timeout 120 python tools/run_kernel.py attn --stage $st --iters 20 2>&1 | tail -n 2
  echo "== attn stage $st (pipe)";  CRF_ATTN_IMPL=pipe timeout 120 python tools/run_kernel.py attn --stage $st --iters 20 2>&1 | tail -n 2
done
for st in 0 2 3; do echo "== gemms stage $st"; timeout 120 python tools/run_kernel.py gemms --stage $st --iters 20 2>&1 | tail -n 8; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown $O/breakdown_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench rc=$?"; head -c 400 $O/bench_$TAG.json; echo
