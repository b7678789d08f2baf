#!/bin/bash
# One gpurun call producing the evidence profiles/ cites: GPU parity tests, the bench line, the ncu launch list of the
# bench's timed region, and `ncu --set full` captures of the top kernels (each ncu run follows a plain run, exit 0).
#   gpurun --timeout 1500 -- 'bash tools/ncu_evidence.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/tests_$TAG.log 2>&1; echo "tests rc=$?" | tee -a $O/tests_$TAG.log
python bench.py --steps 10 --warmup 3 --breakdown $O/breakdown_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-range > $O/plain_bench_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-range \
    > $O/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, run_kernel args...
  local name=$1 rx=$2; shift 2
  python tools/run_kernel.py "$@" --iters 1 > $O/plain_${name}_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o $O/prof_${name}_$TAG \
      python tools/run_kernel.py "$@" --iters 1 > $O/ncu_${name}_$TAG.log 2>&1
  echo "$name rc=$?"
}
cap gemm_fc1 gemm_persistent gemms --stage 0 --only fc1+gelu
cap gemm_dfc2 gemm_persistent gemms --stage 0 --only "d_fc2*gelu'"
cap gemm_fc1_s3 gemm_persistent gemms --stage 3 --only fc1+gelu
cap attn_fwd attn_fwd attn --stage 0
cap attn_bwd attn_bwd attn --stage 0
tail -n 2 $O/tests_$TAG.log; head -c 600 $O/bench_$TAG.json; echo
