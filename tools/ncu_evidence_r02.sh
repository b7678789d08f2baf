#!/bin/bash
# Round-2 evidence in one gpurun call (each ncu run follows a plain run of the same command that exited 0):
#   gpurun --timeout 1700 -- 'bash tools/ncu_evidence_r02.sh'
O=gpurun_out
T=r02final
mkdir -p $O
python -m pytest tests -m gpu -q -p no:cacheprovider > $O/tests_$T.log 2>&1; echo "tests rc=$?" | tee -a $O/tests_$T.log
python bench.py --breakdown $O/breakdown_$T.json > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$T.json 2> $O/bench_ref_$T.err; echo "reference arm rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --profile-range > $O/plain_bench_$T.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --profile-range \
    > $O/ncu_bench_$T.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, run_kernel args...
  local name=$1 rx=$2; shift 2
  python tools/run_kernel.py "$@" --iters 1 > $O/plain_${name}_$T.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o $O/prof_${name}_$T \
      python tools/run_kernel.py "$@" --iters 1 > $O/ncu_${name}_$T.log 2>&1
  echo "$name rc=$?"
}
cap lnbwd128 dgrad_lnbwd lnbwd --stage 0
cap lnbwd256 dgrad_lnbwd lnbwd --stage 1
cap pair_fc1 gemm_pair gemms --stage 3 --only fc1+gelu
cap mlp128 mlp_fused mlp --stage 0
python tools/ncu_block_traffic.py run > $O/blk_traffic_$T.log 2>&1; echo "block traffic rc=$?"
tail -n 2 $O/tests_$T.log; head -c 400 $O/bench_$T.json; echo
