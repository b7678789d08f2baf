O=gpurun_out; T=r02final
cap() { local name=$1 rx=$2; shift 2
  timeout 60 python tools/run_kernel.py "$@" --iters 1 > $O/plain_${name}_$T.log 2>&1 &&
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o $O/prof_${name}_$T python tools/run_kernel.py "$@" --iters 1 > $O/ncu_${name}_$T.log 2>&1
  echo "$name rc=$?"; }
cap lnbwd128 dgrad_lnbwd lnbwd --stage 0
cap pair_fc1 gemm_pair gemms --stage 3 --only fc1+gelu
timeout 240 python tools/ncu_block_traffic.py run > $O/blk_traffic_$T.log 2>&1; echo "block traffic rc=$?"
