#!/usr/bin/env bash
# First GPU call of the next round: runs everything that was written after the round-1 GPU budget was spent
#  and leaves the results under gpurun_out/.  One GPU, ~10 minutes.
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call_next_round.sh'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."

# 0. seconds, no Python: correctness of the opt-in kernels through the C ABI, then CUDA-event timings at config-2 sizes
#    (LayerNorm rows 1 / 2 / 4, Adam over 45 M parameters, attention at head_dim 32 / 64 / 128)
make -C monocular_depth_estimation_b200/csrc hwcheck > /dev/null 2>&1
timeout 60 ./tools/hwcheck > gpurun_out/n_hwcheck.log 2>&1; echo "hwcheck rc=$? ($(grep -c PASS gpurun_out/n_hwcheck.log) PASS, $(grep -c FAIL gpurun_out/n_hwcheck.log) FAIL)" | tee gpurun_out/n_summary.txt
timeout 120 ./tools/hwcheck --bench > gpurun_out/n_hwcheck_bench.log 2>&1; grep bench gpurun_out/n_hwcheck_bench.log | tee -a gpurun_out/n_summary.txt

# 1. the verified suite must still be green on the rebuilt library (new translation units were added to the .so)
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_zz_gpu_unverified.py > gpurun_out/n_tests_verified.log 2>&1
echo "verified suite rc=$?" | tee -a gpurun_out/n_summary.txt

# 2. the unverified groups, each in its own subprocess (logs: gpurun_out/unverified_<group>.log)
timeout 1100 python -m pytest tests/test_zz_gpu_unverified.py -q -rxX > gpurun_out/n_tests_unverified.log 2>&1
echo "unverified groups rc=$? ($(grep -c XPASS gpurun_out/n_tests_unverified.log) xpass, $(grep -c XFAIL gpurun_out/n_tests_unverified.log) xfail)" | tee -a gpurun_out/n_summary.txt

# 3. A/B of the opt-in switches on the bench (three repetitions each: step time is bimodal)
for rep in 1 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_base_$rep.json 2> gpurun_out/n_bench_base_$rep.err
  CRF_LN_ROWS=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_lnrows4_$rep.json 2> gpurun_out/n_bench_lnrows4_$rep.err
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lib-adam > gpurun_out/n_bench_libadam_$rep.json 2> gpurun_out/n_bench_libadam_$rep.err
done
python - <<'PY' | tee -a gpurun_out/n_summary.txt
import glob, json
for tag in ("base", "lnrows4", "libadam"):
    ms = []
    for f in sorted(glob.glob(f"gpurun_out/n_bench_{tag}_*.json")):
        try:
            ms.append(round(json.loads(open(f).read().strip().splitlines()[-1])["ms_per_step"], 3))
        except Exception as e:
            ms.append(f"failed: {type(e).__name__}")
    print(tag, "ms/step:", ms)
PY

# 4. reported baselines and the configs that have no number yet
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gpu-eager-baseline > gpurun_out/n_bench_gpu_eager.json 2> gpurun_out/n_bench_gpu_eager.err
timeout 300 python tools/bench_config5.py > gpurun_out/n_config5.json 2> gpurun_out/n_config5.err
timeout 400 python tools/sweep_config3.py > gpurun_out/n_config3_sweep_wide.md 2> gpurun_out/n_config3_sweep_wide.err
echo "done" | tee -a gpurun_out/n_summary.txt
