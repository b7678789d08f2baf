O=gpurun_out; T=r02final2
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/tests_$T.log 2>&1; echo "tests rc=$?" | tee -a $O/tests_$T.log
timeout 400 python bench.py --breakdown $O/breakdown_$T.json > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"
timeout 300 python tools/sweep_config3.py > $O/config3_$T.md 2> $O/config3_$T.err; echo "config3 rc=$?"
tail -n 2 $O/tests_$T.log
