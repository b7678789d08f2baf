import collections, json, sys
b = json.load(open(sys.argv[1]))
tot = sum(k['ms_per_step'] for k in b)
print("total crf ms/step %.3f" % tot)
agg = collections.defaultdict(float)
for k in b:
    n = k['kernel']
    key = n.split('_M')[0] if n.startswith('gemm') else '_'.join(n.split('_')[:2])
    agg[key] += k['ms_per_step']
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"{k:28s} {v:7.3f} ms")
print()
for k in b[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{k['kernel']:44s} n={k['launches']:3d} {k['ms_per_step']:7.3f} ms/step {k['tflops']:7.1f} TF/s {k['gbs']:7.0f} GB/s")
