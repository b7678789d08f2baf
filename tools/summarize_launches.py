"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share.
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event breakdown."""
import collections
import csv
import re
import sys


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = val / 1e3 if unit in ("ns", "nsecond") else (val * 1e3 if unit in ("ms", "msecond") else val)
        name = r["Kernel Name"]
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)[:90]
        rows.append((name, us, r["Grid Size"], r["Block Size"]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, us, *_ in rows:
        agg[n][0] += 1
        agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "crf::" in k)
    lines = [f"launches: {len(rows)}   total device time: {tot/1e3:.3f} ms   library (crf::) share: {100*ours/tot:.1f} %", "",
             "| kernel | launches | total us | share % |", "|---|---:|---:|---:|"]
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        lines.append(f"| `{k}` | {c} | {us:.1f} | {100*us/tot:.2f} |")
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
