"""DRAM traffic of one CRFBlock forward + backward per decoder scale, from ncu launch lists with the two DRAM byte counters
(no replay of the full metric set: `--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`).

    # on a B200 (one GPU): writes gpurun_out/blk_traffic_s{0..3}.csv
    python tools/ncu_block_traffic.py run
    # anywhere: sums the library kernels of ONE block step per scale and writes profiles/ncu_traffic.json
    python tools/ncu_block_traffic.py summarise gpurun_out profiles/ncu_traffic.json
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = {0: (120, 160, 128), 1: (60, 80, 256), 2: (30, 40, 512), 3: (15, 20, 1024)}


def run():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for s in STAGES:
        cmd = [sys.executable, os.path.join(ROOT, "tools", "run_kernel.py"), "block", "--stage", str(s), "--iters", "1"]
        plain = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
        if plain.returncode != 0:
            print("plain run failed:", plain.stderr[-2000:])
            return 1
        out = os.path.join(ROOT, "gpurun_out", f"blk_traffic_s{s}.csv")
        subprocess.run(["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
                        "--clock-control", "none", "-c", "400", "--csv", "--log-file", out] + cmd, cwd=ROOT)
    return 0


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def summarise(src, dst):
    res = {}
    total = 0.0
    for s, (H, W, C) in STAGES.items():
        path = os.path.join(src, f"blk_traffic_s{s}.csv")
        rows = list(csv.DictReader(ln for ln in open(path) if ln.startswith('"')))
        per_launch = {}
        for r in rows:
            k = per_launch.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "bytes": 0.0, "us": 0.0})
            if r["Metric Name"].startswith("dram__bytes"):
                k["bytes"] += to_bytes(r["Metric Value"], r["Metric Unit"])
            elif r["Metric Name"].startswith("gpu__time"):
                k["us"] = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r["Metric Unit"], 1)
        lib = [k for _, k in sorted(per_launch.items()) if "crf::" in k["name"]]
        # run_kernel.py runs the block step twice (one warm-up, one timed): the second half is the steady state
        half = lib[len(lib) // 2:]
        b = sum(k["bytes"] for k in half)
        res[f"block_B8_{H}x{W}_C{C}"] = {"dram_bytes": b, "kernels": len(half), "sum_kernel_us_cold": sum(k["us"] for k in half),
                                          "by_kernel": [{"kernel": k["name"].split("(")[0][-60:], "dram_MB": round(k["bytes"] / 1e6, 2),
                                                         "us": round(k["us"], 1)} for k in half]}
        total += 2 * b
    res["crf_blocks_step"] = {"dram_bytes": total, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, sum over the "
                              "library kernels of one CRFBlock fwd+bwd per decoder scale (B = 8, shifted), x 2 blocks per "
                              "layer; tools/ncu_block_traffic.py"}
    with open(dst, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: (v["dram_bytes"] / 1e6) for k, v in res.items()}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        sys.exit(run())
    summarise(sys.argv[2], sys.argv[3])
