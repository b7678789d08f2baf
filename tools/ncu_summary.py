"""Summarise `ncu --set full` reports (.ncu-rep) into one JSON: duration, DRAM traffic, tensor-pipe and issue activity,
L2 / DRAM throughput percentages, occupancy limiters and the top warp-stall reasons.

    python tools/ncu_summary.py out.json label=path.ncu-rep[:gflop] [label=path.ncu-rep[:gflop] ...]

Tensor-pipe utilisation.  `sm__pipe_tensor_cycles_active_realtime ... pct_of_peak_sustained_elapsed` is normalised to FOUR
tensor sub-pipes per SM, but a tcgen05.mma is issued by one thread and executes as one SM-wide operation that ncu books
on a single sub-partition: the raw figure is therefore 1/4 of the utilisation that reconciles with flops / time
(round-1 C = 1024 GEMM: raw 4.4 % x 4 = 17.6 % against 482 TFLOP/s / 2303 TFLOP/s nominal at the measured clock = 21 %).
`tensor_pipe_util_pct` below is 4 x the raw metric; when the algorithmic GFLOP of the launch is given after the path,
`tensor_util_from_flops_pct` = flops / duration / (148 SMs x 8192 flop/clk x measured SM clock) is printed beside it.
The two agree for the round-1 C = 1024 GEMM and for the fused-MLP inference capture (16.6 % vs 17.1 %) but NOT for the
fused-MLP training capture (23.1 % vs 12.9 %: the same MMA work in a longer kernel cannot be busier), and the report's
`sm__pipe_tensor_subpipe_hmma_cycles_active_realtime` carries the same value in unrelated captures: on this driver
(580.159) ncu's tensor counters for tcgen05 are not trustworthy on their own -- the flops-derived figure is the one to
quote, the counter is printed beside it.
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "duration_us": "gpu__time_duration.sum",
    "dram_read_MB": "dram__bytes_read.sum",
    "dram_write_MB": "dram__bytes_write.sum",
    "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_active_pct_alt": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers_per_thread": "launch__registers_per_thread",
    "dyn_smem_per_block_KB": "launch__shared_mem_per_block_dynamic",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "limit_blocks_smem": "launch__occupancy_limit_shared_mem",
    "limit_blocks_regs": "launch__occupancy_limit_registers",
    "sm_clock_mhz": "sm__cycles_elapsed.avg.per_second",
    "sm_cycles_elapsed": "sm__cycles_elapsed.avg",
    "tmem_pipe_pct": "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
}


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def summarise(path, gflop=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([ln for ln in out.splitlines() if ln.startswith('"')]))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    s = {"kernel": col.get("Kernel Name", ("", ""))[1][:120]}
    for k, name in KEYS.items():
        if name in col:
            u, v = col[name]
            f = to_float(v)
            if f is None:
                continue
            if u == "Kbyte" and k.endswith("_MB"):
                f /= 1e3
            if u == "Gbyte" and k.endswith("_MB"):
                f *= 1e3
            if u == "byte" and k.endswith("_MB"):
                f /= 1e6
            if u in ("ms", "msecond") and k == "duration_us":
                f *= 1e3
            if u == "Kbyte" and k.endswith("_KB"):
                pass
            if u == "byte" and k.endswith("_KB"):
                f /= 1e3
            s[k] = round(f, 3)
    if "tensor_pipe_active_pct" not in s and "tensor_pipe_active_pct_alt" in s:
        s["tensor_pipe_active_pct"] = s.pop("tensor_pipe_active_pct_alt")
    s.pop("tensor_pipe_active_pct_alt", None)
    if "tensor_pipe_active_pct" in s:
        s["tensor_pipe_raw_pct"] = s.pop("tensor_pipe_active_pct")
        s["tensor_pipe_util_pct"] = round(4.0 * s["tensor_pipe_raw_pct"], 2)
    if gflop and s.get("duration_us") and s.get("sm_clock_mhz"):
        ghz = s["sm_clock_mhz"] if s["sm_clock_mhz"] < 100 else s["sm_clock_mhz"] / 1e3   # ncu prints GHz
        s["algorithmic_gflop"] = gflop
        s["achieved_tflops"] = round(gflop / s["duration_us"] * 1e3, 1)
        s["tensor_util_from_flops_pct"] = round(100.0 * gflop * 1e9 / (s["duration_us"] * 1e-6) / (148 * 8192 * ghz * 1e9), 2)
    stalls = {}
    for h, (u, v) in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            f = to_float(v)
            if f:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(f, 3)
    s["top_stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    if "dram_read_MB" in s and "dram_write_MB" in s:
        s["dram_traffic_MB"] = round(s["dram_read_MB"] + s["dram_write_MB"], 3)
        if s.get("duration_us"):
            s["dram_GBps"] = round(s["dram_traffic_MB"] / s["duration_us"] * 1e3, 1)
    return s


def main():
    out_path = sys.argv[1]
    res = {}
    for a in sys.argv[2:]:
        label, path = a.split("=", 1)
        gflop = None
        if ":" in path and not path.rsplit(":", 1)[1].endswith("rep"):
            path, g = path.rsplit(":", 1)
            gflop = float(g)
        res[label] = summarise(path, gflop)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
