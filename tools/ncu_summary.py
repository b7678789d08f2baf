"""Summarise `ncu --set full` reports (.ncu-rep) into one JSON: duration, DRAM traffic, tensor-pipe and issue activity,
L2 / DRAM throughput percentages, occupancy limiters and the top warp-stall reasons.

    python tools/ncu_summary.py out.json label=path.ncu-rep [label=path.ncu-rep ...]
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
}


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def summarise(path):
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
        res[label] = summarise(path)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
