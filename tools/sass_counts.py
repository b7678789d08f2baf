"""Per-kernel SASS opcode counts of libcrf_sm100.so: the evidence that the hot kernels are Blackwell-native
(`UTCHMMA` = tcgen05.mma, `UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA load / store / reduce, `LDTM` / `STTM` = tcgen05.ld / st,
`UTCBAR` = tcgen05.commit, `LDGSTS` = cp.async, `SYNCS` = mbarrier ops; `HMMA` would be the legacy mma.sync path).

    python tools/sass_counts.py > profiles/r02_sass_counts.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "monocular_depth_estimation_b200", "csrc", "libcrf_sm100.so")
OPS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "LDGSTS", "SYNCS", "FFMA2", "FMUL2", "FADD2",
       "MUFU", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1).split(".")[0]
            if op in OPS:
                cur[op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode counts per kernel, libcrf_sm100.so ({', '.join(arch)}; cuobjdump -sass)\n")
    print("| kernel | " + " | ".join(OPS) + " |")
    print("|---|" + "---:|" * len(OPS))
    tot = collections.Counter()
    for (name, c), dm in zip(per.items(), demangle):
        if not any(c[o] for o in OPS[:8]):   # list the kernels that use the Blackwell / async machinery
            continue
        short = dm.replace("crf::(anonymous namespace)::", "").replace("crf::", "")
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\(CUtensorMap_st.*|\((?:const |unsigned |float|int|long|void|crf|__nv).*", "", short)
        print(f"| `{short[:70]}` | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
        tot.update(c)
    print("| **total** | " + " | ".join(str(tot[o]) for o in OPS) + " |")


if __name__ == "__main__":
    sys.exit(main())
