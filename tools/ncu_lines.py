# attribute ncu SASS-level samples to CUDA source lines using nvdisasm -g line info
import csv, re, subprocess, sys, collections
rep_csv, cubin, func_key = sys.argv[1], sys.argv[2], sys.argv[3]
KERNEL_FILE=sys.argv[5] if len(sys.argv)>5 else "crf_mlp_fused.cu"
dis = subprocess.run(["nvdisasm","-gi","-c",cubin],capture_output=True,text=True).stdout.splitlines()
# locate function section
start=None
for i,l in enumerate(dis):
    if l.startswith(".text.") and func_key in l and l.rstrip().endswith(":"):
        start=i;break
lines=[]  # (offset, file,line)
cur=None; off_re=re.compile(r"/\*([0-9a-f]{4,})\*/")
for l in dis[start+1:]:
    if l.startswith("//---") and ".text." in l: break
    m=re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?',l)
    if m:
        if m.group(3) and KERNEL_FILE in m.group(3):
            cur=(m.group(3).split('/')[-1],int(m.group(4)))
        elif KERNEL_FILE in m.group(1):
            cur=(m.group(1).split('/')[-1],int(m.group(2)))
        elif m.group(3):
            pass  # nested inline: keep the enclosing kernel-file line
        else:
            cur=(m.group(1).split('/')[-1],int(m.group(2)))
        continue
    m=off_re.search(l)
    if m and '/*' in l and ';' in l:
        lines.append((int(m.group(1),16),cur))
offs={o:c for o,c in lines}
rows=list(csv.reader(open(rep_csv)))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
base=int(data[0][ix["Address"]],16)
agg=collections.Counter(); stall=collections.defaultdict(collections.Counter)
stallk=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot=0
for r in data:
    n=float(r[ix["# Samples"]] or 0); tot+=n
    o=int(r[ix["Address"]],16)-base
    key=offs.get(o,("?",0))
    agg[key]+=n
    for k in stallk:
        v=float(r[ix[k]] or 0)
        if v: stall[key][k]+=v
print("total",tot)
for key,n in agg.most_common(int(sys.argv[4]) if len(sys.argv)>4 else 30):
    st=", ".join(f"{k[6:]}={v:.0f}" for k,v in stall[key].most_common(3))
    print(f"{n:7.0f} {100*n/tot:5.1f}%  {key[0]}:{key[1]}   {st}")

# --- instructions executed per source line (optional 6th arg "inst")
if len(sys.argv) > 6 and sys.argv[6] == "inst":
    inst = collections.Counter()
    for r in data:
        o = int(r[ix["Address"]], 16) - base
        inst[offs.get(o, ("?", 0))] += float(r[ix["Instructions Executed"]] or 0)
    tot_i = sum(inst.values())
    print("warp instructions executed:", tot_i)
    for key, n in inst.most_common(25):
        print(f"{n:12.0f} {100*n/tot_i:5.1f}%  {key[0]}:{key[1]}")
