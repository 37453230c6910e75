"""Development tool: join an `ncu --page source --csv` dump of one kernel with nvdisasm line info and
aggregate samples / executed instructions / stall reasons per source function.
Usage: ncu_by_function.py <src.csv> <lib.so> <kernel-substring>"""
import collections, csv, os, re, subprocess, sys, tempfile
csvp, so, pat = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
sass = []          # the library holds one cubin per kernel variant: keep the one that has the kernel
for cubin in sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    if any(l.startswith(".text.") and pat in l for l in txt.split("\n")):
        sass = txt.split("\n")
        break

def func_ranges(path):
    out, cur = [], None
    for l in open(path).read().split("\n"):
        m = re.match(r"^(?:template.*>\s*)?(?:TS_NOINLINE\s+)?(?:HDN|HD|static|__global__|__device__)[\w\s\*&:<>,]*?\b(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            cur = m.group(1)
        out.append(cur)
    return out
srcs, off2fn, insec, curfn = {}, {}, False, "?"
for l in sass:
    if l.startswith(".text."):
        insec = pat in l
        continue
    if l.startswith("//-----"):
        insec = False
    if not insec:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m:
        path, ln = m.group(1), int(m.group(2))
        if path not in srcs:
            srcs[path] = func_ranges(path) if os.path.exists(path) else None
        r = srcs[path]
        curfn = (r[ln - 1] if r and ln - 1 < len(r) else None) or os.path.basename(path)
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
    if m:
        off2fn[int(m.group(1), 16)] = curfn
rows = list(csv.reader(open(csvp)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix["# Samples"]].isdigit()]
base = min(int(r[ix["Address"]], 16) for r in data)
agg = collections.defaultdict(lambda: collections.Counter())
for r in data:
    fn = off2fn.get(int(r[ix["Address"]], 16) - base, "?")
    a = agg[fn]
    a["samples"] += int(r[ix["# Samples"]]); a["exec"] += int(r[ix["Instructions Executed"]]); a["static"] += 1
    for s in stalls:
        a[s] += int(r[ix[s]])
ts = sum(a["samples"] for a in agg.values()); te = sum(a["exec"] for a in agg.values())
print(f"samples {ts} executed warp-instr {te} static {len(data)}")
tot = collections.Counter()
for a in agg.values():
    tot.update(a)
print("stall mix:", ", ".join(f"{s[6:]} {100*tot[s]/ts:.1f}%" for s in sorted(stalls, key=lambda s: -tot[s])[:7]))
print(f"{'function':28s} {'static':>7s} {'exec%':>6s} {'samp%':>6s}  top stalls")
for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:32]:
    top = sorted(stalls, key=lambda s: -a[s])[:3]
    print(f"{fn:28s} {a['static']:7d} {100*a['exec']/te:6.1f} {100*a['samples']/ts:6.1f}  " + ", ".join(f"{s[6:]} {100*a[s]/max(a['samples'],1):.0f}%" for s in top))
