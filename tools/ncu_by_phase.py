"""Development tool: like ncu_by_function.py but attributes every SASS instruction to the OUTERMOST
'phase' function of its inline chain (nvdisasm -gi), so that vector helpers and dual-number operators are
charged to the phase that called them.  Usage: ncu_by_phase.py <src.csv> <lib.so> <kernel-substring>"""
import collections, csv, os, re, subprocess, sys, tempfile
csvp, so, pat = sys.argv[1], sys.argv[2], sys.argv[3]
PHASES = ["tactile_values", "tactile_vjp", "contact_sets", "mass_column", "lu_solve", "kinematics", "joint_dynamics",
          "ground_contacts", "gp_contacts", "inward", "eval_columns", "eval_g", "readout_from_work", "step_eval", "step_post",
          "step_backward", "step_begin", "env_forward", "env_backward", "stage_scene"]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.split("\n")

def func_ranges(path):
    out, cur = [], None
    for l in open(path).read().split("\n"):
        m = re.match(r"^(?:template.*>\s*)?(?:TS_NOINLINE\s+)?(?:HDN|HD|static|__global__|__device__)[\w\s\*&:<>,]*?\b(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            cur = m.group(1)
        out.append(cur)
    return out
srcs = {}
def fn_of(path, ln):
    if path not in srcs:
        srcs[path] = func_ranges(path) if os.path.exists(path) else None
    r = srcs[path]
    return (r[ln - 1] if r and 0 < ln <= len(r) else None) or os.path.basename(path)
off2ph, insec, chain = {}, False, []
fresh = True
for l in sass:
    if l.startswith(".text."):
        insec = pat in l
        continue
    if l.startswith("//-----"):
        insec = False
    if not insec:
        continue
    m = re.match(r'\s*//## File "([^"]*)", line (\d+)(?: inlined at "([^"]*)", line (\d+))?', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append(fn_of(m.group(1), int(m.group(2))))
        if m.group(3):
            chain.append(fn_of(m.group(3), int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
    if m:
        ph = None
        for f in chain:            # chain is innermost -> outermost; keep the outermost phase below the drivers
            if f in PHASES and f not in ("eval_columns", "eval_g", "step_eval", "step_post", "env_forward", "env_backward", "readout_from_work", "step_backward"):
                ph = f
        if ph is None:
            for f in chain:
                if f in PHASES:
                    ph = f
                    break
        off2ph[int(m.group(1), 16)] = ph or (chain[0] if chain else "?")
        fresh = True
rows = list(csv.reader(open(csvp)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix["# Samples"]].isdigit()]
seen, uniq = set(), []
for r in data:
    a = int(r[ix["Address"]], 16)
    if a not in seen:
        seen.add(a); uniq.append(r)
base = min(seen)
agg = collections.defaultdict(collections.Counter)
for r in uniq:
    a = agg[off2ph.get(int(r[ix["Address"]], 16) - base, "?")]
    a["samples"] += int(r[ix["# Samples"]]); a["exec"] += int(r[ix["Instructions Executed"]]); a["static"] += 1
    a["thr"] += int(r[ix["Thread Instructions Executed"]])
    for s in stalls:
        a[s] += int(r[ix[s]])
ts = sum(a["samples"] for a in agg.values()); te = sum(a["exec"] for a in agg.values())
print(f"samples {ts}  executed warp-instr {te}  static {len(uniq)}")
print(f"{'phase':22s} {'static':>7s} {'exec%':>6s} {'samp%':>6s} {'thr/inst':>8s}  top stalls")
for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:24]:
    top = sorted(stalls, key=lambda s: -a[s])[:3]
    print(f"{fn:22s} {a['static']:7d} {100*a['exec']/te:6.1f} {100*a['samples']/ts:6.1f} {a['thr']/max(a['exec'],1):8.1f}  " + ", ".join(f"{s[6:]} {100*a[s]/max(a['samples'],1):.0f}%" for s in top))
