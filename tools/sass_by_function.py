"""Development tool: static SASS instruction count of one kernel, attributed to the source function that
the (innermost) line-info entry points at.  Usage: sass_by_function.py <lib.so> <kernel-substring>"""
import re, subprocess, sys, os, tempfile, collections
so, pat = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
sass = []          # the library holds one cubin per kernel variant: keep the one that has the kernel
for cubin in sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    if any(l.startswith(".text.") and pat in l for l in txt.split("\n")):
        sass = txt.split("\n")
        break
# function line ranges from the sources
def func_ranges(path):
    out = []
    lines = open(path).read().split("\n")
    cur = None
    for i, l in enumerate(lines, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:TS_NOINLINE\s+)?(?:HDN|HD|static|__global__|__device__)[\w\s\*&:<>,]*?\b(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            cur = m.group(1)
        out.append(cur)
    return out
srcs = {}
insec = False
cnt = collections.Counter()
curfn = "?"
for l in sass:
    if l.startswith(".text."):
        insec = pat in l
        continue
    if l.startswith("//-----") :
        insec = False
    if not insec:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        path, ln = m.group(1), int(m.group(2))
        if path not in srcs:
            srcs[path] = func_ranges(path) if os.path.exists(path) else None
        r = srcs[path]
        curfn = (r[ln - 1] if r and ln - 1 < len(r) else None) or os.path.basename(path) + ":" + str(ln)
        continue
    if re.match(r"\s*/\*[0-9a-f]+\*/", l):
        cnt[curfn] += 1
tot = sum(cnt.values())
print("total", tot)
for k, v in cnt.most_common(40):
    print(f"{v:7d} {100*v/tot:5.1f}%  {k}")
