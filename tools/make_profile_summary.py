"""Builds the committed profile summaries under profiles/ from gpurun_out/<tag>_* dumps.
Usage: python tools/make_profile_summary.py <tag> <round>   (e.g. r01e r01)"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
tag, rnd = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list: per-kernel time shares of one bench step
rows = list(csv.reader(l for l in open(os.path.join(G, f"{tag}_launches.csv")) if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "")
    if "at::" in short:
        short = "torch elementwise/copy/reduce kernels (cotangent + gradient glue)"
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v_ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    a = agg.setdefault(short, [0, 0.0, r[ix["Block Size"]], r[ix["Grid Size"]]])
    a[0] += 1; a[1] += v_ms
tot = sum(a[1] for a in agg.values())
lines = [f"# {rnd}: ncu launch list of `python bench.py --steps 1 --warmup 1 --horizon 40 --e2e-steps 1 --no-cpu-baseline`",
         "", "(`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES)", "",
         "| kernel | launches | total ms | share | block | grid |", "|---|---:|---:|---:|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100*a[1]/tot:.1f}% | {a[2]} | {a[3]} |")
open(os.path.join(P, f"{rnd}_launches.md"), "w").write("\n".join(lines) + "\n")

# ---- ncu --set full raw metrics of the two kernels
rows = list(csv.reader(open(os.path.join(G, f"{tag}_raw.csv"))))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
out = {}
md = [f"# {rnd}: `ncu --set full --clock-control none` of the two kernels", "",
      "Command: `python tools/perf_probe.py --B 4096 --T 20 --lanes 8 --reps 1 --grad-only` (bench scene and inputs, 20 steps per launch).", ""]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
    d = {}
    md += [f"## {name}", "", "| metric | value | unit |", "|---|---:|---|"]
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            d[k] = r[i]
            md.append(f"| {k} | {r[i]} | {units[i]} |")
    out[name] = d
    md.append("")
json.dump(out, open(os.path.join(P, f"{rnd}_ncu_metrics.json"), "w"), indent=1)

# ---- stall reasons and per-function attribution
for kern in ("fwd", "bwd"):
    src = os.path.join(G, f"{tag}_src_{kern}.csv")
    if not os.path.exists(src):
        continue
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), src,
                          os.path.join(ROOT, "tactilesimulation_b200", "libtactilesim_b200.so"), f"{kern}_kernelILi8E"],
                         capture_output=True, text=True).stdout
    md += [f"## {kern}_kernel<8>: warp-state samples per source function (ncu source page joined with nvdisasm line info)", "", "```", txt.rstrip(), "```", ""]
open(os.path.join(P, f"{rnd}_ncu_summary.md"), "w").write("\n".join(md) + "\n")

def fnum(x):
    return float(x.replace(",", ""))
f = out.get("fwd_kernel<8>") or next(iter(out.values()))
tr = {"fwd_kernel_dram_bytes_per_launch": None}
try:
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    rd = fnum(f["dram__bytes_read.sum"]) * mult[units[hdr.index("dram__bytes_read.sum")]]
    wr = fnum(f["dram__bytes_write.sum"]) * mult[units[hdr.index("dram__bytes_write.sum")]]
    tr = {"capture": "B=4096, T=20 per launch", "fwd_kernel_dram_bytes_captured": rd + wr,
          "fwd_kernel_dram_bytes_per_env_step": (rd + wr) / (4096 * 20),
          "fwd_kernel_dram_bytes_per_launch": (rd + wr) / (4096 * 20) * 4096 * 200,
          "note": "per_launch = captured bytes per env-step x the 4096 x 200 env-steps of one bench launch"}
except Exception as e:
    tr["error"] = str(e)
json.dump(tr, open(os.path.join(P, f"{rnd}_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{rnd}_launches.md")).read())
print(json.dumps(tr))
