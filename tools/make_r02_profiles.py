"""Builds the committed round-2 profile summaries under profiles/ from the gpurun_out/r02_* dumps of tools/gpu_ncu_r02.sh.
Usage: python tools/make_r02_profiles.py"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from tactilesimulation_b200.build import source_sha  # noqa: E402

G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
B, T = 4096, 200


def rows_of(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    return rows[1:], ix


def short(name):
    s = name.split("(")[0].replace("void ", "")
    return "torch / memset kernels (cotangent + gradient glue)" if ("at::" in s or "elementwise" in s or "Memset" in s or "reduce" in s) else s


# ---- launch list of one bench step
rows, ix = rows_of(os.path.join(G, "r02_launches.csv"))
agg = collections.OrderedDict()
for r in rows:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
    a = agg.setdefault(short(r[ix["Kernel Name"]]), [0, 0.0, r[ix["Block Size"]], r[ix["Grid Size"]]])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
md = ["# r02: ncu launch list of `python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline` (B=4096, T=200)", "",
      "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400`; the list covers the warm-up step, the timed step and the",
      "two e2e steps (4 passes of the hot path).  Cold-cache, serialised: compare SHARES with the bench line, not absolutes.", "",
      "| kernel | launches | total ms | share | block | grid |", "|---|---:|---:|---:|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |")
open(os.path.join(P, "r02_launches.md"), "w").write("\n".join(md) + "\n")

# ---- counters of the timed step's kernels at the bench size
rows, ix = rows_of(os.path.join(G, "r02_counters.csv"))
per = collections.OrderedDict()
for r in rows:
    k = short(r[ix["Kernel Name"]]) + " #" + r[ix["ID"]]
    per.setdefault(k, {})[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
fwd = [v for k, v in per.items() if k.startswith("fwd_kernel")][0]
flops = 2 * fwd["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"][0] + fwd["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"][0] + \
    fwd["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"][0]
dram = fwd["dram__bytes_read.sum"][0] + fwd["dram__bytes_write.sum"][0]
cap = {"source_sha": source_sha(), "workload": "push", "B": B, "T": T,
       "capture": "ncu --metrics ... --clock-control none of `python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline`, the timed step's launches (tools/gpu_ncu_r02.sh)",
       "fwd_kernel_dram_bytes_per_launch": dram, "fwd_kernel_dram_bytes_per_env_step": dram / (B * T),
       "fwd_kernel_fp64_flops_per_launch": flops, "fwd_kernel_fp64_flops_per_env_step": flops / (B * T),
       "fwd_kernel_ms_under_ncu": fwd["gpu__time_duration.sum"][0] / 1e6,
       "kernels": {k: {m: v[0] for m, v in d.items()} for k, d in per.items()}}
json.dump(cap, open(os.path.join(P, "r02_counters.json"), "w"), indent=1)

md = ["# r02: ncu counters of the kernels of one bench step (TactilePush 32x13, B=4096, T=200)", "",
      "`tools/gpu_ncu_r02.sh`: `ncu --metrics <list> --clock-control none -k regex:'fwd_kernel|tape_kernel|tac_kernel|vjp_kernel|bwd_kernel' --launch-skip 6 -c 6`",
      "on `python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline` (the six launches of the timed step).  Kernel sources: sha " + source_sha() + ".", ""]
names = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
         "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
         "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
         "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sector_hit_rate.pct",
         "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
stalls = [n for n in fwd if "issue_stalled" in n]
for k, d in per.items():
    md += [f"## {k}", "", "| metric | value | unit |", "|---|---:|---|"]
    for n in names:
        if n in d:
            md.append(f"| {n} | {d[n][0]:,.6g} | {d[n][1]} |")
    tots = sum(d[s][0] for s in stalls if s in d)
    if tots:
        mix = sorted(((d[s][0] / tots, s.split("issue_stalled_")[1].split("_per_")[0]) for s in stalls if s in d), reverse=True)
        md.append("| warp stall mix (per issue) | " + ", ".join(f"{nm} {100 * f:.1f}%" for f, nm in mix) + " | |")
    f64 = 2 * d.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", (0,))[0] + d.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", (0,))[0] + \
        d.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", (0,))[0]
    ms = d["gpu__time_duration.sum"][0] / 1e6
    md.append(f"| fp64 flops (2 dfma + dmul + dadd, thread level) | {f64:,.4g} | = {f64 / ms / 1e9:.3f} TFLOP/s under ncu |")
    md.append(f"| DRAM bytes per env-step | {(d['dram__bytes_read.sum'][0] + d['dram__bytes_write.sum'][0]) / (B * T):,.0f} | byte |")
    md.append("")
# ---- source-level attribution of the ncu --set full capture (T=20 per launch), per source function
rep = os.path.join(G, "r02_fwd_full.ncu-rep")
if os.path.exists(rep):
    src = os.path.join(G, "r02_fwd_src.csv")
    with open(src, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
    tab = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), src,
                          os.path.join(ROOT, "tactilesimulation_b200", "libtactilesim_b200.so"), "fwd_kernelILi8ELb0"],
                         capture_output=True, text=True).stdout
    md += ["## fwd_kernel<8, 0>: warp-state samples per source function", "",
           "`ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -c 1` on `tools/perf_probe.py --B 4096 --T 20 --lanes 8",
           "--grad-only`, source page joined with nvdisasm line info (`tools/ncu_by_function.py`).  `kernels.cu` = the block-wide vote of a",
           "round; `gp_points_coop` = the cooperative contact-point phase (its barriers); `mkdual` = inlined dual-number arithmetic.", "", "```"]
    md += tab.strip().split("\n")[:34] + ["```", ""]
open(os.path.join(P, "r02_ncu_summary.md"), "w").write("\n".join(md) + "\n")
print("fwd flops/launch %.4g, dram %.4g B, sha %s" % (flops, dram, source_sha()))
