"""Collects the bench lines of the round's gpurun calls (gpurun_out/*.json) into profiles/r02_bench_lines.md."""
import json
import os

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
FILES = [("i_bench.json", "headline, 1 GPU: `python bench.py --gpus 1 --steps 20 --warmup 5` (the driver's flags)"),
         ("fin_ref.json", "reference arm: `python bench.py --impl reference --gpus 1 --steps 20 --warmup 5`"),
         ("h_bench_push_fwd.json", "configs[1]: `python bench.py --workload push_fwd --steps 5 --warmup 3`"),
         ("h_bench_dclaw.json", "configs[3]: `python bench.py --workload dclaw --steps 5 --warmup 3`"),
         ("h_bench_insertion.json", "configs[4] on 1 GPU (B=1024): `python bench.py --workload insertion --steps 5 --warmup 3`"),
         ("h_bench_stepsim.json", "gd.py shape: `python bench.py --workload stepsim --steps 5 --warmup 3`"),
         ("n8_push.json", "headline, 8 GPUs: `torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3`"),
         ("n8_insertion.json", "configs[4], 8 GPUs (8192 environments): `torchrun --nproc-per-node 8 bench.py --gpus 8 --workload insertion`"),
         ("n8_dclaw.json", "configs[3], 8 GPUs: `torchrun --nproc-per-node 8 bench.py --gpus 8 --workload dclaw`")]
md = ["# r02: bench lines measured during the round (gpurun, fresh B200 boxes; the driver's own runs are BENCH_r02 / SCALE_r02)", "",
      "| run | value | unit | ms/step | e2e | reference on the host cores (same run) | kernels (ms) |", "|---|---:|---|---:|---:|---:|---|"]
full = []
for f, what in FILES:
    p = os.path.join(G, f)
    if not os.path.exists(p):
        continue
    try:
        l = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        continue
    k = {n: round(v["ms"], 2) for n, v in l.get("roofline", {}).get("kernels", {}).items()}
    cpu = l.get("cpu_baseline") or {}
    md.append(f"| {what} | {l['value']:.4g} | {l['unit']} | {l['ms_per_step']:.1f} | {l['e2e']['value']:.4g} | "
              f"{(str(round(cpu['value'])) + ' (' + str(cpu['cores']) + ' cores)') if cpu else '-'} | {k if k else '-'} |")
    full.append((what, l))
md += ["", "## full JSON lines", ""]
for what, l in full:
    md += [f"### {what}", "", "```json", json.dumps(l), "```", ""]
open(os.path.join(P, "r02_bench_lines.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[:14]))
