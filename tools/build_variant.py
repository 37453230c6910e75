"""Development tool: build an A/B variant of the library with extra nvcc flags for the variant-8 kernels (TactilePush).

    python tools/build_variant.py NAME -DTS_ROUNDS_PER_VOTE=1 -DTS_TILE_STEPS=0 ...

Compiles csrc/kernels_v8.cu with the extra flags and links it with the stock v16 / v17 / cabi objects into
tactilesimulation_b200/_variants/NAME.so (git-ignored; travels to the GPU box).  Select it with TSIM_B200_LIB=<path>
(tools/gpu_variants.sh).  --all-variants recompiles v16 / v17 with the flags too."""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from tactilesimulation_b200 import build as b  # noqa: E402


def main():
    name, extra = sys.argv[1], [x for x in sys.argv[2:] if x != "--all-variants"]
    allv = "--all-variants" in sys.argv
    b.build()                                   # stock objects
    out = os.path.join(b.HERE, "_variants")
    os.makedirs(out, exist_ok=True)
    objs, jobs = [], []
    for v in b.VARIANTS:
        if v == 8 or allv:
            o = os.path.join(out, f"{name}_v{v}.o")
            jobs.append(subprocess.Popen(["nvcc"] + b.NVCC_FLAGS + extra + ["-c", f"kernels_v{v}.cu", "-o", o], cwd=b.CSRC))
        else:
            o = os.path.join(b.OBJ, f"kernels_v{v}.o")
        objs.append(o)
    if any(j.wait() for j in jobs):
        raise SystemExit("nvcc failed")
    objs.append(os.path.join(b.OBJ, "cabi.o"))
    objs.append(os.path.join(b.OBJ, "microbench.o"))
    lib = os.path.join(out, name + ".so")
    subprocess.check_call(["nvcc", "-shared", "-o", lib] + objs, cwd=b.CSRC)
    print(lib)


if __name__ == "__main__":
    main()
