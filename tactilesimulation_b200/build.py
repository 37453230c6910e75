"""In-tree build of the CUDA library (sm_100a only).  nvcc cross-compiles without a GPU.

The kernels are compiled three times from csrc/kernels.cu (compile-time capacities and feature gates,
csrc/kernel_layout.h): variants 8, 16 and 17; csrc/cabi.cpp owns the public C ABI and dispatches per scene;
csrc/microbench.cu is the fp64 peak measurement aid.  The objects compile in parallel."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtactilesim_b200.so")
OBJ = os.path.join(HERE, "_obj")
VARIANTS = (8, 16, 17)
DEPS = ["microbench.cu", "kernels.cu", "kernels_v8.cu", "kernels_v16.cu", "kernels_v17.cu", "cabi.cpp", "sim_core.cuh", "dual.cuh", "scene_layout.h", "kernel_layout.h", "scene_lower.h",
        os.path.join("..", "..", "include", "tactilesim_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def source_sha() -> str:
    """Hash of the kernel sources: profiles/*.json captured under ncu carry it, so that bench.py quotes a captured figure
    (DRAM traffic, flop counts) only for the code it was captured on."""
    import hashlib
    h = hashlib.sha256()
    for d in sorted(DEPS):
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    flags = NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else [])
    jobs, objs = [], []
    for v in VARIANTS:
        o = os.path.join(OBJ, f"kernels_v{v}.o")
        objs.append(o)
        jobs.append(subprocess.Popen([nvcc] + flags + ["-c", f"kernels_v{v}.cu", "-o", o], cwd=CSRC))
    o = os.path.join(OBJ, "microbench.o")
    objs.append(o)
    jobs.append(subprocess.Popen([nvcc] + flags + ["-c", "microbench.cu", "-o", o], cwd=CSRC))
    o = os.path.join(OBJ, "cabi.o")
    objs.append(o)
    jobs.append(subprocess.Popen([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-c", "cabi.cpp", "-o", o], cwd=CSRC))
    rcs = [j.wait() for j in jobs]
    if any(rcs):
        raise RuntimeError(f"nvcc failed (exit codes {rcs})")
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
