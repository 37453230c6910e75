"""ctypes binding of libtactilesim_b200.so (C ABI: include/tactilesim_b200.h).

There is no CPU fallback: if the CUDA library is missing or no CUDA device is visible, every
entry point raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TSIM_B200_LIB: development override to A/B-test kernel build variants (still a CUDA library of this ABI)
LIB_PATH = os.environ.get("TSIM_B200_LIB") or os.path.join(HERE, "libtactilesim_b200.so")

SYMBOLS = ["tsim_last_error", "tsim_scene_create", "tsim_scene_destroy", "tsim_scene_sizes", "tsim_scene_set_lanes", "tsim_scene_set_option",
           "tsim_scene_set_env_scenes",
           "tsim_forward", "tsim_forward_multistep", "tsim_scene_kernel_times", "tsim_readout", "tsim_backward",
           "tsim_debug_fp64_peak", "tsim_debug_last_error", "tsim_debug_lu_solve"]
KERNELS = ("fwd_kernel", "tape_kernel", "tac_kernel", "vjp_kernel", "bwd_kernel")
(NJ, NDOF_R, NDOF_M, NDOF_U, NDOF_VAR, NDOF_TACTILE, N_MARKERS, TAPE_DOUBLES, CMASK_WORDS, INTEGRATOR, N_SIZES) = range(11)
INT_BDF1, INT_BDF2, INT_SDIRK2 = 0, 1, 2

_lib = None


class TactileSimError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TactileSimError(
            f"{LIB_PATH} is missing: build the sm_100a library first (python -m tactilesimulation_b200.build "
            "or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.tsim_last_error.restype = ctypes.c_char_p
    lib.tsim_scene_create.argtypes = [vp, i64, vp, i64, ctypes.c_int, ctypes.POINTER(vp)]
    lib.tsim_scene_destroy.argtypes = [vp]
    lib.tsim_scene_destroy.restype = None
    lib.tsim_scene_sizes.argtypes = [vp, vp]
    lib.tsim_scene_set_lanes.argtypes = [vp, ctypes.c_int]
    lib.tsim_scene_set_option.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    lib.tsim_scene_set_env_scenes.argtypes = [vp, i32, vp, i64, vp, i64]
    lib.tsim_forward.argtypes = [vp, i32, i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.tsim_forward_multistep.argtypes = [vp, i32, i32, vp, vp, vp, vp, i32, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.tsim_scene_kernel_times.argtypes = [vp, vp]
    lib.tsim_readout.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.tsim_backward.argtypes = [vp, i32, i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.tsim_debug_fp64_peak.argtypes = [ctypes.c_int, vp]
    lib.tsim_debug_last_error.restype = ctypes.c_char_p
    lib.tsim_debug_lu_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]
    _lib = lib
    return lib


def fp64_peak(device: int = 0):
    """Measured fp64 FMA peak of the device: dict(gflops, ms, sms, sm_mhz).  tsim_debug_fp64_peak."""
    import numpy as np
    lib = load()
    out = np.zeros(4, dtype=np.float64)
    if lib.tsim_debug_fp64_peak(int(device), out.ctypes.data) != 0:
        raise TactileSimError(lib.tsim_debug_last_error().decode("utf-8", "replace"))
    return dict(gflops=float(out[0]), ms=float(out[1]), sms=int(out[2]), sm_mhz=float(out[3]))


def lu_solve(A, b, device: int = 0):
    """The kernels' n x n solve (row-owner elimination with partial pivoting) on a batch of systems: A [nsys, n, n],
    b [nsys, n] with n = 8 or 16 -> x [nsys, n].  Test aid, tsim_debug_lu_solve."""
    import numpy as np
    lib = load()
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    nsys, n = b.shape
    if A.shape != (nsys, n, n):
        raise ValueError("A must be [nsys, n, n] for b [nsys, n]")
    x = np.zeros_like(b)
    check(lib.tsim_debug_lu_solve(int(n), int(device), int(nsys), A.ctypes.data, b.ctypes.data, x.ctypes.data), lib)
    return x


def check(rc, lib):
    if rc != 0:
        raise TactileSimError(lib.tsim_last_error().decode("utf-8", "replace"))
