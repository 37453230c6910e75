"""The C-ABI library loads and exports every symbol include/tactilesim_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tactilesim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsim_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = _declared_symbols()
    for s in ["tsim_scene_create", "tsim_scene_destroy", "tsim_scene_sizes", "tsim_forward", "tsim_readout",
              "tsim_backward", "tsim_last_error"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from tactilesimulation_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), s
    assert sorted(_lib.SYMBOLS) == _declared_symbols()


def test_scene_create_rejects_garbage_without_touching_the_gpu():
    import numpy as np
    from tactilesimulation_b200 import _lib
    lib = _lib.load()
    ib = np.zeros(64, dtype=np.int32)
    db = np.zeros(64, dtype=np.float64)
    h = ctypes.c_void_p()
    rc = lib.tsim_scene_create(ib.ctypes.data, ib.size, db.ctypes.data, db.size, 0, ctypes.byref(h))
    assert rc != 0
    assert b"scene blob" in lib.tsim_last_error()


def test_product_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from tactilesimulation_b200 import _lib
    from tactilesimulation_b200.sim import BatchedSim
    g = np.load(os.path.join(ROOT, "tests", "golden", "pusher13x10_stepsim_s0.npz"))
    with pytest.raises(_lib.TactileSimError):
        BatchedSim((g["ibuf"], g["dbuf"]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tactilesimulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "redmax_oracle" not in text, f
                assert "tests.emu" not in text and "libtsim_emu" not in text, f


def test_scene_create_rejects_truncated_and_corrupted_blobs():
    """lower_scene validates every count and section of the blob against the buffer sizes before indexing it."""
    import numpy as np
    from tactilesimulation_b200 import _lib
    lib = _lib.load()
    g = np.load(os.path.join(ROOT, "tests", "golden", "pusher13x10_stepsim_s0.npz"))
    ib0, db0 = g["ibuf"].astype(np.int32), g["dbuf"].astype(np.float64)

    def create(ib, db):
        h = ctypes.c_void_p()
        rc = lib.tsim_scene_create(ib.ctypes.data, ib.size, db.ctypes.data, db.size, 0, ctypes.byref(h))
        return rc, lib.tsim_last_error()
    rc, msg = create(ib0, np.ascontiguousarray(db0[: db0.size // 2]))            # truncated double buffer
    assert rc != 0 and b"malformed scene blob" in msg
    rc, msg = create(np.ascontiguousarray(ib0[:40]), db0)                          # truncated int buffer
    assert rc != 0 and b"malformed scene blob" in msg
    ib = ib0.copy()
    ib[2] = -3                                                                    # negative joint count
    rc, msg = create(ib, db0)
    assert rc != 0 and b"malformed scene blob" in msg
    ib = ib0.copy()
    ib[24] = 10 ** 8                                                              # joint data offset past the buffer
    rc, msg = create(ib, db0)
    assert rc != 0 and b"malformed scene blob" in msg
