"""Helpers shared by the DClaw parity tests: expected contact bitmask words from the golden per-force id lists."""
import numpy as np


def expected_words(scene, ground_ids_t, gp_ids_t):
    """ground_ids_t [forces, width], gp_ids_t [forces, width] (-1 padded) -> list of 32-bit words in the ABI's
    layout (include/tactilesim_b200.h: ground contacts first, then general-primitive contacts, ceil(points/32) words each)."""
    words = []
    for f, g in enumerate(scene.ground_contacts):
        npts = len(scene.contact_points[g["body"]])
        ws = [0] * ((npts + 31) // 32)
        for k in ground_ids_t[f]:
            if k >= 0:
                ws[int(k) >> 5] |= 1 << (int(k) & 31)
        words += ws
    for f, gp in enumerate(scene.gp_contacts):
        npts = len(scene.contact_points[gp["body1"]])
        ws = [0] * ((npts + 31) // 32)
        for k in gp_ids_t[f]:
            if k >= 0:
                ws[int(k) >> 5] |= 1 << (int(k) & 31)
        words += ws
    return np.array(words, dtype=np.uint64)
