"""Shared helpers for the parity tests (the seeded frame generators live in imageflow_b200.synth so that
bench.py and the tests produce identical bytes)."""
from __future__ import annotations

import numpy as np

from imageflow_b200.synth import gradient_np as gradient  # noqa: F401
from imageflow_b200.synth import noise_np as noise  # noqa: F401


def padded(a: np.ndarray, align: int = 64) -> np.ndarray:
    """Copy into a buffer whose row stride is padded to `align` bytes (bitmaps.rs:803-804); returns the (h,w,4) view."""
    h, w, _ = a.shape
    stride = (w * 4 + align - 1) // align * align
    buf = np.zeros((h, stride), np.uint8)
    view = np.lib.stride_tricks.as_strided(buf, shape=(h, w, 4), strides=(stride, 4, 1))
    view[...] = a
    return view


def diff_stats(a: np.ndarray, b: np.ndarray):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return int(d.max()), int((d > 0).sum())
