"""TEST INFRASTRUCTURE - numpy restatement of the output side of a generation run (never imported by the product package).

reference ``utils/callbacks.py:52-57``:  sample.continuous = sample.continuous * std + mean;  sample.apply_mask()
reference ``utils/tensorclass.py:97-108, 197-201``:  apply_mask zeroes continuous and discrete at padded slots; save_to writes
the datasets ``time``, ``continuous``, ``discrete``, ``mask``.

``pack_records`` produces the per-jet record the CUDA kernel ``mmf_pack_sample`` must match byte for byte (layout in
include/mmf_b200.h): [D][3] f32 de-standardised kinematics, zero at pads | [D] u8 token | mask << 7 | zero padding to 16 bytes.
"""
from __future__ import annotations

import numpy as np


def record_bytes(D: int) -> int:
    return (D * 13 + 15) // 16 * 16


def postprocess(continuous, discrete, mask, mean=None, std=None):
    """What FlowGeneratorCallback._gather_results_global does to the gathered sample (callbacks.py:52-57)."""
    x = np.asarray(continuous, np.float32)
    m = np.asarray(mask).reshape(x.shape[0], x.shape[1], 1)
    if mean is not None or std is not None:
        sig = np.asarray(std if std is not None else [1, 1, 1], np.float32)
        mu = np.asarray(mean if mean is not None else [0, 0, 0], np.float32)
        x = (x * sig).astype(np.float32) + mu
    x = (x * (m != 0)).astype(np.float32)
    k = None if discrete is None else (np.asarray(discrete).reshape(m.shape) * (m != 0)).astype(np.int64)
    return x, k, (m != 0).astype(np.int64)


def pack_records(continuous, discrete, mask, mean=None, std=None) -> np.ndarray:
    x, k, m = postprocess(continuous, discrete, mask, mean, std)
    B, D = x.shape[:2]
    rec = np.zeros((B, record_bytes(D)), np.uint8)
    rec[:, : D * 12] = np.where(m != 0, x, np.float32(0.0)).astype(np.float32).reshape(B, D * 3).view(np.uint8)
    tok = np.zeros((B, D), np.uint8) if k is None else (k[..., 0] & 0x7F).astype(np.uint8)
    rec[:, D * 12: D * 13] = tok | (m[..., 0].astype(np.uint8) << 7)
    return rec
