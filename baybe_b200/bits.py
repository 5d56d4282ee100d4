"""Bit-packed candidate layout (BB_BITS_U8 of include/baybe_b200.h): numpy helpers."""
from __future__ import annotations

import numpy as np


def pack_bits(x01: np.ndarray) -> np.ndarray:
    """(N, d) 0/1 matrix -> (N, ceil(d/8)) uint8, feature j in bit (j & 7) of byte j >> 3."""
    return np.packbits(np.asarray(x01) != 0, axis=1, bitorder="little")


def unpack_bits(packed: np.ndarray, d: int) -> np.ndarray:
    """Inverse of pack_bits: (N, ceil(d/8)) uint8 -> (N, d) float64 0/1 matrix."""
    return np.unpackbits(np.asarray(packed, dtype=np.uint8), axis=1, bitorder="little")[:, :d].astype(np.float64)
