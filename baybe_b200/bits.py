"""Bit-packed candidate layout (BB_BITS_U8 of include/baybe_b200.h): numpy helpers."""
from __future__ import annotations

import numpy as np


def pack_bits(x01: np.ndarray) -> np.ndarray:
    """(N, d) 0/1 matrix -> (N, ceil(d/8)) uint8, feature j in bit (j & 7) of byte j >> 3."""
    return np.packbits(np.asarray(x01) != 0, axis=1, bitorder="little")


def unpack_bits(packed: np.ndarray, d: int) -> np.ndarray:
    """Inverse of pack_bits: (N, ceil(d/8)) uint8 -> (N, d) float64 0/1 matrix."""
    return np.unpackbits(np.asarray(packed, dtype=np.uint8), axis=1, bitorder="little")[:, :d].astype(np.float64)


def encode_levels(x: np.ndarray) -> tuple[np.ndarray, np.ndarray, int]:
    """Level-coded form of a discrete comp-rep matrix: (codes, table, bits).

    ``table[j]`` lists the distinct float32 values of column j (ascending, zero padded to a common length) and
    ``codes`` holds, per row, the index of each column's value -- packed two per byte (low nibble = even column)
    when every column has at most 16 levels (bits = 4), one byte per column otherwise (bits = 8).  Decoding with
    ``bb_decode_codes`` reproduces ``x.astype(float32)`` exactly.  Raises if a column has more than 256 levels
    (then the space is not usefully level-coded; ship float32)."""
    x32 = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    n, d = x32.shape
    levels, idx = [], np.empty((n, d), dtype=np.uint8)
    for j in range(d):
        vals, inv = np.unique(x32[:, j], return_inverse=True)
        if len(vals) > 256:
            raise ValueError(f"column {j} has {len(vals)} distinct values: not a level-coded space")
        levels.append(vals)
        idx[:, j] = inv.astype(np.uint8)
    width = max(len(v) for v in levels)
    table = np.zeros((d, width), dtype=np.float32)
    for j, v in enumerate(levels):
        table[j, : len(v)] = v
    if width <= 16:
        if d % 2:
            idx = np.concatenate([idx, np.zeros((n, 1), dtype=np.uint8)], axis=1)
        return np.ascontiguousarray(idx[:, 0::2] | (idx[:, 1::2] << 4)), table, 4
    return idx, table, 8


def decode_levels(codes: np.ndarray, table: np.ndarray, bits: int, d: int) -> np.ndarray:
    """NumPy twin of ``bb_decode_codes`` (tests)."""
    codes = np.asarray(codes, dtype=np.uint8)
    if bits == 4:
        idx = np.empty((codes.shape[0], 2 * codes.shape[1]), dtype=np.uint8)
        idx[:, 0::2] = codes & 15
        idx[:, 1::2] = codes >> 4
        idx = idx[:, :d]
    else:
        idx = codes[:, :d]
    return np.take_along_axis(np.asarray(table, dtype=np.float32).T, idx.astype(np.int64), axis=0) if False else \
        np.stack([np.asarray(table, dtype=np.float32)[j][idx[:, j]] for j in range(d)], axis=1)
