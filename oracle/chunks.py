"""Chunk grid + crop-and-place restated.  TEST INFRASTRUCTURE (oracle/__init__.py).

Follows predict.py:38-61 (`make_chunks`) and predict.py:64-96
(`process_chunks`).  Pinned against the verbatim reference in
tests/test_oracle_vs_reference.py and by tests/golden/chunk_grids.json.
"""
import itertools

import numpy as np


def axis_grid(arr, chk, mrg):
    """One axis of make_chunks (predict.py:44-58): chunk starts step by
    chk-2*mrg, the last start is clamped to arr-chk (and dropped if it repeats
    the previous one); each chunk keeps [mrg, chk-mrg) except the first, which
    keeps from 0, and the last, which keeps whatever is still uncovered."""
    if arr < chk:
        raise ValueError(f'array extent {arr} smaller than chunk extent {chk}')
    starts = list(range(0, arr - 2 * mrg, chk - 2 * mrg))
    starts[-1] = arr - chk
    if len(starts) > 1 and starts[-1] == starts[-2]:
        starts.pop()
    crops = [[mrg, chk - mrg] for _ in starts]
    crops[0][0] = 0
    covered = sum(c[1] - c[0] for c in crops[:-1])
    crops[-1] = [chk - (arr - covered), chk]
    return starts, [tuple(c) for c in crops]


def make_chunks(arr_shape, chunk_shape, margin):
    ndim = len(arr_shape)
    if isinstance(margin, int):
        margin = [margin] * ndim
    per_axis = [axis_grid(arr_shape[d], chunk_shape[d], margin[d]) for d in range(ndim)]
    starts = list(itertools.product(*[p[0] for p in per_axis]))
    crops = list(itertools.product(*[p[1] for p in per_axis]))
    return starts, crops


def process_chunks(input_volume, chunk_size, output_volume, margin, fn):
    """fn(chunk (cz,cy,cx) f32) -> (C,cz,cy,cx); the cropped interior of every
    chunk's prediction is written to output_volume (C,Z,Y,X) (predict.py:81-95)."""
    starts, crops = make_chunks(input_volume.shape[-3:], chunk_size, margin)
    for st, cr in zip(starts, crops):
        sl = tuple(slice(s, s + c) for s, c in zip(st, chunk_size))
        pred = fn(input_volume[sl])
        crs = tuple(slice(a, b) for a, b in cr)
        output_volume[(slice(None),) + sl][(slice(None),) + crs] = pred[(slice(None),) + crs]
    return output_volume
