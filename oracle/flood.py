"""Python face of the C flood oracle + a pure-Python heapq restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).
Restates watershed.py:17-92 (`affinity_watershed`, `_prep_data`,
`_indices_to_raveled_affinities`) and wraps flood.c (watershed.py:95-159).
"""
import ctypes
import heapq
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libisg_oracle.so')
_lib = None


def build(force=False):
    """gcc -O2 the C restatement into oracle/_build/ (git-ignored)."""
    src = os.path.join(_HERE, 'flood.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-o', _SO, src])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.isg_oracle_flood.restype = ctypes.c_int64
        _lib.isg_oracle_flood.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def neighbor_table(shape):
    """(2*ndim, 2) rows (affinity axis, flat offset): the 6-connected offsets in
    the order [-YX, -X, -1, +1, +X, +YX] paired with axes [0,1,2,2,1,0]
    (watershed.py:84-92 over skimage's _offsets_to_raveled_neighbors)."""
    ndim = len(shape)
    strides = [int(np.prod(shape[i + 1:])) for i in range(ndim)]
    offs = [-s for s in strides] + [s for s in reversed(strides)]
    axes = list(range(ndim)) + list(range(ndim))[::-1]
    return np.array(list(zip(axes, offs)), dtype=np.int64)


def ravel_seeds(coords, shape):
    coords = np.asarray(coords, dtype=np.int64).reshape(-1, len(shape))
    strides = np.array([int(np.prod(shape[i + 1:])) for i in range(len(shape))],
                       dtype=np.int64)
    return coords @ strides


def raveled_flood_c(image_raveled, seeds, offsets, mask, output):
    """C oracle; `output` (uint32, flat) is modified in place."""
    lib = _load()
    image_raveled = np.ascontiguousarray(image_raveled, dtype=np.float32)
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    mask8 = np.ascontiguousarray(mask).view(np.uint8) if mask.dtype == bool \
        else np.ascontiguousarray(mask, dtype=np.uint8)
    assert output.dtype == np.uint32 and output.flags.c_contiguous
    age = lib.isg_oracle_flood(
        image_raveled.ctypes.data, image_raveled.shape[1],
        seeds.ctypes.data, len(seeds), offsets.ctypes.data, len(offsets),
        mask8.ctypes.data, output.ctypes.data)
    if age < 0:
        raise MemoryError('oracle flood: heap allocation failed')
    return age


def raveled_flood_py(image_raveled, seeds, offsets, mask, output):
    """Pure-Python heapq restatement (small cases only)."""
    npix = image_raveled.shape[1]
    half = len(offsets) // 2
    heap = [(np.float32(0.0), 0, int(s)) for s in seeds]
    heapq.heapify(heap)
    age = 0
    while heap:
        _, _, p = heapq.heappop(heap)
        for i, (axis, off) in enumerate(offsets):
            nb = p + int(off)
            if nb < 0 or nb >= npix or not mask[nb] or output[nb]:
                continue
            output[nb] = output[p]
            age += 1
            aoff = 0 if i < half else int(off)
            heapq.heappush(heap, (image_raveled[axis, aoff + p], age, nb))
    return age


def affinity_watershed(image, marker_coords, mask=None, scale=None, out=None,
                       impl='c'):
    """Restatement of affinity_watershed (watershed.py:17-35) + _prep_data
    (:38-63): image (3,Z,Y,X) f32, marker_coords (N,3), mask (Z,Y,X) bool."""
    shape = image.shape[1:]
    raveled = np.stack([image[i].ravel() for i in range(image.shape[0])]).astype(
        image.dtype, copy=True)
    if scale is not None:
        raveled *= np.abs(np.asarray(scale, dtype=raveled.dtype)).reshape(-1, 1)
    if mask is None:
        mask = np.pad(np.ones([s - 2 for s in shape], dtype=bool), 1,
                      constant_values=False)
    seeds = ravel_seeds(marker_coords, shape)
    output = np.zeros(mask.size, dtype=np.uint32) if out is None else out
    output[seeds] = np.arange(1, len(seeds) + 1, dtype=output.dtype)
    fn = raveled_flood_c if impl == 'c' else raveled_flood_py
    if output.dtype != np.uint32:
        tmp = output.astype(np.uint32)
        fn(raveled, seeds, neighbor_table(shape), mask.ravel(), tmp)
        output[:] = tmp
    else:
        fn(raveled, seeds, neighbor_table(shape), mask.ravel(), output)
    return output.reshape(shape)
