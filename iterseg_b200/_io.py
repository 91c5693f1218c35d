"""Label output: OME-zarr v0.4 label stores, written without the zarr package.

Mirrors the pieces of src/iterseg/_io.py the segmentation path calls:
`napari_to_ome` (:99-138), `save_labels_to_ome` (:142-166), `open_zarr` (:325-386).
zarr / ome-zarr are not dependencies here: the store is written directly in the
zarr v2 directory layout (".zgroup" / ".zattrs" / ".zarray" + one raw C-order
file per chunk, no compressor), which any zarr v2 reader opens.  The metadata
is what the reference writes: group attribute `image-label: {}`, `multiscales`
v0.4 with axes t/z/y/x (seconds / micrometers) and the `scale` + `translate`
coordinate transformations on dataset "0" (the reference writes the type
string 'translate', _io.py:128-131; so do we).
"""
import json
import os
import pathlib

import numpy as np


def napari_to_ome(layer_meta):
    scale = list(map(float, layer_meta['scale']))
    translate = list(map(float, layer_meta['translate']))
    ndim = len(scale)
    axes = [{'name': 't', 'type': 'time', 'unit': 'second'},
            {'name': 'z', 'type': 'space', 'unit': 'micrometer'},
            {'name': 'y', 'type': 'space', 'unit': 'micrometer'},
            {'name': 'x', 'type': 'space', 'unit': 'micrometer'}][-ndim:]
    coordtfs = [{'type': 'scale', 'scale': scale}, {'type': 'translate', 'translate': translate}]
    datasets = [{'coordinateTransformations': coordtfs, 'path': '0'}]
    return {'datasets': datasets, 'axes': axes, 'name': layer_meta['name']}


def normalize_chunks(chunks, shape):
    """zarr v2's rule for under-specified chunks: missing trailing dimensions span the
    whole axis -- so the reference's chunks=chunk_size (3 values) on tzyx data chunks
    (t,z,y) and keeps x whole (segmentation.py:776-782)."""
    chunks = tuple(int(c) for c in chunks)
    if len(chunks) > len(shape):
        raise ValueError('too many dimensions in chunks')
    chunks = chunks + tuple(int(s) for s in shape[len(chunks):])
    return tuple(min(c, s) if s > 0 else c for c, s in zip(chunks, shape))


class LabelArray:
    """Chunked integer array: zarr-v2 directory store (path given) or in memory."""

    def __init__(self, shape, chunks, dtype=np.int32, path=None):
        self.path = None if path is None else str(path)
        self._mem = None
        self.sep = '.'
        meta_fn = None if self.path is None else os.path.join(self.path, '.zarray')
        if meta_fn is not None and os.path.exists(meta_fn):
            # an existing store is re-opened with ITS OWN metadata, like zarr.open(mode='a') in the
            # reference's open_zarr (_io.py:325-386): a warm restart must find the frames where
            # the first run put them, whatever chunk_size this call was given
            with open(meta_fn) as f:
                meta = json.load(f)
            if meta.get('zarr_format') != 2:
                raise NotImplementedError(f'{self.path}: only zarr format 2 stores are supported')
            if meta.get('compressor') is not None or meta.get('filters'):
                raise NotImplementedError(
                    f'{self.path} was written with compressor {meta.get("compressor")!r} / filters '
                    f'{meta.get("filters")!r}: this writer reads and writes raw (uncompressed) chunks only; '
                    'resume with the tool that created the store or start a fresh one')
            if meta.get('order', 'C') != 'C':
                raise NotImplementedError(f'{self.path}: only C-order chunks are supported')
            if shape is not None and tuple(int(v) for v in shape) != tuple(meta['shape']):
                raise ValueError(f'{self.path} holds an array of shape {tuple(meta["shape"])}, '
                                 f'not the requested {tuple(shape)}')
            self.shape = tuple(int(v) for v in meta['shape'])
            self.chunks = tuple(int(v) for v in meta['chunks'])
            self.dtype = np.dtype(meta['dtype'])
            self.sep = meta.get('dimension_separator', '.')
            self.ndim = len(self.shape)
            return
        self.shape = tuple(int(s) for s in shape)
        self.chunks = normalize_chunks(chunks, self.shape)
        self.dtype = np.dtype(dtype)
        self.ndim = len(self.shape)
        if self.path is None:
            self._mem = np.zeros(self.shape, dtype=self.dtype)
        else:
            os.makedirs(self.path, exist_ok=True)
            meta = {'zarr_format': 2, 'shape': list(self.shape), 'chunks': list(self.chunks),
                    'dtype': self.dtype.newbyteorder('<').str, 'compressor': None,
                    'fill_value': 0, 'order': 'C', 'filters': None,
                    'dimension_separator': '.'}
            tmp = meta_fn + f'.tmp{os.getpid()}'
            with open(tmp, 'w') as f:
                json.dump(meta, f, indent=1)
            os.replace(tmp, meta_fn)                     # atomic: several ranks may create the same store

    # ---- chunk files ----------------------------------------------------------
    def _chunk_fn(self, idx):
        fn = os.path.join(self.path, self.sep.join(str(i) for i in idx))
        if self.sep == '/':
            os.makedirs(os.path.dirname(fn), exist_ok=True)
        return fn

    def _read_chunk(self, idx):
        fn = self._chunk_fn(idx)
        if not os.path.exists(fn):
            return np.zeros(self.chunks, dtype=self.dtype)
        a = np.fromfile(fn, dtype=self.dtype.newbyteorder('<'))
        n = int(np.prod(self.chunks))
        if a.size < n:                                   # a chunk file that is still being extended
            a = np.concatenate([a, np.zeros(n - a.size, dtype=a.dtype)])
        return a[:n].reshape(self.chunks)

    def _update_chunk(self, idx, src, block):
        """Write `block` into region `src` of chunk `idx` IN PLACE (no read-modify-write of the
        whole chunk file): raw C-order chunks have a fixed byte layout, so the region is stored
        through a shared mapping of the file.  Processes that write DISJOINT regions of one chunk
        file (ranks that own different frames of a t-chunk, distributed.shard_frames) never
        touch each other's bytes; a file that does not exist yet is created sparse, its holes
        read as the fill value 0."""
        fn = self._chunk_fn(idx)
        nbytes = int(np.prod(self.chunks)) * self.dtype.itemsize
        fd = os.open(fn, os.O_RDWR | os.O_CREAT, 0o644)
        try:
            if os.fstat(fd).st_size < nbytes:
                os.ftruncate(fd, nbytes)                 # extending keeps what other writers stored
        finally:
            os.close(fd)
        # leading axes that are indexed, trailing axes that are taken whole: every index tuple of the
        # leading part is ONE contiguous byte run of the file -> positional writes (pwrite), which
        # fill the page cache in bulk (a frame of a tzyx t-chunk is a single 17 MB run); anything
        # else goes through a shared mapping
        k = self.ndim
        while k > 0 and src[k - 1].start == 0 and src[k - 1].stop == self.chunks[k - 1]:
            k -= 1
        k = min(k + 1, self.ndim)                       # the innermost indexed axis is contiguous too
        lead = [range(sl.start, sl.stop) for sl in src[:k - 1]]
        n_runs = int(np.prod([len(r) for r in lead])) if lead else 1
        if n_runs <= 64:
            import itertools
            blk = np.ascontiguousarray(block, dtype=self.dtype.newbyteorder('<'))
            strides = [int(np.prod(self.chunks[i + 1:])) * self.dtype.itemsize for i in range(self.ndim)]
            fd = os.open(fn, os.O_RDWR)
            try:
                for idx_lead in itertools.product(*lead):
                    off = sum(i * st for i, st in zip(idx_lead, strides)) + src[k - 1].start * strides[k - 1]
                    sub = blk[tuple(i - sl.start for i, sl in zip(idx_lead, src[:k - 1]))]
                    view = memoryview(np.ascontiguousarray(sub)).cast('B')
                    done = 0
                    while done < len(view):
                        done += os.pwrite(fd, view[done:], off + done)
            finally:
                os.close(fd)
            return
        mm = np.memmap(fn, dtype=self.dtype.newbyteorder('<'), mode='r+', shape=self.chunks)
        mm[src] = block
        mm.flush()
        del mm

    def _norm_key(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if Ellipsis in key:
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        out, squeeze = [], []
        for k, s in zip(key, self.shape):
            if isinstance(k, (int, np.integer)):
                k = int(k) + (s if k < 0 else 0)
                out.append((k, k + 1))
                squeeze.append(True)
            else:
                a, b, st = k.indices(s)
                if st != 1:
                    raise NotImplementedError('strided access')
                out.append((a, max(a, b)))
                squeeze.append(False)
        return out, squeeze

    def __getitem__(self, key):
        if self._mem is not None:
            return self._mem[key]
        rng, squeeze = self._norm_key(key)
        res = np.zeros([b - a for a, b in rng], dtype=self.dtype)
        n_chunk = int(np.prod(self.chunks))
        for idx, src, dst in self._overlaps(rng):
            fn = self._chunk_fn(idx)
            if not os.path.exists(fn):
                continue                                 # fill value 0
            if os.path.getsize(fn) >= n_chunk * self.dtype.itemsize:
                # only the pages of the requested region are read (a t-chunk of the reference's
                # chunks=chunk_size layout holds 10 frames, segmentation.py:776-782)
                mm = np.memmap(fn, dtype=self.dtype.newbyteorder('<'), mode='r', shape=self.chunks)
                res[dst] = mm[src]
                del mm
            else:
                res[dst] = self._read_chunk(idx)[src]
        return res.reshape([n for n, sq in zip(res.shape, squeeze) if not sq])

    def __setitem__(self, key, value):
        if self._mem is not None:
            self._mem[key] = value
            return
        rng, squeeze = self._norm_key(key)
        full = [b - a for a, b in rng]
        kept = [n for n, sq in zip(full, squeeze) if not sq]
        value = np.broadcast_to(np.asarray(value).astype(self.dtype, copy=False), kept).reshape(full)
        for idx, src, dst in self._overlaps(rng):
            whole = all(s.start == 0 and s.stop == c for s, c in zip(src, self.chunks))
            if whole:
                np.ascontiguousarray(value[dst]).astype(self.dtype.newbyteorder('<'), copy=False) \
                    .tofile(self._chunk_fn(idx))
            else:
                self._update_chunk(idx, src, value[dst])

    def _overlaps(self, rng):
        import itertools
        spans = []
        for (a, b), c in zip(rng, self.chunks):
            spans.append(range(a // c, (max(b, a + 1) - 1) // c + 1) if b > a else range(0))
        for idx in itertools.product(*spans):
            src, dst = [], []
            for i, (a, b), c in zip(idx, rng, self.chunks):
                lo, hi = max(a, i * c), min(b, (i + 1) * c)
                src.append(slice(lo - i * c, hi - i * c))
                dst.append(slice(lo - a, hi - a))
            yield idx, tuple(src), tuple(dst)

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.shape[0]


def open_zarr(path, *, shape=None, chunks=None, dtype=None, **kwargs):
    """_io.py:325-386: open (mode 'a') the array store at `path`; an existing store keeps its own
    shape / chunks / dtype (shape is checked when given)."""
    return LabelArray(shape, chunks, dtype=dtype if dtype is not None else np.int32, path=path)


def zeros(shape, chunks, dtype=np.int32):
    """Stand-in for zarr.zeros(...) (segmentation.py:784-786): an in-memory store."""
    return LabelArray(shape, chunks, dtype=dtype, path=None)


def save_labels_to_ome(path, data=None, layer_meta=None, shape=None, chunks=None, dtype=np.uint32):
    path = pathlib.Path(path)
    if data is None and (shape is None or chunks is None):
        raise ValueError('either data or shape/chunks must be provided')
    os.makedirs(path, exist_ok=True)
    with open(path / '.zgroup', 'w') as f:
        json.dump({'zarr_format': 2}, f)
    metadata = napari_to_ome(layer_meta)
    attrs = {'image-label': {},
             'multiscales': [{'version': '0.4', 'name': metadata['name'], 'axes': metadata['axes'],
                              'datasets': metadata['datasets']}]}
    with open(path / '.zattrs', 'w') as f:
        json.dump(attrs, f, indent=1)
    if data is not None:
        shape, dtype = data.shape, data.dtype
        if chunks is None:
            chunks = getattr(data, 'chunks', None) or (1,) * (data.ndim - 2) + tuple(data.shape[-2:])
    arr = open_zarr((path / '0').as_posix(), shape=shape, chunks=chunks, dtype=dtype)
    if data is not None:
        arr[...] = np.asarray(data)
    return arr
