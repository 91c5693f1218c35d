"""Spatial (z-slab) sharding of ONE large volume over the GPUs of one box
(BASELINE.json configs[3]: 256x2048x2048 over 8 B200; SURVEY.md section 8e, second row).

The reference is single-process: `segment_single_volume` (segmentation.py:885-900) runs the
whole volume through `affinity_watershed_for_chunks` (:147-195).  Here every rank owns a
contiguous range of z planes and the result is IDENTICAL to the single-device run:

 1. U-Net: the GLOBAL `make_chunks` list (predict.py:38-61) is partitioned by the z interval
    of each chunk's cropped interior, so every chunk (its own BatchNorm batch) is computed
    exactly once, by the rank that owns the planes it writes.  No collective.
 2. Halo: each rank receives `halo` feature planes from both neighbours (NCCL send/recv over
    NVLink).  The sigma=2 smoothing behind the Otsu threshold needs 8 of them, the 3x3x3
    peak test 1, the +z affinity edge 1; the rest is there so that whole objects fit (step 5).
 3. The three quantities that couple the whole volume are all-reduced: input maximum
    (`vol /= max`, segmentation.py:889), per-channel affinity maxima (watershed.py:195; MAX),
    min / max (MIN / MAX) and the 256-bin histogram (SUM) of the smoothed mask channel
    -> the same Otsu threshold on every rank (watershed.py:227).
 4. Every rank runs the unchanged post stage (`isg_segment_features`) on its extended slab
    with those global scalars.  Mask components, the size window, the seeds inside a
    component and the ordered flood of a component depend on nothing outside the component
    (oracle-verified decomposition, iterseg_b200/csrc/flood.cuh), so every component that
    lies completely inside the extended slab is segmented exactly.
 5. Guard: a component that reaches the rank's own planes AND an open face of the extended
    slab is only partially known -> the kernel raises a flag and this module raises
    `HaloTooSmall` (re-run with a larger halo).  There is no silent approximation.
 6. Seam label merge: label ids are the rank of the seed in the global order
    (-smoothed centre value, C-order index) (peak_local_max + watershed.py:61).  The ranks
    all-gather the sort keys of the kept seeds in their own planes, sort them, and every rank
    rewrites its local ids (own planes + halo objects alike) to the global ones.

Not covered: a volume whose minimum is 0 -- the reference then strips all-zero slices before
normalising (segmentation.py:887-888), which changes the global chunk grid; `segment_volume_slabs`
all-reduces the minimum and raises NotImplementedError instead of silently differing.

`SlabWorker` holds one rank's state and exposes the phases; `segment_volume_slabs` wires them
to torch.distributed (NCCL), `segment_volume_emulated` runs R virtual ranks phase by phase on
one GPU (tests, single-GPU bench).
"""
import ctypes

import numpy as np
import torch

from . import _lib, predict, watershed

__all__ = ['plan_slabs', 'SlabWorker', 'segment_volume_slabs', 'segment_volume_emulated', 'HaloTooSmall']


class HaloTooSmall(RuntimeError):
    pass


class Slab:
    def __init__(self, rank, z0, z1, chunks, in0, in1):
        self.rank, self.z0, self.z1 = rank, int(z0), int(z1)
        self.chunks = chunks            # indices into the global chunk list
        self.in0, self.in1 = int(in0), int(in1)   # input planes the chunks read

    def __repr__(self):
        return f'Slab(rank={self.rank}, z=[{self.z0},{self.z1}), chunks={len(self.chunks)}, in=[{self.in0},{self.in1}))'


def plan_slabs(shape, chunk_size, margin, world):
    """Partition the z axis into `world` slabs whose borders coincide with the borders of the
    chunks' cropped interiors, balanced in planes.  Returns (slabs, (starts, crop_lo, crop_hi))."""
    st, lo, hi = predict._chunk_tables(tuple(shape), tuple(chunk_size), tuple(margin))
    a = st[:, 0] + lo[:, 0]
    b = st[:, 0] + hi[:, 0]
    layers = sorted(set(zip(a.tolist(), b.tolist())))
    for (a0, b0), (a1, b1) in zip(layers[:-1], layers[1:]):
        if b0 != a1:
            raise ValueError('chunk interiors do not tile the z axis')
    if world > len(layers):
        raise ValueError(f'{world} ranks but only {len(layers)} z-layers of chunks: use fewer ranks '
                         f'or a smaller chunk depth')
    Z = int(shape[0])
    # greedy split at the layer border closest to the ideal one
    cuts = [0]
    for r in range(1, world):
        ideal = r * Z / world
        cands = [L[0] for L in layers[1:] if L[0] > cuts[-1]]
        cands = cands[:len(cands) - (world - r - 1)] if world - r - 1 > 0 else cands
        cuts.append(min(cands, key=lambda c: abs(c - ideal)))
    cuts.append(Z)
    slabs = []
    for r in range(world):
        z0, z1 = cuts[r], cuts[r + 1]
        idx = np.nonzero((a >= z0) & (b <= z1))[0]
        in0 = int(st[idx, 0].min())
        in1 = int(st[idx, 0].max()) + int(chunk_size[0])
        slabs.append(Slab(r, z0, z1, idx, in0, in1))
    return slabs, (st, lo, hi)


def _u64_buf(n, device):
    return torch.zeros(max(int(n), 1), dtype=torch.int64, device=device)


class SlabWorker:
    """One rank of the slab-sharded segmentation (see module docstring for the phases)."""

    def __init__(self, net, shape, chunk_size, margin, rank, world, halo=24, device=None,
                 affinities_channels=(0, 1, 2), centroids_channel=4, thresholding_channel=3, scale=None):
        _lib.require_device()
        self.net, self.shape = net, tuple(int(s) for s in shape)
        self.chunk_size, self.margin = tuple(chunk_size), tuple(margin)
        self.rank, self.world = int(rank), int(world)
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.slabs, self.tables = plan_slabs(self.shape, chunk_size, margin, world)
        self.slab = self.slabs[self.rank]
        thinnest = min(s.z1 - s.z0 for s in self.slabs)
        self.halo = int(min(halo, thinnest)) if world > 1 else 0
        if world > 1 and self.halo < 8:
            raise ValueError(f'slabs of {thinnest} planes cannot provide the 8 halo planes the '
                             f'sigma=2 smoothing needs: use fewer ranks')
        self.h_lo = self.halo if self.rank > 0 else 0
        self.h_hi = self.halo if self.rank < self.world - 1 else 0
        self.channels = (tuple(affinities_channels), int(centroids_channel), int(thresholding_channel))
        self.scale = scale
        if int(np.prod(self.shape)) >= 2 ** 31:
            raise ValueError('volume too large for 31-bit global voxel ids')
        self.feats_own = self.feats_ext = None

    # ---- phase 1: input + U-Net ------------------------------------------------------------
    def local_input_max(self, volume):
        """Maximum of the own input planes (combine with MAX; segmentation.py:889)."""
        s = self.slab
        return float(np.max(volume[s.z0:s.z1]))

    def load_input(self, volume):
        """H2D of the input planes this rank's chunks read (one contiguous copy; asynchronous when
        `volume` is pinned).  Returns a device float32[2] = (min, max) of the OWN planes, to be
        combined with MIN / MAX over the ranks (segmentation.py:887-889)."""
        s = self.slab
        sub = volume[s.in0:s.in1]
        if isinstance(sub, torch.Tensor):
            host = sub.to(torch.float32).contiguous()
        else:
            host = torch.from_numpy(np.ascontiguousarray(sub, dtype=np.float32))
        self.frame = host.to(self.device, non_blocking=True)
        if self.frame.data_ptr() == host.data_ptr():
            self.frame = self.frame.clone()
        own = self.frame[s.z0 - s.in0:s.z1 - s.in0]
        return torch.stack([own.amin(), own.amax()])

    def unet_device(self, gmax):
        """Own chunks of the global chunk list -> feature planes [z0, z1) (5, nz, Y, X).
        gmax: 1-element float32 device tensor, the maximum of the WHOLE volume."""
        s = self.slab
        st, lo, hi = self.tables
        # vol /= max (segmentation.py:889): a true IEEE float32 division, tensor / tensor (torch
        # turns a division by a Python scalar into a multiplication by the reciprocal)
        frame = self.frame / gmax.reshape(()).to(self.device, torch.float32)
        self.frame = None
        st_l = st[s.chunks].copy()
        st_l[:, 0] -= s.in0
        tabs = (np.ascontiguousarray(st_l), np.ascontiguousarray(lo[s.chunks]),
                np.ascontiguousarray(hi[s.chunks]))
        out = torch.zeros((5,) + tuple(frame.shape), dtype=torch.float32, device=self.device)
        predict.predict_frame_device(self.net, frame, self.chunk_size, self.margin, out=out, tables=tabs)
        self.feats_own = out[:, s.z0 - s.in0:s.z1 - s.in0].contiguous()
        return self.feats_own

    def unet(self, volume, global_max):
        """load_input + unet_device with a host-side global maximum (emulated ranks, tests)."""
        self.load_input(volume)
        return self.unet_device(torch.tensor([np.float32(global_max)], dtype=torch.float32, device=self.device))

    def set_features(self, feats_own):
        """Bypass the U-Net (tests / feature maps from elsewhere): planes [z0, z1)."""
        assert tuple(feats_own.shape[1:]) == (self.slab.z1 - self.slab.z0,) + self.shape[1:]
        self.feats_own = feats_own.to(self.device, torch.float32).contiguous()

    # ---- phase 2: halo ---------------------------------------------------------------------
    def halo_for_lower(self):
        return self.feats_own[:, :self.halo].contiguous()

    def halo_for_upper(self):
        return self.feats_own[:, self.feats_own.shape[1] - self.halo:].contiguous()

    def set_halos(self, from_lower, from_upper):
        parts = []
        if self.h_lo:
            parts.append(from_lower)
        parts.append(self.feats_own)
        if self.h_hi:
            parts.append(from_upper)
        self.feats_ext = torch.cat(parts, dim=1).contiguous() if len(parts) > 1 else self.feats_own
        self.own0 = self.h_lo
        self.own1 = self.h_lo + (self.slab.z1 - self.slab.z0)
        self.ext_z0 = self.slab.z0 - self.h_lo

    # ---- phase 3: global scalars -----------------------------------------------------------
    def _params(self, absolute_thresh=None):
        aff, cent, thr = self.channels
        p, w1, w2 = watershed.post_params(aff, cent, thr, self.scale, absolute_thresh)
        p.own_z0, p.own_z1 = self.own0, self.own1
        p.open_faces = (1 if self.h_lo else 0) | (2 if self.h_hi else 0)
        return p, w1, w2

    def _stats(self, stage, minmax):
        lib = _lib.load()
        C, Z, Y, X = self.feats_ext.shape
        p, _, w2 = self._params()
        n = Z * Y * X
        ws = torch.empty(2 * (4 * n + 256) + 4096, dtype=torch.uint8, device=self.device)
        chan = torch.zeros(3, dtype=torch.float32, device=self.device)
        hist = torch.zeros(256, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.isg_slab_stats(self.feats_ext.data_ptr(), C, Z, Y, X, ctypes.byref(p), w2.ctypes.data,
                                    stage, minmax.data_ptr(), chan.data_ptr(), hist.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, 'isg_slab_stats')
        return chan, hist

    def stats0(self):
        """-> (affinity channel maxima [3], smoothed-mask min [1], max [1]) of the own planes."""
        mm = torch.zeros(2, dtype=torch.float32, device=self.device)
        chan, _ = self._stats(0, mm)
        return chan, mm[0:1].clone(), mm[1:2].clone()

    def stats1(self, gmin, gmax):
        """-> 256-bin histogram (int64) of the own planes over the GLOBAL [gmin, gmax]."""
        mm = torch.cat([gmin.reshape(1), gmax.reshape(1)]).to(self.device, torch.float32).contiguous()
        _, hist = self._stats(1, mm)
        return hist

    @staticmethod
    def otsu(hist, gmin, gmax):
        lib = _lib.load()
        dev = hist.device
        mm = torch.cat([gmin.reshape(1), gmax.reshape(1)]).to(dev, torch.float32).contiguous()
        thr = torch.zeros(1, dtype=torch.float32, device=dev)
        scratch = torch.empty(2048, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.isg_otsu_from_hist(hist.contiguous().data_ptr(), mm.data_ptr(), thr.data_ptr(),
                                        scratch.data_ptr(), scratch.numel(), _lib.stream_ptr())
        _lib.check(rc, 'isg_otsu_from_hist')
        return float(thr.item())

    # ---- phase 4/5: post stage on the extended slab ----------------------------------------
    def segment(self, thr, chan_max, max_seeds=None):
        """Runs the post stage; returns the global sort keys (int64) of the kept seeds that lie
        in the own planes.  Raises HaloTooSmall if an object that reaches the own planes is cut
        by an open face of the extended slab."""
        C, Z, Y, X = self.feats_ext.shape
        if max_seeds is None:
            max_seeds = max(1 << 16, (Z * Y * X) // 8)
        aff, cent, tch = self.channels
        self.labels_ext = torch.zeros((Z + 2, Y + 2, X + 2), dtype=torch.int32, device=self.device)
        keys = _u64_buf(max_seeds, self.device)
        slab = {'aff_div': [float(c) for c in chan_max.tolist()], 'own_z0': self.own0, 'own_z1': self.own1,
                'open_faces': (1 if self.h_lo else 0) | (2 if self.h_hi else 0), 'seed_keys': keys}
        # the ordered flood's compact arenas are sized for a quarter of the voxels lying in
        # multi-seed components (a real mask covers a few per cent); if a slab needs more, the call
        # fails loudly and is repeated with the worst-case workspace
        kw = dict(scale=self.scale, absolute_thresh=thr, max_seeds=max_seeds, slab=slab)
        try:
            seeds, counts, mask, _ = watershed.segment_features_device(
                self.feats_ext, self.labels_ext, aff, cent, tch, max_flood_nodes=(Z * Y * X) // 4, **kw)
        except _lib.IsgError as e:
            if e.status != _lib.ISG_ERR_WORKSPACE:
                raise
            self.labels_ext.zero_()
            seeds, counts, mask, _ = watershed.segment_features_device(
                self.feats_ext, self.labels_ext, aff, cent, tch, **kw)
        c = counts.cpu().numpy()
        # an object that reaches the own planes is cut by an open face of the extended slab: the caller
        # combines this flag over the ranks (MAX) and calls `check_halo` -- all ranks raise together,
        # none is left waiting in a collective
        self.halo_violation = int(c[4])
        self.n_local = int(c[0])
        k = keys[:self.n_local]
        # local key = (~ord(value) << 32) | local unpadded flat index  ->  global flat index
        v = k & 0xFFFFFFFF
        zl = torch.div(v, Y * X, rounding_mode='floor')
        self.keys_global = (k - v) + (v + self.ext_z0 * (Y * X))          # z shift only
        own = (zl >= self.own0) & (zl < self.own1)
        self.mask_ext = mask
        return self.keys_global[own].contiguous()

    def check_halo(self, any_violation):
        if any_violation:
            raise HaloTooSmall(f'an object that reaches the own planes of a rank is cut by the halo of {self.halo} '
                               f'planes (rank {self.rank}: {"yes" if self.halo_violation else "no"}); re-run with a '
                               f'larger halo')

    # ---- phase 6: global label ids -----------------------------------------------------------
    def relabel(self, global_sorted_keys):
        """Local label ids -> global ones; returns the own planes (nz, Y, X) as int32 (uint32 bits)."""
        lib = _lib.load()
        C, Z, Y, X = self.feats_ext.shape
        own = self.labels_ext[1 + self.own0:1 + self.own1, 1:-1, 1:-1].contiguous()
        lut = torch.zeros(max(self.n_local, 1), dtype=torch.int32, device=self.device)
        missing = torch.zeros(1, dtype=torch.int32, device=self.device)
        g = global_sorted_keys.to(self.device).contiguous()
        with torch.cuda.device(self.device):
            rc = lib.isg_relabel_by_keys(own.data_ptr(), own.numel(), self.keys_global.contiguous().data_ptr(),
                                         self.n_local, g.data_ptr(), g.numel(), lut.data_ptr(),
                                         missing.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, 'isg_relabel_by_keys')
        if int(missing.item()):
            raise RuntimeError(f'rank {self.rank}: a label of the own planes has no global id '
                               f'(inconsistent seed exchange)')
        return own


def sort_keys(keys):
    """Ascending sort of int64 keys (all < 2^63) on the device with the library's radix sort."""
    lib = _lib.load()
    keys = keys.contiguous().clone()
    n = keys.numel()
    if n > 1:
        nb = lib.isg_sort_tmp_bytes(n)
        tmp = torch.empty(nb, dtype=torch.uint8, device=keys.device)
        with torch.cuda.device(keys.device):
            rc = lib.isg_sort_keys_u64(keys.data_ptr(), n, tmp.data_ptr(), nb, _lib.stream_ptr())
        _lib.check(rc, 'isg_sort_keys_u64')
    return keys


def segment_volume_emulated(volume, net, chunk_size, margin, world, halo=24, features=None, **kw):
    """R virtual ranks, phase by phase, on the current GPU.  Returns the (Z,Y,X) uint32 labels
    (numpy) assembled from the ranks' own planes and the number of labels.  `features`
    (5,Z,Y,X) bypasses the U-Net (numpy or tensor)."""
    shape = tuple(volume.shape) if volume is not None else tuple(features.shape[1:])
    ws = [SlabWorker(net, shape, chunk_size, margin, r, world, halo=halo, **kw) for r in range(world)]
    if features is None:
        if float(np.min(volume)) == 0.0:
            raise NotImplementedError('the volume contains zeros (remove_sum_zero_slices is not sharded)')
        gmax = max(w.local_input_max(volume) for w in ws)
        for w in ws:
            w.unet(volume, gmax)
    else:
        f = torch.as_tensor(features)
        for w in ws:
            w.set_features(f[:, w.slab.z0:w.slab.z1])
    for r, w in enumerate(ws):
        w.set_halos(ws[r - 1].halo_for_upper() if r > 0 else None,
                    ws[r + 1].halo_for_lower() if r < world - 1 else None)
    s0 = [w.stats0() for w in ws]
    chan = torch.stack([s[0] for s in s0]).amax(0)
    gmin = torch.stack([s[1] for s in s0]).amin(0)
    gmax_s = torch.stack([s[2] for s in s0]).amax(0)
    hist = torch.stack([w.stats1(gmin, gmax_s) for w in ws]).sum(0)
    thr = SlabWorker.otsu(hist, gmin, gmax_s)
    keys = torch.cat([w.segment(thr, chan) for w in ws])
    bad = any(w.halo_violation for w in ws)
    for w in ws:
        w.check_halo(bad)
    gsorted = sort_keys(keys)
    out = np.zeros(shape, dtype=np.uint32)
    for w in ws:
        out[w.slab.z0:w.slab.z1] = w.relabel(gsorted).cpu().numpy().view(np.uint32)
    return out, int(gsorted.numel()), {'otsu': thr, 'aff_max': chan.tolist(), 'slabs': [repr(w.slab) for w in ws]}


def segment_volume_slabs(volume, net, chunk_size, margin, halo=24, group=None, features=None, **kw):
    """Every rank calls this with the same (host, e.g. memory-mapped) volume; returns
    (own labels int32 CUDA tensor (nz,Y,X), (z0, z1), number of labels in the whole volume)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shape = tuple(volume.shape) if volume is not None else tuple(features.shape[1:])
    w = SlabWorker(net, shape, chunk_size, margin, rank, world, halo=halo, **kw)
    dev = w.device
    if features is None:
        mm = w.load_input(volume)                       # H2D + device min / max of the own planes
        lo, hi = mm[0:1].clone(), mm[1:2].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        if float(lo.item()) == 0.0:
            # segment_single_volume strips the slices that sum to zero when min() == 0
            # (segmentation.py:887-888): that changes the chunk grid of the WHOLE volume
            raise NotImplementedError(
                'the volume contains zeros: the reference strips all-zero slices first '
                '(remove_sum_zero_slices), which the slab-sharded path does not implement -- strip '
                'them before sharding or run the volume on one device')
        w.unet_device(hi)
    else:
        w.set_features(torch.as_tensor(features)[:, w.slab.z0:w.slab.z1])
    # halo planes to / from the neighbours (NCCL point-to-point over NVLink)
    Y, X = shape[1:]
    lo = torch.empty((5, w.halo, Y, X), dtype=torch.float32, device=dev) if w.h_lo else None
    hi = torch.empty((5, w.halo, Y, X), dtype=torch.float32, device=dev) if w.h_hi else None
    ops, keep = [], []
    if w.h_lo:
        t = w.halo_for_lower(); keep.append(t)
        ops += [dist.P2POp(dist.isend, t, rank - 1, group), dist.P2POp(dist.irecv, lo, rank - 1, group)]
    if w.h_hi:
        t = w.halo_for_upper(); keep.append(t)
        ops += [dist.P2POp(dist.isend, t, rank + 1, group), dist.P2POp(dist.irecv, hi, rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    w.set_halos(lo, hi)
    chan, mn, mx = w.stats0()
    dist.all_reduce(chan, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    hist = w.stats1(mn, mx)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    thr = SlabWorker.otsu(hist, mn, mx)
    own_keys = w.segment(thr, chan)
    bad = torch.tensor([w.halo_violation], dtype=torch.int32, device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
    w.check_halo(int(bad.item()))
    # seam label merge: all-gather the (padded) key lists, sort, relabel
    cnt = torch.tensor([own_keys.numel()], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    cmax = max(int(c.item()) for c in cnts)
    pad = torch.full((max(cmax, 1),), -1, dtype=torch.int64, device=dev)
    pad[:own_keys.numel()] = own_keys
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    allk = torch.cat([b[:int(c.item())] for b, c in zip(bufs, cnts)])
    gsorted = sort_keys(allk)
    own = w.relabel(gsorted)
    return own, (w.slab.z0, w.slab.z1), int(gsorted.numel())
