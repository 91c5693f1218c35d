"""Seeded synthetic platelet volumes and analytic feature maps.

There is no network for datasets and the reference's bundled network file is
missing, so benchmarks and parity tests run on synthetic data:

* `platelet_labels`  -- oblate ellipsoids ("platelets"), rejection-placed so
  that each keeps >= 80 % of its volume (objects may touch and form small
  aggregates).
* `platelet_frame`   -- a fluorescence-like zyx float32 frame rendered from the
  labels (PSF blur, noise, strictly positive so that the reference's
  remove_sum_zero_slices branch, segmentation.py:887-888, is never taken).
* `analytic_features` -- the 5 channels the U-Net would predict, derived from
  the ground truth with the reference's target conventions: affinities
  aff[a][v] = 1 where labels differ between v-1 and v along axis a
  (labels.py:87-109), mask = labels > 0, centreness peaking at each centroid
  (labels.py:143-205).

Pure numpy/scipy host code: data generation is not on the measured path.
"""
import numpy as np
from scipy import ndimage as ndi


def platelet_labels(shape=(33, 512, 512), n_objects=None, seed=0, max_tries=None):
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    if n_objects is None:
        n_objects = int(round(1500 * (Z * Y * X) / (33 * 512 * 512)))
    labels = np.zeros(shape, dtype=np.int32)
    placed, tries = 0, 0
    max_tries = max_tries or 20 * n_objects + 100
    while placed < n_objects and tries < max_tries:
        tries += 1
        rz = rng.uniform(1.0, 2.0)
        ra, rb = rng.uniform(4.0, 8.0, 2)
        th = rng.uniform(0, np.pi)
        cz, cy, cx = rng.uniform(0, Z), rng.uniform(0, Y), rng.uniform(0, X)
        R = int(np.ceil(max(ra, rb))) + 1
        Rz = int(np.ceil(rz)) + 1
        z0, z1 = max(0, int(cz) - Rz), min(Z, int(cz) + Rz + 1)
        y0, y1 = max(0, int(cy) - R), min(Y, int(cy) + R + 1)
        x0, x1 = max(0, int(cx) - R), min(X, int(cx) + R + 1)
        if z0 >= z1 or y0 >= y1 or x0 >= x1:
            continue
        zz, yy, xx = np.meshgrid(np.arange(z0, z1) - cz, np.arange(y0, y1) - cy,
                                 np.arange(x0, x1) - cx, indexing='ij')
        u = np.cos(th) * yy + np.sin(th) * xx
        v = -np.sin(th) * yy + np.cos(th) * xx
        inside = (zz / rz) ** 2 + (u / ra) ** 2 + (v / rb) ** 2 <= 1.0
        vol = int(inside.sum())
        if vol < 20:
            continue
        sub = labels[z0:z1, y0:y1, x0:x1]
        free = inside & (sub == 0)
        if free.sum() < 0.8 * vol:
            continue
        placed += 1
        sub[free] = placed
    return labels


def platelet_frame(shape=(33, 512, 512), seed=0, n_objects=None, return_labels=False):
    labels = platelet_labels(shape, n_objects, seed)
    rng = np.random.default_rng(seed + 7919)
    n = int(labels.max())
    inten = np.concatenate([[0.0], rng.uniform(0.4, 1.0, n)]).astype(np.float32)
    vol = inten[labels]
    vol = ndi.gaussian_filter(vol, (0.7, 1.5, 1.5), mode='nearest')
    vol = vol + rng.normal(0.0, 0.03, shape).astype(np.float32) + np.float32(0.02)
    vol = np.clip(vol, 1e-3, None).astype(np.float32)
    vol /= vol.max()
    return (vol, labels) if return_labels else vol


def timeseries(n_frames, shape=(33, 512, 512), seed0=0):
    for t in range(n_frames):
        yield platelet_frame(shape, seed0 + t)


def analytic_features(labels, seed=0, noise=0.01):
    """(5,Z,Y,X) float32: z/y/x affinities, mask, centreness."""
    rng = np.random.default_rng(seed + 104729)
    shape = labels.shape
    feats = np.zeros((5,) + shape, dtype=np.float32)
    for a in range(3):
        pad = [(0, 0)] * 3
        pad[a] = (1, 0)
        lp = np.pad(labels, pad, mode='reflect')
        sl0 = [slice(None)] * 3
        sl1 = [slice(None)] * 3
        sl0[a] = slice(0, shape[a])
        sl1[a] = slice(1, shape[a] + 1)
        aff = (lp[tuple(sl0)] != lp[tuple(sl1)]).astype(np.float32)
        feats[a] = ndi.gaussian_filter(aff, 0.5, mode='nearest')
    feats[3] = ndi.gaussian_filter((labels > 0).astype(np.float32), 0.7, mode='nearest')
    n = int(labels.max())
    if n > 0:
        idx = np.arange(1, n + 1)
        com = np.array(ndi.center_of_mass(np.ones(shape, np.float32), labels, idx))
        zz, yy, xx = np.nonzero(labels)
        lab = labels[zz, yy, xx]
        c = com[lab - 1]
        d = np.sqrt((4.0 * (zz - c[:, 0])) ** 2 + (yy - c[:, 1]) ** 2 + (xx - c[:, 2]) ** 2)
        feats[4][zz, yy, xx] = (1.0 / (1.0 + 0.35 * d)).astype(np.float32)
    if noise:
        feats += rng.normal(0.0, noise, feats.shape).astype(np.float32)
    np.clip(feats, 1e-4, 1.0, out=feats)
    return feats


def structured_state_dict(seed=0, noise=0.25):
    """A deterministic network file (the reference's state_dict format, train.py:414-420)
    whose outputs are platelet-shaped WITHOUT training: the bundled network is missing
    from the reference checkout and there is no network to fetch or time to train one.

    Every convolution gets dense random weights (torch's default scale x `noise`) plus a
    hand-placed "carrier": channel 0 of every layer copies channel 0 of its input (the first
    layer applies an in-plane 3x3 box blur; the decoder averages the upsampled and the skip
    carrier), so a smoothed copy of the
    intensity image survives the encoder/decoder.  The head maps it to
      affinities (ch 0-2) = -blur (valleys between touching objects are expensive),
      mask (ch 3)         = +fine blur,      centre (ch 4) = +coarse blur.
    FLOPs, shapes and dtypes are those of any UNet(1,5) file; only the values differ.
    """
    import torch
    rng = np.random.default_rng(seed)
    sd = {}
    enc = [('c0', 1, 32), ('c1', 32, 64), ('c2', 64, 128), ('c3', 128, 256), ('c4', 256, 256)]
    dec = [('c5_0', 512, 128), ('c6_0', 256, 64), ('c7_0', 128, 32), ('c8_0', 64, 5)]
    ups = [('up0', 256, (2, 2, 2)), ('up1', 128, (1, 2, 2)), ('up2', 64, (1, 2, 2)), ('up3', 32, (1, 2, 2))]

    def conv(prefix, cin, cout, carriers, blur=False):
        b = noise / np.sqrt(cin * 27)
        w = rng.uniform(-b, b, (cout, cin, 3, 3, 3)).astype(np.float32)
        for co, ci, gain in carriers:
            if blur:
                w[co, ci, 1] += np.float32(gain / 9.0)      # in-plane 3x3 box blur
            else:
                w[co, ci, 1, 1, 1] += np.float32(gain)      # identity
        sd[prefix + '.weight'] = torch.from_numpy(w)
        sd[prefix + '.bias'] = torch.zeros(cout)

    def bn(prefix, c, beta0=1.5):
        g = rng.uniform(0.8, 1.2, c).astype(np.float32)
        be = rng.uniform(-0.2, 0.2, c).astype(np.float32)
        g[0], be[0] = 1.0, beta0
        sd[prefix + '.weight'] = torch.from_numpy(g)
        sd[prefix + '.bias'] = torch.from_numpy(be)
        sd[prefix + '.running_mean'] = torch.zeros(c)
        sd[prefix + '.running_var'] = torch.ones(c)
        sd[prefix + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    for name, cin, cout in enc:
        conv(name + '.conv0', cin, cout, [(0, 0, 1.0)], blur=(name == 'c0'))
        conv(name + '.conv1', cout, cout, [(0, 0, 1.0)])
        bn(name + '.batch0', cout)
        bn(name + '.batch1', cout)
    for name, cin, cout in dec:
        half = cin // 2                      # concat = [upsampled (0..half), skip (half..)]
        if name != 'c8_0':
            conv(name + '.conv0', cin, cout, [(0, 0, 0.5), (0, half, 0.5)])
            conv(name + '.conv1', cout, cout, [(0, 0, 1.0)])
            bn(name + '.batch0', cout)
            bn(name + '.batch1', cout)
        else:
            # head: ch0-2 = -fine, ch3 = +fine, ch4 = +coarse
            conv(name + '.conv0', cin, cout,
                 [(0, half, -1.0), (1, half, -1.0), (2, half, -1.0), (3, half, 1.0), (4, 0, 1.0)])
            conv(name + '.conv1', cout, cout, [(k, k, 1.0) for k in range(5)])
            for i in (0, 1):
                sd[f'{name}.batch{i}.weight'] = torch.ones(cout)
                sd[f'{name}.batch{i}.bias'] = torch.full((cout,), 1.0 if i == 0 else 0.0)
                sd[f'{name}.batch{i}.running_mean'] = torch.zeros(cout)
                sd[f'{name}.batch{i}.running_var'] = torch.ones(cout)
                sd[f'{name}.batch{i}.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    # fix the key order of c8_0 to conv0, conv1, batch0, batch1 (state_dict order)
    ordered = {}
    for name, _, _ in enc + dec:
        for k in ('conv0.weight', 'conv0.bias', 'conv1.weight', 'conv1.bias'):
            ordered[f'{name}.{k}'] = sd[f'{name}.{k}']
        for i in (0, 1):
            for k in ('weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked'):
                ordered[f'{name}.batch{i}.{k}'] = sd[f'{name}.batch{i}.{k}']
    for name, c, k in ups:
        w = rng.uniform(-1, 1, (c, 1) + k).astype(np.float32) * np.float32(0.5 * noise)
        w[0] = 1.0                            # carrier: nearest-neighbour upsampling
        bb = rng.uniform(-0.1, 0.1, c).astype(np.float32) * np.float32(noise)
        bb[0] = 0.0
        ordered[name + '.weight'] = torch.from_numpy(w)
        ordered[name + '.bias'] = torch.from_numpy(bb)
    return ordered


class JitteredSeries:
    """A tzyx float32 series of `n_frames` synthetic platelet frames that costs `n_base` frame
    renderings instead of n_frames (BASELINE.json configs[2]: 192 frames; rendering one frame
    takes ~1 s of host time): frame t is base frame t % n_base rolled by (0, 3k, 5k) voxels with
    k = t // n_base -- the platelets drift frame to frame, every frame is distinct, min > 0.
    Only the frames listed in `own` are materialised (a rank of a frame-sharded run holds its own
    frames), in pinned host memory when `pin` is set.  Duck-types what `segmentation_loop`
    reads: .shape, .ndim, .dtype, data[t]."""

    def __init__(self, n_frames, shape=(33, 512, 512), n_base=4, seed0=0, own=None, pin=False):
        self.shape = (int(n_frames),) + tuple(int(s) for s in shape)
        self.ndim = 4
        self.dtype = np.dtype(np.float32)
        self.n_base = int(n_base)
        own = range(n_frames) if own is None else own
        base = {}
        self.frames = {}
        buf = None
        own = list(own)
        if pin:
            import torch
            self._pinned = torch.empty((len(own),) + self.shape[1:], dtype=torch.float32).pin_memory()
            buf = self._pinned.numpy()
        for i, t in enumerate(own):
            b = t % self.n_base
            if b not in base:
                base[b] = platelet_frame(self.shape[1:], seed=seed0 + b)
            k = t // self.n_base
            fr = np.roll(base[b], (3 * k, 5 * k), axis=(1, 2)) if k else base[b]
            if buf is not None:
                buf[i] = fr
                self.frames[t] = buf[i]
            else:
                self.frames[t] = np.ascontiguousarray(fr)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, t):
        if isinstance(t, tuple):
            return self.frames[int(t[0])][t[1:]]
        return self.frames[int(t)]


def big_volume_planes(shape, z0, z1, tile=(32, 512, 512), seed0=1000, n_base=3, out=None):
    """Planes [z0, z1) of the synthetic big volume of BASELINE.json configs[3] (256 x 2048 x 2048,
    same platelet density): a mosaic of `n_base` rendered (tile z + 1, tile y, tile x) blocks,
    each mosaic cell a different block / in-plane roll, the z tiling shifted by half a tile so
    that objects straddle the z-slab borders of an 8-way split.  Deterministic in (shape, seed0):
    every rank renders the planes it needs and they fit together."""
    Z, Y, X = (int(s) for s in shape)
    tz, ty, tx = tile
    base = [platelet_frame((tz, ty, tx), seed=seed0 + b) for b in range(n_base)]
    planes = np.empty((z1 - z0, Y, X), np.float32) if out is None else out
    for z in range(z0, z1):
        zz = z + tz // 2
        cz, lz = zz // tz, zz % tz
        for iy in range(0, Y, ty):
            for ix in range(0, X, tx):
                cell = cz * 131 + (iy // ty) * 17 + (ix // tx) * 5
                src = base[cell % n_base][lz]
                sh = (7 * cell) % ty, (11 * cell) % tx
                blk = np.roll(src, sh, axis=(0, 1))
                planes[z - z0, iy:iy + ty, ix:ix + tx] = blk[:min(ty, Y - iy), :min(tx, X - ix)]
    return planes
