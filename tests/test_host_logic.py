"""Host-side logic of the package (no GPU): chunk grids, config handling, label store,
frame sharding with a 2-rank gloo group."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_make_chunks_matches_reference_golden(golden_dir):
    from iterseg_b200 import predict
    grids = json.load(open(os.path.join(golden_dir, 'chunk_grids.json')))
    for key, g in grids.items():
        shape, chunk, margin = json.loads(key)
        st, cr = predict.make_chunks(tuple(shape), tuple(chunk), margin if isinstance(margin, int) else tuple(margin))
        assert [list(map(int, s)) for s in st] == g['starts'], key
        assert [[list(map(int, ab)) for ab in c] for c in cr] == g['crops'], key


def test_make_chunks_rejects_small_arrays():
    from iterseg_b200 import predict
    with pytest.raises(ValueError):
        predict.make_chunks((8, 256, 256), (10, 256, 256), (1, 64, 64))


def test_process_chunks_generic_function_provenance(golden_dir):
    from iterseg_b200 import predict
    prov = np.load(os.path.join(golden_dir, 'provenance_12x300x300.npz'))['provenance']
    out = np.zeros((1, 12, 300, 300), np.float32)
    k = {'i': 0}

    def fake(input_volume, sl, **kw):
        k['i'] += 1
        return np.full((1, 1) + input_volume[sl[1:]].shape, k['i'], np.float32)

    predict.process_chunks(np.zeros((12, 300, 300), np.float32), (10, 256, 256), out, (1, 64, 64), fake)
    assert np.array_equal(out[0].astype(np.uint8), prov)


def test_label_store_roundtrip_and_zarr_chunk_rule(tmp_path):
    from iterseg_b200 import _io
    meta = {'scale': (1, 4, 1, 1), 'translate': (0, 0, 0, 0), 'name': 'n'}
    a = _io.save_labels_to_ome(tmp_path / 'x.ome.zarr', layer_meta=meta, shape=(3, 12, 40, 50),
                               chunks=(2, 5, 16), dtype=np.int32)
    assert a.chunks == (2, 5, 16, 50)        # zarr v2: missing trailing dims span the axis
    ref = np.random.default_rng(0).integers(0, 100, (3, 12, 40, 50)).astype(np.int32)
    assert not np.any(a[1])
    a[1, ...] = ref[1]
    a[0] = ref[0]
    a[2, 3:5] = ref[2, 3:5]
    assert np.array_equal(a[1], ref[1]) and np.array_equal(a[0], ref[0])
    assert np.array_equal(a[2, 3:5], ref[2, 3:5]) and not np.any(a[2, 0])
    # a second handle on the same directory sees the data (it is a real store)
    b = _io.open_zarr(str(tmp_path / 'x.ome.zarr' / '0'), shape=(3, 12, 40, 50), chunks=(2, 5, 16), dtype=np.int32)
    assert np.array_equal(np.asarray(b)[:2], ref[:2])
    attrs = json.load(open(tmp_path / 'x.ome.zarr' / '.zattrs'))
    assert attrs['multiscales'][0]['datasets'][0]['coordinateTransformations'][1]['type'] == 'translate'
    with pytest.raises(ValueError):
        _io.save_labels_to_ome(tmp_path / 'y.ome.zarr', layer_meta=meta)


def test_prep_config_errors(tmp_path):
    from iterseg_b200 import segmentation, viewer
    layer = viewer.Image(np.zeros((10, 256, 256), np.float32))
    with pytest.raises(ValueError):
        segmentation.affinity_watershed_prep_config(layer, 'network.txt', None)
    with pytest.raises(AssertionError):
        segmentation.affinity_watershed_prep_config(layer, str(tmp_path / 'missing.pt'), None)
    assert segmentation.a_w_output_volume(np.zeros((4, 10, 20, 30)), 5).shape == (5, 10, 20, 30)
    assert set(segmentation.segmenters) == {'affinity-unet-watershed', 'DoG-blob-watershed'}


def test_remove_sum_zero_slices():
    from iterseg_b200 import segmentation
    x = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    x[:, 1, :] = 0
    x[0] = 0
    assert segmentation.remove_sum_zero_slices(x).shape == (1, 2, 4)


def test_state_dict_keys_match_reference_format():
    from iterseg_b200 import synth, unet
    from oracle import unet_ref
    net = unet.UNet()
    assert list(net.state_dict()) == list(unet_ref.synth_state_dict(0))
    net.load_state_dict(synth.structured_state_dict(0))
    with pytest.raises(NotImplementedError):
        unet.UNet(in_channels=2)


def test_no_device_fails_loudly():
    import torch
    from iterseg_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(_lib.IsgError):
        _lib.require_device()
    from iterseg_b200 import watershed
    with pytest.raises(_lib.IsgError):
        watershed.segment_output_image(np.zeros((5, 4, 8, 8), np.float32), (0, 1, 2), 4, 3)


def test_frame_sharding_and_offsets():
    from iterseg_b200 import distributed as d
    assert d.shard_frames(7, 1, 3) == [1, 4]
    assert sorted(sum((d.shard_frames(192, r, 8) for r in range(8)), [])) == list(range(192))
    assert d.exclusive_offsets([3, 0, 5, 2]).tolist() == [0, 3, 3, 8]
    lab = np.array([0, 1, 2, 0], np.int32)
    assert d.add_label_offset_host(lab, 10).tolist() == [0, 11, 12, 0]


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from iterseg_b200 import distributed as d
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
T = 7
mine = d.shard_frames(T, rank, world)
local = {t: 10 * t + 1 for t in mine}            # pretend frame t holds 10t+1 labels
counts = d.gather_label_counts(local, T, rank, world)
assert counts.tolist() == [10 * t + 1 for t in range(T)], counts
off = d.exclusive_offsets(counts)
# every rank offsets its own frames; together the ids are globally unique and dense
out = {}
for t in mine:
    lab = np.arange(0, counts[t] + 1).astype(np.int64)
    out[t] = d.add_label_offset_host(lab, int(off[t]))
    assert out[t][0] == 0 and out[t][1] == off[t] + 1 and out[t][-1] == off[t] + counts[t]
dist.barrier()
dist.destroy_process_group()
sys.stdout.write('rank ' + str(rank) + ' ok\n')
sys.stdout.flush()
'''


def test_label_count_allgather_two_ranks_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(GLOO_WORKER % {'root': ROOT})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29611', str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'rank 0 ok' in r.stdout and 'rank 1 ok' in r.stdout


def test_plan_slabs_partitions_chunks_exactly_once():
    """Slab borders coincide with chunk-interior borders; every chunk of the GLOBAL list belongs to
    exactly one rank (its BatchNorm batch is never split or recomputed)."""
    from iterseg_b200 import slab
    for shape, world in (((256, 2048, 2048), 8), ((33, 512, 512), 2), ((96, 160, 160), 4)):
        slabs, (st, lo, hi) = slab.plan_slabs(shape, (10, 256, 256) if shape[1] >= 256 else (10, 64, 64),
                                              (1, 64, 64) if shape[1] >= 256 else (1, 16, 16), world)
        assert slabs[0].z0 == 0 and slabs[-1].z1 == shape[0]
        seen = np.zeros(len(st), int)
        for a, b in zip(slabs[:-1], slabs[1:]):
            assert a.z1 == b.z0
        for s in slabs:
            seen[s.chunks] += 1
            za, zb = st[s.chunks, 0] + lo[s.chunks, 0], st[s.chunks, 0] + hi[s.chunks, 0]
            assert za.min() == s.z0 and zb.max() == s.z1
            assert s.in0 <= s.z0 and s.in1 >= s.z1
        assert (seen == 1).all()
    slabs, _ = slab.plan_slabs((256, 2048, 2048), (10, 256, 256), (1, 64, 64), 8)
    sizes = [s.z1 - s.z0 for s in slabs]
    assert max(sizes) - min(sizes) <= 8 and len(slabs[0].chunks) * 8 >= 7200 * 0.9
    with pytest.raises(ValueError):
        slab.plan_slabs((33, 512, 512), (10, 256, 256), (1, 64, 64), 8)


OFFSETS_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from iterseg_b200 import _io, distributed as d
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
assert d.world() == (rank, world)
T, shape = 7, (3, 8, 6)
mine = d.shard_frames(T, rank, world)
n_steps = (T + world - 1) // world
counts = {t: 10 * t + 1 for t in range(T)}          # pretend frame t holds 10t+1 labels
off = d.LabelOffsets(rank, world, torch.device('cpu'))
store = _io.open_zarr(%(store)r, shape=(T,) + shape, chunks=(4, 2, 8), dtype=np.int32)   # t-chunk of 4 frames
got = {}
for s in range(n_steps):
    t = s * world + rank
    o = off.step(torch.tensor([counts[t]]) if t < T else None)     # every rank, every step
    if t < T:
        got[t] = int(o.item())
        store[t, ...] = np.full(shape, got[t] + 1, np.int32)       # ranks share t-chunk files
want = np.concatenate([[0], np.cumsum([counts[t] for t in range(T)])])
assert all(got[t] == want[t] for t in mine), (got, want)
assert int(off.total.item()) == want[-1]
dist.barrier()
if rank == 0:
    a = np.asarray(_io.open_zarr(%(store)r))
    for t in range(T):
        assert (a[t] == want[t] + 1).all(), (t, a[t].ravel()[:4], want[t] + 1)
dist.barrier()
dist.destroy_process_group()
sys.stdout.write('rank ' + str(rank) + ' ok\n')
sys.stdout.flush()
'''


def test_running_label_offsets_and_shared_store_two_ranks_gloo(tmp_path):
    """World-size-2 run of the host logic behind the frame-sharded series (BASELINE configs[2]):
    frames t = rank (mod 2), one all-gather per step giving the exclusive prefix of the label
    counts, and both ranks writing their frames into the SAME t-chunk files of one zarr store."""
    script = tmp_path / 'worker2.py'
    script.write_text(OFFSETS_WORKER % {'root': ROOT, 'store': str(tmp_path / 'lab')})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29613', str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'rank 0 ok' in r.stdout and 'rank 1 ok' in r.stdout


def test_label_store_reopens_with_its_own_metadata(tmp_path):
    """an existing store keeps its shape / chunks / dtype on re-open (warm restart with a
    different chunk_size), a shape mismatch raises, and a compressed store is refused."""
    from iterseg_b200 import _io
    a = _io.open_zarr(str(tmp_path / 's'), shape=(6, 4, 10, 12), chunks=(2, 4, 5), dtype=np.int32)
    v = np.arange(4 * 10 * 12, dtype=np.int32).reshape(4, 10, 12)
    a[3, ...] = v
    b = _io.open_zarr(str(tmp_path / 's'), shape=(6, 4, 10, 12), chunks=(1, 2, 2), dtype=np.int64)
    assert b.chunks == (2, 4, 5, 12) and b.dtype == np.int32
    assert np.array_equal(b[3], v) and not b[2].any()
    with pytest.raises(ValueError):
        _io.open_zarr(str(tmp_path / 's'), shape=(5, 4, 10, 12), chunks=(2, 4, 5))
    meta_fn = tmp_path / 's' / '.zarray'
    meta = json.loads(meta_fn.read_text())
    meta['compressor'] = {'id': 'blosc', 'cname': 'lz4', 'clevel': 5, 'shuffle': 1, 'blocksize': 0}
    meta_fn.write_text(json.dumps(meta))
    with pytest.raises(NotImplementedError):
        _io.open_zarr(str(tmp_path / 's'), shape=(6, 4, 10, 12), chunks=(2, 4, 5))


def test_series_and_volume_synth_are_rank_consistent():
    from iterseg_b200 import synth
    a = synth.JitteredSeries(12, (6, 40, 40), n_base=2, own=[1, 5, 9])
    b = synth.JitteredSeries(12, (6, 40, 40), n_base=2, own=range(12))
    assert a.shape == (12, 6, 40, 40) and a.ndim == 4
    for t in (1, 5, 9):
        assert np.array_equal(a[t], b[t]) and a[t].min() > 0
    assert not np.array_equal(b[1], b[3])                 # same base frame, different jitter
    p = synth.big_volume_planes((40, 96, 96), 10, 14, tile=(8, 32, 32))
    q = synth.big_volume_planes((40, 96, 96), 0, 40, tile=(8, 32, 32))
    assert np.array_equal(p, q[10:14]) and q.min() > 0


def test_label_store_partial_writes_in_place(tmp_path):
    """Partial chunk updates (a frame of a multi-frame t-chunk, sub-blocks, single rows) are written in
    place -- positional writes for contiguous runs, a shared mapping otherwise -- for every chunk shape,
    and read back (region reads) exactly."""
    from iterseg_b200 import _io
    rng = np.random.default_rng(0)
    for k, chunks in enumerate([(10, 4, 16), (4, 2, 7, 5), (3, 5, 40, 30), (1, 1, 1, 1), (25, 5, 40)]):
        a = _io.open_zarr(str(tmp_path / f'a{k}'), shape=(25, 5, 40, 30), chunks=chunks, dtype=np.int32)
        ref = np.zeros((25, 5, 40, 30), np.int32)
        for t in (3, 11, 24, 0, 19):
            v = rng.integers(1, 100, (5, 40, 30)).astype(np.int32)
            a[t, ...] = v
            ref[t] = v
        a[5:7, 1:3, 7:29, 3:17] = 7
        ref[5:7, 1:3, 7:29, 3:17] = 7
        a[8, 2, 5] = np.arange(30)
        ref[8, 2, 5] = np.arange(30)
        assert np.array_equal(np.asarray(a), ref), chunks
        assert np.array_equal(a[11], ref[11]) and np.array_equal(a[5:7, 1:3], ref[5:7, 1:3])
        b = _io.open_zarr(str(tmp_path / f'a{k}'))                    # re-open: own metadata
        assert np.array_equal(np.asarray(b), ref)


def test_band_flat_conv_tiling_covers_every_voxel_once():
    """The band-flat tiling of conv3d_tc (host-side planner, no device needed): a tile is
    128 consecutive positions f = y*P + x' of a column band.  For many plane sizes: every output
    voxel lies in exactly one (band, tile, row), and every tap of every row stays inside the R*P rows
    of the plane slot whose TMA box starts at padded row y0 = 128*fb // P."""
    import ctypes
    from iterseg_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int64 * 4)()
    sizes = [(256, 256), (129, 129), (65, 65), (33, 33), (17, 17), (9, 9), (5, 5), (3, 3), (96, 160), (33, 40), (2, 7)]
    for H, W in sizes:
        for rb, cap in ((64, 32 << 10), (128, 32 << 10), (128, 24 << 10), (64, 20 << 10)):
            rc = lib.isg_debug_flat_tiling(H, W, rb, cap, out)
            if rc != 0:
                continue
            P, R, tiles, slot = (int(v) for v in out)
            Wt = P - 2
            bands = -(-W // Wt)
            per_band = -(-(H * P) // 128)
            assert tiles == bands * per_band and slot >= R * P * rb and slot <= cap
            f = np.arange(per_band * 128)
            y, xx = f // P, f % P
            cover = np.zeros((H, W), int)
            for wb in range(bands):
                x = wb * Wt + xx
                valid = (xx < Wt) & (y < H) & (x < W)
                np.add.at(cover, (y[valid], x[valid]), 1)
            assert (cover == 1).all(), (H, W, P)
            fb = np.arange(per_band)
            start = (fb * 128) % P                                  # first position inside the slot
            assert (start + 127 + 2 * P + 2 < R * P).all(), (H, W, P, R)
            y0 = (fb * 128) // P                                    # slot = padded rows y0 .. y0 + R - 1
            last_row_needed = (fb * 128 + 127 + 2 * P + 2) // P
            assert (last_row_needed <= y0 + R - 1).all()


LOOP_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch.distributed as dist
from iterseg_b200 import _io, segmentation
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
T, shape = 7, (4, 8, 6)
data = np.stack([np.full(shape, t + 1, np.float32) for t in range(T)])
data[:, 0, 0, 0] = 8.0                                   # the frame maximum: vol /= max leaves (t + 1) / 8
store = _io.open_zarr(%(store)r, shape=(T,) + shape, chunks=(4, 4, 8), dtype=np.int32)
if rank == 0:
    store[3, ...] = np.full(shape, 99, np.int32)          # already segmented: warm restart must skip it
dist.barrier()
calls = []

def fake_segmenter(input_volume, current_output, chunk_size, margin, **config):
    """A plug-in processing function (segmentation.py:898-899 protocol): label = 8 * normalised value."""
    calls.append(float(input_volume[1, 1, 1]))
    current_output[1:-1, 1:-1, 1:-1] = np.rint(input_volume * 8).astype(np.uint32)

done = list(segmentation.segmentation_loop(None, data, (4, 4, 8), (0, 0, 0), store, fake_segmenter, {}))
mine = [t for t in range(rank, T, world) if t != 3]
assert done == mine, (rank, done, mine)
assert len(calls) == len(mine)
dist.barrier()
a = np.asarray(_io.open_zarr(%(store)r))
for t in range(T):
    want = 99 if t == 3 else t + 1
    assert (a[t].ravel()[1:] == want).all(), (t, a[t].ravel()[:4])
# opting out: every rank runs the whole series
out = np.zeros((T,) + shape, np.int32)
done = list(segmentation.segmentation_loop(None, data, (4, 4, 8), (0, 0, 0), out, fake_segmenter, {'shard': False}))
assert done == list(range(T))
dist.barrier()
dist.destroy_process_group()
sys.stdout.write('rank ' + str(rank) + ' ok\n')
sys.stdout.flush()
'''


def test_segmentation_loop_shards_frames_over_two_ranks_gloo(tmp_path):
    """The public frame loop under torch.distributed (gloo, world size 2, CPU) with a plug-in processing
    function of the reference's protocol: frames t = rank (mod 2), warm restart honoured, both ranks
    writing one store, `shard: False` opting out (segmentation.py:833-882)."""
    script = tmp_path / 'worker3.py'
    script.write_text(LOOP_WORKER % {'root': ROOT, 'store': str(tmp_path / 'lab')})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29615', str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'rank 0 ok' in r.stdout and 'rank 1 ok' in r.stdout
