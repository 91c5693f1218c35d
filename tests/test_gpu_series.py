"""Series loop (segmentation.py:833-882) on the device pipeline: stores, global label offsets,
the DoG segmenter, plan / workspace reuse, the capped flood workspace."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPE, CHUNK, MARGIN = (10, 128, 128), (10, 64, 64), (1, 16, 16)


@pytest.fixture(scope='module')
def net():
    from iterseg_b200 import _lib, synth, unet
    _lib.require_device()
    n = unet.UNet()
    n.load_state_dict(synth.structured_state_dict(0))
    return n.cuda()


@pytest.fixture(scope='module')
def data():
    from iterseg_b200 import synth
    return np.stack([synth.platelet_frame(SHAPE, seed=s) * np.float32(0.7) for s in range(31, 38)])


def _plain(net, data, monkeypatch):
    from iterseg_b200 import segmentation
    monkeypatch.setenv('ISG_NO_PIPELINE', '1')
    out = np.zeros(data.shape, np.int32)
    order = list(segmentation.segmentation_loop(None, data, CHUNK, MARGIN, out,
                                                segmentation.affinity_watershed_for_chunks,
                                                {'unet': net, 'output_volume': np.zeros(1)}))
    monkeypatch.delenv('ISG_NO_PIPELINE')
    return order, out


def test_series_into_zarr_store_and_numpy(net, data, tmp_path, monkeypatch):
    """Seven frames through the pipelined loop into (a) a pageable numpy array, (b) a pinned
    array (direct D2H), (c) a zarr store whose t-chunk holds several frames (in-place partial
    chunk writes by the writer threads): all equal the frame-at-a-time loop."""
    from iterseg_b200 import _io, segmentation
    order_s, want = _plain(net, data, monkeypatch)
    cfg = {'unet': net, 'output_volume': np.zeros(1)}
    outs = {'numpy': np.zeros(data.shape, np.int32),
            'pinned': torch.zeros(data.shape, dtype=torch.int32).pin_memory().numpy(),
            'zarr': _io.open_zarr(str(tmp_path / 'lab'), shape=data.shape, chunks=(4, 10, 64), dtype=np.int32)}
    for name, out in outs.items():
        order = list(segmentation.segmentation_loop(None, data, CHUNK, MARGIN, out,
                                                    segmentation.affinity_watershed_for_chunks, cfg))
        assert order == order_s == list(range(len(data))), name
        assert np.array_equal(np.asarray(out), want), name
    assert want.max() > 0
    # non-float32, non-contiguous input takes the loader's cast (np.asarray(data[t]).astype(float32))
    d64 = data.astype(np.float64)[:, :, ::1, :]
    out = np.zeros(data.shape, np.int32)
    list(segmentation.segmentation_loop(None, d64, CHUNK, MARGIN, out,
                                        segmentation.affinity_watershed_for_chunks, cfg))
    assert np.array_equal(out, want)


def test_series_zero_frame_takes_the_reference_route(net, data, monkeypatch):
    """A frame whose minimum is 0 leaves the pipeline (segmentation.py:887-888) and still gives what
    the frame-at-a-time loop gives; the frames around it stay pipelined."""
    from iterseg_b200 import segmentation
    d = data[:4].copy()
    d[1, 0, 0, 0] = 0.0                                  # min == 0, no all-zero slice: shape unchanged
    order_s, want = _plain(net, d, monkeypatch)
    out = np.zeros(d.shape, np.int32)
    order = list(segmentation.segmentation_loop(None, d, CHUNK, MARGIN, out,
                                                segmentation.affinity_watershed_for_chunks,
                                                {'unet': net, 'output_volume': np.zeros(1)}))
    assert order == order_s == [0, 1, 2, 3]
    assert np.array_equal(out, want)


def test_series_global_label_offsets(net, data, monkeypatch):
    """config['global_label_offsets']: frame t's non-zero labels are shifted by the number of
    labels of the frames before it (device-resident running prefix, no host read-back)."""
    from iterseg_b200 import segmentation
    _, want = _plain(net, data, monkeypatch)
    out = np.zeros(data.shape, np.int32)
    cfg = {'unet': net, 'output_volume': np.zeros(1), 'global_label_offsets': True}
    list(segmentation.segmentation_loop(None, data, CHUNK, MARGIN, out,
                                        segmentation.affinity_watershed_for_chunks, cfg))
    off = 0
    for t in range(len(data)):
        exp = np.where(want[t] > 0, want[t] + off, 0)
        assert np.array_equal(out[t], exp), t
        off += int(want[t].max())
    assert int(segmentation.LAST_COUNTS['global_total'].item()) == off


def test_dog_series_pipelined_equals_plain(monkeypatch):
    from iterseg_b200 import segmentation, synth
    data = np.stack([synth.platelet_frame((12, 96, 96), seed=s) for s in (41, 42, 43)])
    cfg = {'min_sigma': 1, 'max_sigma': 1.5, 'threshold': 0.02}

    def run():
        out = np.zeros(data.shape, np.int32)
        order = list(segmentation.segmentation_loop(None, data, CHUNK, MARGIN, out,
                                                    segmentation.dog_blob_watershed_for_chunks, dict(cfg)))
        return order, out

    order_p, out_p = run()
    monkeypatch.setenv('ISG_NO_PIPELINE', '1')
    order_s, out_s = run()
    assert order_p == order_s == [0, 1, 2]
    assert np.array_equal(out_p, out_s) and out_p.max() > 3


def test_plans_share_one_workspace_and_tables_are_per_call(net):
    """plans are keyed by geometry, the chunk tables are uploaded per call on the caller's
    stream, all plans of a network share one workspace, and batches of different sizes alternate
    without rebuilding anything."""
    from iterseg_b200 import predict, synth
    vol = torch.from_numpy(synth.platelet_frame((12, 160, 160), seed=5)).cuda()
    st, lo, hi = predict._chunk_tables(vol.shape, CHUNK, MARGIN)
    want = predict.predict_frame_device(net, vol, CHUNK, MARGIN).clone()
    ws0 = net._workspace
    n_plans = len(net._plans)
    got = torch.zeros_like(want)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())       # vol / got were produced on the default stream
    with torch.cuda.stream(s):                       # a non-default stream, batches of 5 then the rest
        for b in range(0, len(st), 5):
            net.forward_chunks(vol, CHUNK, st[b:b + 5], lo[b:b + 5], hi[b:b + 5], out=got)
    s.synchronize()
    assert torch.equal(got, want)
    assert net._workspace is ws0                      # smaller batches reuse the block
    assert len(net._plans) <= n_plans + 2
    # the same plan, different tables, back to back: results must not mix
    a = torch.zeros_like(want)
    b_ = torch.zeros_like(want)
    net.forward_chunks(vol, CHUNK, st[:5], lo[:5], hi[:5], out=a)
    net.forward_chunks(vol, CHUNK, st[5:10], lo[5:10], hi[5:10], out=b_)
    torch.cuda.synchronize()
    sel_a = torch.zeros(want.shape[1:], dtype=torch.bool, device='cuda')
    for i in range(5):
        z, y, x = (int(v) for v in st[i])
        sel_a[z + lo[i][0]:z + hi[i][0], y + lo[i][1]:y + hi[i][1], x + lo[i][2]:x + hi[i][2]] = True
    assert torch.equal(a[:, sel_a], want[:, sel_a])
    assert not b_[:, sel_a].any()
    # back to the whole-frame plan after other plans used (and overwrote) the shared workspace
    assert torch.equal(predict.predict_frame_device(net, vol, CHUNK, MARGIN), want)


def test_capped_flood_workspace_fails_loudly_then_succeeds():
    """isg_segment_features with a workspace that cannot hold the multi-seed components raises
    ISG_ERR_WORKSPACE (nothing silently dropped); the full workspace gives the oracle's labels."""
    from iterseg_b200 import _lib, synth, watershed as ws
    from oracle import post as opost
    lab = synth.platelet_labels((8, 96, 96), n_objects=60, seed=3)
    feats_h = synth.analytic_features(lab, 3)
    feats = torch.from_numpy(feats_h).cuda()
    labels = torch.zeros((10, 98, 98), dtype=torch.int32, device='cuda')
    with pytest.raises(_lib.IsgError) as ei:
        ws.segment_features_device(feats, labels, max_flood_nodes=4)
    assert ei.value.status == _lib.ISG_ERR_WORKSPACE
    labels.zero_()
    ws.segment_features_device(feats, labels, max_flood_nodes=8 * 96 * 96 // 2)
    want = np.zeros((10, 98, 98), np.uint32)
    opost.segment_output_image(feats_h, out=want.ravel())
    assert np.array_equal(labels.cpu().numpy().view(np.uint32), want)


@pytest.mark.parametrize('staging', ['fused', 'zerocopy'])
def test_input_staging_modes_are_bit_identical(net, data, monkeypatch, staging):
    """north_star (1): the chunks staged straight from the pinned volume (`zerocopy`: no H2D copy, the
    first U-Net kernel and the min/max kernel read the pinned frame in place, `vol /= max` fused into
    the loads) and the copy-engine path with the fused division give the labels of the default path."""
    from iterseg_b200 import segmentation
    _, want = _plain(net, data, monkeypatch)
    monkeypatch.setenv('ISG_INPUT_STAGING', staging)
    cfg = {'unet': net, 'output_volume': np.zeros(1)}
    for src in (torch.from_numpy(data.copy()).pin_memory().numpy(), data.astype(np.float64)):
        out = np.zeros(data.shape, np.int32)
        order = list(segmentation.segmentation_loop(None, src, CHUNK, MARGIN, out,
                                                    segmentation.affinity_watershed_for_chunks, cfg))
        assert order == list(range(len(data)))
        assert np.array_equal(out, want)
