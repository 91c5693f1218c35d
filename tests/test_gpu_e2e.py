"""End-to-end GPU tests through the reference-facing entry points."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CHUNK, MARGIN = (10, 256, 256), (1, 64, 64)


@pytest.fixture(scope='module')
def net():
    from iterseg_b200 import _lib, synth, unet
    _lib.require_device()
    n = unet.UNet()
    n.load_state_dict(synth.structured_state_dict(0))
    return n.cuda()


@pytest.fixture(scope='module')
def frame():
    from iterseg_b200 import synth
    return synth.platelet_frame((12, 300, 300), seed=1)


def _segment(net, vol):
    from iterseg_b200 import segmentation
    cur = np.zeros(tuple(s + 2 for s in vol.shape), np.uint32)
    segmentation.affinity_watershed_for_chunks(vol.copy(), cur, CHUNK, MARGIN, unet=net,
                                               output_volume=np.zeros((5,) + vol.shape, np.float32))
    return cur


def test_reproducible_and_batch_invariant(net, frame):
    """BatchNorm statistics are reduced with integer atomics: bitwise identical run to run,
    and a chunk gives the same features alone or batched with others."""
    from iterseg_b200 import predict
    vol = torch.from_numpy(frame).cuda()
    a = predict.predict_frame_device(net, vol, CHUNK, MARGIN).clone()
    b = predict.predict_frame_device(net, vol, CHUNK, MARGIN).clone()
    assert torch.equal(a, b)
    st, lo, hi = predict._chunk_tables(vol.shape, CHUNK, MARGIN)
    c = torch.zeros_like(a)
    for i in range(len(st)):
        predict.predict_frame_device(net, vol, CHUNK, MARGIN, out=c,
                                     tables=(st[i:i + 1], lo[i:i + 1], hi[i:i + 1]))
    assert torch.equal(a, c)
    assert np.array_equal(_segment(net, frame), _segment(net, frame))


def test_pipeline_equals_oracle_post_on_gpu_features(net, frame):
    """Plumbing: the fused frame pipeline == the oracle's post-U-Net stage applied to the
    features the GPU U-Net produced (bit-exact integer/index work)."""
    from iterseg_b200 import predict
    from oracle import post
    cur = _segment(net, frame)
    feats = predict.predict_frame_device(net, torch.from_numpy(frame).cuda(), CHUNK, MARGIN).cpu().numpy()
    want = np.zeros_like(cur)
    post.segment_output_image(feats, out=want.ravel())
    assert np.array_equal(cur, want)


def test_end_to_end_vs_cpu_oracle(net, frame):
    """BASELINE.json gate: VI <= 0.01 and matched-object F1 >= 0.99 against the reference
    CPU path (fp32 U-Net + scipy/numpy stage + heap flood) on a platelet-shaped frame."""
    from iterseg_b200 import synth
    from oracle import metrics, post, unet_ref
    sd = synth.structured_state_dict(0)
    feats = unet_ref.predict_frame(frame, sd)
    seg_o, _, _ = post.segment_output_image(feats)
    seg_g = _segment(net, frame)[1:-1, 1:-1, 1:-1]
    vi = sum(metrics.variation_of_information(seg_o, seg_g))
    f1 = metrics.matched_f1(seg_o, seg_g, 0.5)
    assert vi <= 0.01, (vi, int(seg_o.max()), int(seg_g.max()), int((seg_o > 0).sum()), int((seg_g > 0).sum()))
    assert f1 >= 0.99, f1


def test_end_to_end_gate_on_a_full_baseline_frame(net):
    """The same gate at the size BASELINE.json quotes (configs[0]/[1]): one 33x512x512 frame, 36 chunks of
    (10,256,256) margin (1,64,64).  Reference arm: fp32 torch-CPU U-Net (train-mode BN) on all 36
    chunks + the scipy/numpy post stage + the heap flood."""
    from iterseg_b200 import synth
    from oracle import metrics, post, unet_ref
    vol = synth.platelet_frame((33, 512, 512), seed=0)
    sd = synth.structured_state_dict(0)
    feats = unet_ref.predict_frame(vol, sd)
    seg_o, _, _ = post.segment_output_image(feats)
    seg_g = _segment(net, vol)[1:-1, 1:-1, 1:-1]
    assert seg_o.max() > 1000
    vi = sum(metrics.variation_of_information(seg_o, seg_g))
    f1 = metrics.matched_f1(seg_o, seg_g, 0.5)
    assert vi <= 0.01, vi
    assert f1 >= 0.99, f1


def test_segment_data_timeseries_zarr_and_warm_restart(net, tmp_path):
    from iterseg_b200 import _dock_widgets, synth, viewer
    shape = (10, 256, 256)
    data = np.stack([synth.platelet_frame(shape, seed=s) for s in (3, 4)])
    path = str(tmp_path / 'net.pt')
    torch.save(net.state_dict(), path)
    v = viewer.HeadlessViewer()
    layer = viewer.Image(data, name='img', scale=(1, 4, 1, 1), translate=(0, 0, 0, 0))
    out_layer = _dock_widgets.segment_data(v, layer, save_dir=str(tmp_path), name='seg',
                                           segmenter='affinity-unet-watershed',
                                           network_or_config_file=path, chunk_size=CHUNK,
                                           margin=MARGIN, debug=False)
    store = tmp_path / 'seg.ome.zarr'
    attrs = json.load(open(store / '.zattrs'))
    assert attrs['image-label'] == {}
    ms = attrs['multiscales'][0]
    assert [a['name'] for a in ms['axes']] == ['t', 'z', 'y', 'x']
    assert ms['datasets'][0]['coordinateTransformations'][0] == {'type': 'scale', 'scale': [1.0, 4.0, 1.0, 1.0]}
    zarray = json.load(open(store / '0' / '.zarray'))
    assert zarray['dtype'] == '<i4' and zarray['shape'] == [2, 10, 256, 256]
    labels = np.asarray(out_layer.data)
    assert labels.shape == data.shape and labels.dtype == np.int32
    for t in range(2):
        want = _segment(net, (data[t] / data[t].max()).astype(np.float32))[1:-1, 1:-1, 1:-1]
        assert np.array_equal(labels[t].astype(np.uint32), want)
        assert labels[t].max() > 10
    # warm restart (segmentation.py:874-876): frames that already hold labels are skipped
    from iterseg_b200 import segmentation
    arr = out_layer.data
    arr[0] = np.full(shape, 7, np.int32)
    done = list(segmentation.segmentation_loop(v, data, CHUNK, MARGIN, arr,
                                               segmentation.affinity_watershed_for_chunks,
                                               {'unet': net, 'output_volume': np.zeros(1)}))
    assert done == []
    assert (np.asarray(arr[0]) == 7).all()


def test_single_volume_3d_input(net):
    from iterseg_b200 import segmentation, synth, viewer
    vol = synth.platelet_frame((10, 256, 256), seed=9)
    v = viewer.HeadlessViewer()
    out = segmentation._io.zeros(vol.shape, CHUNK, np.int32)
    done = list(segmentation.segmentation_loop(v, vol, CHUNK, MARGIN, out,
                                               segmentation.affinity_watershed_for_chunks,
                                               {'unet': net, 'output_volume': np.zeros(1)}))
    assert done == [0] and np.asarray(out).max() > 10


def test_errors_match_reference(net):
    from iterseg_b200 import segmentation
    cur = np.zeros((12, 258, 258), np.uint32)
    vol = np.ones((10, 256, 256), np.float32)
    with pytest.raises(ValueError):
        segmentation.affinity_watershed_for_chunks(vol, cur, CHUNK, MARGIN, unet=net, output_volume=None)
    with pytest.raises(ValueError):
        segmentation.affinity_watershed_for_chunks(vol, cur, CHUNK, MARGIN, unet=None,
                                                   output_volume=np.zeros(1))


def test_label_offsets_kernel():
    from iterseg_b200 import distributed as d
    lab = torch.tensor([0, 1, 2, 0, 5], dtype=torch.int32, device='cuda')
    d.add_label_offset_(lab, 10)
    assert lab.tolist() == [0, 11, 12, 0, 15]


def test_frame_pipeline_equals_sequential(net):
    """Two frames in flight (post of frame i || U-Net of frame i+1 on two streams) must give the
    very same labels as the sequential calls."""
    from iterseg_b200 import predict, synth, watershed as ws
    from iterseg_b200.pipeline import FramePipeline
    shape, chunk, margin = (20, 128, 128), (10, 64, 64), (1, 16, 16)
    dev = net.device
    frames = [torch.from_numpy(synth.platelet_frame(shape, seed=s)).to(dev) for s in (11, 12, 13)]
    want = []
    for f in frames:
        feats = predict.predict_frame_device(net, f, chunk, margin)
        lab = torch.zeros(tuple(s + 2 for s in shape), dtype=torch.int32, device=dev)
        ws.segment_features_device(feats, lab)
        want.append(lab.cpu().numpy())
    pipe = FramePipeline(net, shape, chunk, margin)
    got = []
    pipe.submit(frames[0])
    for i in range(3):
        if i + 1 < 3:
            pipe.submit(frames[i + 1])
        lab, counts = pipe.collect()
        pipe.drain_to()
        torch.cuda.synchronize()
        got.append(lab.cpu().numpy().copy())
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert want[0].max() > 0


def test_series_loop_pipelined_equals_plain(net, monkeypatch):
    """segmentation_loop on a tzyx series: the pipelined loop (two frames in flight) yields the same
    time points in the same order and writes the same labels as the frame-at-a-time loop, including
    the warm-restart skip of a frame in the middle."""
    from iterseg_b200 import segmentation, synth
    shape = (10, 128, 128)
    chunk, margin = (10, 64, 64), (1, 16, 16)
    data = np.stack([synth.platelet_frame(shape, seed=s) * np.float32(0.7) for s in (21, 22, 23, 24)])
    cfg = {'unet': net, 'output_volume': np.zeros(1)}

    def run():
        out = np.zeros(data.shape, np.int32)
        out[2] = 5                                           # already segmented: must be skipped
        order = list(segmentation.segmentation_loop(None, data, chunk, margin, out,
                                                    segmentation.affinity_watershed_for_chunks, cfg))
        return order, out

    order_p, out_p = run()
    monkeypatch.setenv('ISG_NO_PIPELINE', '1')
    order_s, out_s = run()
    assert order_p == order_s == [0, 1, 3]
    assert np.array_equal(out_p, out_s)
    assert (out_p[2] == 5).all() and out_p[0].max() > 0
