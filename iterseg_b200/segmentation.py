"""Segmenter registry + per-frame driver: the reference's segmentation module.

Mirrors src/iterseg/segmentation.py for the affinity U-Net watershed path:
`affinity_unet_watershed` (:24-73), `affinity_watershed_prep_config` (:80-135),
`a_w_output_volume` (:138-140), `affinity_watershed_for_chunks` (:147-195),
`read_config_json` (:687-690), `segmentation_wrapper` (:700-830),
`segmentation_loop` (:833-882), `segment_single_volume` (:885-900),
`remove_sum_zero_slices` (:903-916) and the `segmenters` registry (:924-930).

What differs from the reference, on purpose:
* one frame goes host -> device once, the feature volume never leaves the
  device (the reference moves 2.6 MB in / 13.1 MB out per chunk, predict.py:119-123),
  labels come back once; `output_volume` (the reference's host scratch, zeroed at
  the end of every frame, :195) is left untouched -- i.e. zero, as the reference
  leaves it;
* without napari the frame loop runs inline (the reference uses a napari
  thread_worker unless debug=True, :808-828);
* the `.json` config branch accepts what the reference's documentation promises
  (`{"unet": "default" | "labels layer" | path}`) instead of raising
  UnboundLocalError (:98-107).
"""
import json
import os
import pathlib
from typing import Callable, Union

import numpy as np
import torch

from . import _io, _lib
from . import watershed as ws
from .predict import (load_unet, predict_chunk_feature_map, predict_frame_device,  # noqa: F401
                      process_chunks)
from . import unet as unet_mod


# ------------------------
# Affinity U-net Watershed
# ------------------------

def affinity_unet_watershed(napari_viewer, input_volume_layer, save_dir: Union[str, None] = None,
                            name: str = 'my-segmentation',
                            unet_or_config_file: Union[str, None] = None,
                            layer_reference: Union[str, None] = None,
                            chunk_size: Union[tuple, None] = (10, 256, 256),
                            margin: Union[tuple, None] = (1, 64, 64), debug: bool = False):
    """Same arguments as the reference (segmentation.py:24-73).  Returns the output labels
    layer (the reference's wrapper computes it but drops the return value)."""
    return segmentation_wrapper(affinity_watershed_for_chunks, affinity_watershed_prep_config,
                                napari_viewer, input_volume_layer, save_dir, name,
                                unet_or_config_file, layer_reference, chunk_size, margin, debug)


def affinity_watershed_prep_config(input_volume_layer, unet_or_config_file, reference_layer):
    unet = None
    affinities_extent = 1
    if isinstance(unet_or_config_file, pathlib.PurePath):
        unet_or_config_file = str(unet_or_config_file)
    if isinstance(unet_or_config_file, str):
        if unet_or_config_file.endswith('.json'):
            config = read_config_json(unet_or_config_file)
            unet = config.get('unet')
            if config.get('affinities_extent') is not None:
                affinities_extent = int(config['affinities_extent'])
            if unet == 'labels layer':
                unet = reference_layer.metadata['unet']
            if unet == 'default':
                unet = None
        elif unet_or_config_file.endswith('.pt') or unet_or_config_file.endswith('.pth'):
            unet = unet_or_config_file
        else:
            raise ValueError('Please provide a valid path to a pytorch unet - must end with .pt or .pth')
    elif unet_or_config_file is not None:
        raise ValueError('Please provide a valid path to a pytorch unet - must end with .pt or .pth')
    if unet is not None:
        if isinstance(unet, str):
            m = f'There was not file at the provided location: {unet}\nMake sure a pytorch unet lives here...'
            assert os.path.exists(unet), m
            assert unet.endswith('.pt') or unet.endswith('.pth')
        else:
            raise ValueError('Please provide a valid path to a pytorch unet - must end with .pt or .pth')
    if affinities_extent != 1:
        raise NotImplementedError('only affinities_extent = 1 (5 prediction channels) is supported')
    num_pred_channels = 3 * affinities_extent + 2
    output_volume = a_w_output_volume(input_volume_layer.data, num_pred_channels)
    unet = load_unet(unet)
    return {'unet': unet, 'output_volume': output_volume}


def a_w_output_volume(data, num_pred_channels, **kwargs):
    return np.zeros((num_pred_channels,) + tuple(data.shape[-3:]), dtype=np.float32)


def affinity_watershed_for_chunks(input_volume, current_output, chunk_size, margin, unet=None,
                                  output_volume=None, **kwargs):
    """The processing function of the plug-in protocol (segmentation.py:147-195):
    writes the labels of one frame IN PLACE into `current_output`, the padded
    (Z+2,Y+2,X+2) uint32 array the driver allocates (:890-895)."""
    if output_volume is None:
        raise ValueError('output_volume must not be None. Please ensure the output volume is supplied in the config dict')
    if unet is None:
        raise ValueError('unet must not be none. Please ensure a unet was loaded and added to the config dict in the config function')
    if not isinstance(unet, unet_mod.UNet):
        raise TypeError('unet must be an iterseg_b200.unet.UNet (use iterseg_b200.predict.load_unet)')
    _lib.require_device()
    dev = unet.device
    vol = input_volume if isinstance(input_volume, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(input_volume, dtype=np.float32))
    frame = vol.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    feats = predict_frame_device(unet, frame, tuple(int(c) for c in chunk_size), margin)
    shape_p = tuple(int(s) + 2 for s in frame.shape)
    out_is_dev = isinstance(current_output, torch.Tensor) and current_output.is_cuda
    labels = current_output.view(shape_p) if out_is_dev else \
        torch.zeros(shape_p, dtype=torch.int32, device=dev)
    _, counts, _, _ = ws.segment_features_device(feats, labels, affinities_channels=(0, 1, 2),
                                                 thresholding_channel=3, centroids_channel=4)
    LAST_COUNTS['counts'] = counts           # device int64[8]; [0] = number of labels of this frame
    if not out_is_dev:
        direct = (isinstance(current_output, np.ndarray) and current_output.flags.c_contiguous
                  and current_output.dtype in (np.uint32, np.int32)
                  and current_output.size == labels.numel())
        if direct:                                 # one D2H copy straight into the caller's array
            torch.from_numpy(current_output.reshape(-1).view(np.int32)).copy_(labels.view(-1))
        else:
            flat = current_output.reshape(-1)      # a view, like current_output.ravel() in :194
            flat[...] = labels.cpu().numpy().view(np.uint32).reshape(-1).astype(flat.dtype, copy=False)
        unet.check_overflow()                      # the copy above waited for the frame


# (addition) device-side counters of the most recent frame, for callers that need the number of
# labels without scanning the label volume (frame-sharded series: global label offsets)
LAST_COUNTS = {}


# ---------------
# Common funcions
# ---------------

def read_config_json(path_to_json):
    with open(path_to_json, 'r') as f:
        return json.load(f)


def segmentation_wrapper(processing_function: Callable, config_prep_function: Callable,
                         napari_viewer, input_volume_layer, save_dir: Union[str, None], name: str,
                         network_or_config_file: Union[str, None],
                         layer_reference: Union[str, None], chunk_size: tuple, margin: tuple,
                         debug: bool = False):
    viewer = napari_viewer
    config = config_prep_function(input_volume_layer, network_or_config_file, layer_reference)
    if config is None:
        config = {}
    save_path = None
    if save_dir is not None and not debug:
        save_path = os.path.join(str(save_dir), name + '.ome.zarr')
    data = input_volume_layer.data
    shape = data.shape
    scale = input_volume_layer.scale
    translate = input_volume_layer.translate
    layer_meta = {'scale': scale, 'translate': translate, 'name': name}
    if save_path is not None:
        output_labels = _io.save_labels_to_ome(save_path, layer_meta=layer_meta, shape=shape,
                                               chunks=chunk_size, dtype=np.int32)
    else:
        output_labels = _io.zeros(shape=shape, chunks=chunk_size, dtype=np.int32)
    output_layer = viewer.add_labels(output_labels, name=name, scale=scale, translate=translate)

    def handle_yields(yielded_val):
        viewer.dims.current_step = (yielded_val, 0, 0, 0)
        print(f"Segmented t = {yielded_val}")

    worker_factory = None
    if not debug:
        try:                                                  # the reference's threaded mode
            from napari.qt import thread_worker
            worker_factory = thread_worker(
                segmentation_loop, progress={'total': data.shape[0], 'desc': 'thread-progress'},
                connect={'yielded': handle_yields, 'errored': _raise})
        except Exception:
            worker_factory = None
    if worker_factory is not None:
        worker_factory(viewer, data, chunk_size, margin, output_labels, processing_function, config)
    else:
        for t in segmentation_loop(viewer, data, chunk_size, margin, output_labels,
                                   processing_function, config):
            if debug:
                print(f'Segmented frame {t}')
            else:
                handle_yields(t)
    return output_layer


def _raise(err):
    raise err


def segmentation_loop(viewer, data, chunk_size, margin, output_labels, processing_function, config):
    """One 3-D frame at a time; yields the time point when done (segmentation.py:833-882),
    including the warm restart: frames whose output is already non-zero are skipped.

    Additions that leave single-process results unchanged:
    * tzyx series run through `iterseg_b200.pipeline.SeriesPipeline` (several frames in flight,
      no host synchronisation on the compute streams);
    * under an initialised `torch.distributed` job (one process per GPU) the frames of a series
      are sharded `t = rank (mod world)` and every rank writes ITS frames into the shared output
      store (a zarr store on a file system all ranks see: partial chunk writes are in place, so
      ranks that share a t-chunk never touch each other's bytes); a single 3-D volume is sharded
      into z-slabs (`iterseg_b200.slab`).  Opt out with config['shard'] = False;
    * config['global_label_offsets'] = True makes label ids unique across the series (an
      exclusive prefix of the per-frame label counts, one NCCL all-gather of R int64 per step;
      the reference restarts at 1 in every frame, watershed.py:61-62)."""
    from . import distributed as idist
    rank, world = idist.world(config.get('process_group'))
    if not config.get('shard', True):
        rank, world = 0, 1
    ndim = data.ndim
    if ndim == 3:
        if world > 1 and processing_function is affinity_watershed_for_chunks:
            yield from _slab_volume(data, chunk_size, margin, output_labels, config)
            return
        output = segment_single_volume(np.asarray(data).astype(np.float32), chunk_size, config,
                                       margin, processing_function)
        output_labels[...] = output
        yield 0
        return
    frames = idist.shard_frames(data.shape[0], rank, world)
    core_kind = _pipeline_kind(processing_function, config, data)
    if core_kind is not None:
        yield from _pipelined_series(data, chunk_size, margin, output_labels, config, frames, core_kind,
                                     processing_function)
        return
    if config.get('global_label_offsets'):
        raise NotImplementedError('global_label_offsets needs the pipelined series loop')
    for t in frames:
        if np.any(output_labels[t]):
            continue
        input_volume = np.asarray(data[t]).astype(np.float32)
        current_output = segment_single_volume(input_volume, chunk_size, config, margin,
                                               processing_function)
        output_labels[t, ...] = current_output
        yield t


def _pipeline_kind(processing_function, config, data):
    if data.ndim != 4 or os.environ.get('ISG_NO_PIPELINE') is not None:
        return None
    if processing_function is affinity_watershed_for_chunks and isinstance(config.get('unet'), unet_mod.UNet) \
            and config.get('output_volume') is not None:
        return 'affinity'
    if processing_function is dog_blob_watershed_for_chunks and \
            all(k in config for k in ('min_sigma', 'max_sigma', 'threshold')):
        return 'dog'
    return None


def _slab_volume(data, chunk_size, margin, output_labels, config):
    """One 3-D volume under torch.distributed: z-slabs with halo planes (iterseg_b200.slab,
    BASELINE.json configs[3]); every rank stores its own planes."""
    from . import slab
    vol = data if isinstance(data, np.ndarray) else np.asarray(data)
    own, (z0, z1), n_labels = slab.segment_volume_slabs(
        vol, config['unet'], tuple(int(c) for c in chunk_size), margin,
        halo=int(config.get('slab_halo', 24)), group=config.get('process_group'))
    LAST_COUNTS['n_labels'] = n_labels
    output_labels[z0:z1, ...] = own.cpu().numpy()
    yield 0


def _pipelined_series(data, chunk_size, margin, output_labels, config, frames, kind, processing_function):
    """The frame loop of `segmentation_loop` for the two built-in processing functions: identical
    results and yield order, but the host never waits for a compute stream (see
    iterseg_b200/pipeline.py::SeriesPipeline): loader threads cast frame t+1.. into pinned
    buffers and evaluate the warm-restart test (segmentation.py:875-877), the copy stream moves
    frame t+1 and computes its min / max, the U-Net of frame t and the post stage of frame t-1
    run on their streams, writer threads store finished frames."""
    import collections
    import concurrent.futures as cf
    from . import distributed as idist
    from .pipeline import DogCore, FramePipeline, SeriesPipeline
    _lib.require_device()
    dev = config['unet'].device if kind == 'affinity' else torch.device('cuda', torch.cuda.current_device())
    shape = tuple(int(v) for v in data.shape[1:])
    chunk = tuple(int(c) for c in chunk_size)
    rank, world = idist.world(config.get('process_group')) if config.get('shard', True) else (0, 1)
    want_offsets = bool(config.get('global_label_offsets'))
    if want_offsets and kind != 'affinity':
        raise NotImplementedError('global_label_offsets is implemented for the affinity U-Net watershed only')
    # global label ids: one all-gather per step over ALL ranks of the job, sharded (step s = frames
    # sR..sR+R-1) or not (every rank runs its own copy of the loop, e.g. one series per rank: step s =
    # frame s of every rank); every rank must take the same number of steps
    o_rank, o_world = idist.world(config.get('process_group')) if want_offsets else (0, 1)
    n_steps = (int(data.shape[0]) + world - 1) // world if world > 1 else int(data.shape[0])
    offsets = idist.LabelOffsets(o_rank, o_world, dev, config.get('process_group')) if want_offsets else None

    depth = int(os.environ.get('ISG_PIPE_DEPTH', '2')) if kind == 'affinity' else 2   # U-Nets submitted ahead of the post stage

    def make_pipe():
        if kind == 'affinity':
            core = FramePipeline(config['unet'], shape, chunk, margin, depth=depth)
        else:
            core = DogCore(shape, dev, min_sigma=config['min_sigma'], max_sigma=config['max_sigma'],
                           threshold=config['threshold'])
        return SeriesPipeline(core)

    import time
    t_start = time.perf_counter()
    timing = {'frames': 0}
    LAST_COUNTS['timing'] = timing
    pipe = make_pipe()
    timing['setup_s'] = time.perf_counter() - t_start
    loaders = cf.ThreadPoolExecutor(max_workers=2, thread_name_prefix='isg-load')
    writers = cf.ThreadPoolExecutor(max_workers=4, thread_name_prefix='isg-write')
    SKIP = object()

    def load(t):
        if np.any(output_labels[t]):                             # warm restart (:875-876)
            return SKIP
        return pipe.load(data[t])

    def store(t, o, host, direct):
        try:
            pipe.outs[o]['ev_d2h'].synchronize()
            if not direct:
                output_labels[t, ...] = host.numpy()
        finally:
            pipe.release(o)            # also when the store raised: the main thread must never wait for this slot
        return t

    def direct_dst(t):
        """The caller's own memory for frame t when one D2H copy can land there."""
        if isinstance(output_labels, np.ndarray) and output_labels.dtype in (np.int32, np.uint32):
            dst = output_labels[t]
            if dst.flags.c_contiguous and dst.shape == shape:
                ten = torch.from_numpy(dst.view(np.int32))
                if ten.is_pinned():
                    return ten
        return None

    staged = collections.deque()        # (t, staging slot): H2D enqueued
    running = collections.deque()       # t: U-Net enqueued
    storing = collections.deque()       # writer futures, in frame order
    steps_done = 0

    def collect_one():
        nonlocal steps_done
        t = running.popleft()
        dst = direct_dst(t)
        on_counts = None
        if offsets is not None:
            on_counts = lambda counts: offsets.step(counts[0:1])      # noqa: E731
            steps_done += 1
        o, host, counts = pipe.collect(dst=dst, on_counts=on_counts)
        LAST_COUNTS['counts'] = counts
        storing.append(writers.submit(store, t, o, host, dst is not None))

    def finished(block_if_more_than):
        while storing and (storing[0].done() or len(storing) > block_if_more_than):
            t_done = storing.popleft().result()
            if kind == 'affinity':
                config['unet'].check_overflow()                  # that frame's D2H has completed
            now = time.perf_counter() - t_start
            timing.setdefault('first_frame_s', now)
            timing['last_frame_s'] = now
            timing['frames'] += 1
            yield t_done

    def slow_zero_frame(t):
        """A frame that contains zeros takes the reference's host route (segmentation.py:887-889:
        slices that sum to zero are stripped before `vol /= max`); shapes may differ from the
        series', so it runs outside the pipeline, after everything in flight has drained."""
        while running:
            collect_one()
        yield from finished(0)
        vol = np.asarray(data[t]).astype(np.float32)
        output_labels[t, ...] = segment_single_volume(vol, chunk_size, config, margin, processing_function)
        yield t

    try:
        todo = collections.deque(frames)
        loading = collections.deque()       # (t, future)
        while todo or loading or staged or running:
            # keep two frames staged: their H2D copies run on the copy stream beside everything else
            while (todo or loading) and len(staged) < 2:
                while todo and len(loading) < 3:
                    t = todo.popleft()
                    loading.append((t, loaders.submit(load, t)))
                t, fut = loading.popleft()
                res = fut.result()
                if res is SKIP:
                    if want_offsets:
                        raise NotImplementedError('a warm restart of a globally numbered series is not supported')
                    continue
                staged.append((t, pipe.h2d(res)))
            if staged and len(running) < depth:                  # a U-Net slot is free
                t, j = staged.popleft()
                if pipe.submit(j):
                    running.append(t)
                else:
                    if want_offsets:
                        raise NotImplementedError('global_label_offsets with a frame that contains zeros')
                    yield from slow_zero_frame(t)
                continue
            if running:
                collect_one()
            yield from finished(3)
        while running:
            collect_one()
        if offsets is not None:                                  # ranks with fewer frames keep the lock-step
            with torch.cuda.stream(pipe.core.s_post):
                while steps_done < n_steps:
                    offsets.step(None)
                    steps_done += 1
        yield from finished(0)
        if offsets is not None:
            LAST_COUNTS['global_total'] = offsets.total
    finally:
        pipe.close()                                             # wakes loader threads waiting for a buffer
        loaders.shutdown(wait=True, cancel_futures=True)
        writers.shutdown(wait=True)
        torch.cuda.synchronize(dev)
        pipe.recycle()                                           # pinned staging buffers back to the pool


def segment_single_volume(input_volume, chunk_size, config, margin, processing_function):
    if input_volume.min() == 0:
        input_volume = remove_sum_zero_slices(input_volume)
    input_volume /= np.max(input_volume)
    current_output = np.zeros(tuple(s + 2 for s in input_volume.shape), dtype=np.uint32)
    crop = (slice(1, -1),) * current_output.ndim
    processing_function(input_volume, current_output, chunk_size, margin, **config)
    return current_output[crop]


def remove_sum_zero_slices(input_volume):
    for ax_i in range(input_volume.ndim):
        keep = [i for i in range(input_volume.shape[ax_i])
                if np.take(input_volume, i, axis=ax_i).sum() != 0]
        input_volume = np.take(input_volume, keep, axis=ax_i)
    return input_volume


# ------------------
# DoG Blob Watershed
# ------------------

def dog_blob_watershed(napari_viewer, input_volume_layer, save_dir=None, name='labels-prediction',
                       config_file=None, layer_reference=None, chunk_size=(10, 256, 256),
                       margin=(1, 64, 64), debug=False):
    """segmentation.py:548-589, same arguments."""
    return segmentation_wrapper(dog_blob_watershed_for_chunks, dog_blob_watershed_prep_config,
                                napari_viewer, input_volume_layer, save_dir, name, config_file,
                                layer_reference, chunk_size, margin, debug)


def dog_blob_watershed_prep_config(input_volume_layer, unet_or_config_file, reference_layer,
                                   max_sigma=1.5, min_sigma=1, threshold=0.02):
    """segmentation.py:653-675.  (The reference subscripts `config.get[...]`, a TypeError for any
    config file; here the JSON keys max_sigma / min_sigma / threshold are honoured.)"""
    if unet_or_config_file is not None:
        config = read_config_json(unet_or_config_file)
        max_sigma = config.get('max_sigma', max_sigma) if config.get('max_sigma') is not None else max_sigma
        min_sigma = config.get('min_sigma', min_sigma) if config.get('min_sigma') is not None else min_sigma
        threshold = config.get('threshold', threshold) if config.get('threshold') is not None else threshold
    return {'max_sigma': max_sigma, 'min_sigma': min_sigma, 'threshold': threshold}


def _dog_params(min_sigma, max_sigma, threshold, sigma_ratio=1.6, overlap=0.5):
    """isg_dog_params for blob_dog(min_sigma, max_sigma, threshold) with scikit-image's defaults
    (sigma_ratio 1.6, overlap 0.5, exclude_border False): k = int(log(max/min)/log(ratio) + 1) DoG
    layers from the sigma list min_sigma * ratio^i, i = 0..k (segmentation.py:637-638)."""
    import math
    if not (np.isscalar(min_sigma) and np.isscalar(max_sigma)):
        raise NotImplementedError('only scalar min_sigma / max_sigma are supported')
    k = int(math.log(float(max_sigma) / float(min_sigma)) / math.log(sigma_ratio) + 1)
    if k < 1:
        raise ValueError('max_sigma must not be smaller than min_sigma')
    if k + 1 > _lib.DOG_MAX_SIGMAS:
        raise NotImplementedError(f'blob_dog with {k} DoG layers: at most {_lib.DOG_MAX_SIGMAS - 1} are supported')
    sl = [float(min_sigma) * sigma_ratio ** i for i in range(k + 1)]
    p = _lib.DogParams()
    for i, sg in enumerate((float(min_sigma), float(max_sigma), sl[0], sl[1])):
        w, r = ws.gaussian_half_kernel(sg)
        if k > 1:
            if i >= 2:
                continue
            if r > _lib.GAUSS_MAX_RADIUS:
                raise ValueError(f'sigma {sg} needs a Gaussian radius of {r} > {_lib.GAUSS_MAX_RADIUS}')
            p.mask_radius[i] = r
            for j in range(r + 1):
                p.mask_weights[i][j] = float(w[j])
            continue
        if r > 11:
            raise ValueError(f'sigma {sg} needs a Gaussian radius of {r} > 11')
        p.radius[i] = r
        for j in range(r + 1):
            p.weights[i][j] = float(w[j])
    p.threshold = float(threshold)
    p.scale_factor = float(np.float32(1.0 / (sigma_ratio - 1)))
    # _prune_blobs with ONE common sigma: spheres of radius sigma * sqrt(3); blobs closer than this
    # overlap by > `overlap`
    r = sl[0] * math.sqrt(3)
    d2max = 0
    for d2 in range(1, int((2 * r) ** 2) + 2):
        d = math.sqrt(d2)
        if d > 2 * r:
            break
        vol = math.pi / (12 * d) * (2 * r - d) ** 2 * (d * d + 4 * d * r)
        if vol / (4.0 / 3 * math.pi * r ** 3) > overlap:
            d2max = d2
    p.prune_d2 = d2max
    p.prune_radius = int(math.isqrt(d2max)) if d2max else 0
    p.n_layers = k
    p.overlap = float(overlap)
    if k > 1:
        for i, sg in enumerate(sl):
            w, r = ws.gaussian_half_kernel(sg)
            if r > _lib.GAUSS_MAX_RADIUS:
                raise ValueError(f'sigma {sg:.3f} (layer {i}) needs a Gaussian radius of {r} > {_lib.GAUSS_MAX_RADIUS}')
            p.layer_sigma[i] = sg
            p.layer_radius[i] = r
            for j in range(r + 1):
                p.layer_weights[i][j] = float(w[j])
    return p


def dog_blob_segment_device(frame, labels, min_sigma=1, max_sigma=1.5, threshold=0.02, distance=None,
                            max_seeds=None):
    """Device core of the DoG blob segmenter: frame (Z,Y,X) float32 CUDA tensor, labels
    (Z+2,Y+2,X+2) int32 zeros (written in place).  Returns (mask uint8 padded, counts int64[4])."""
    lib = _lib.load()
    assert frame.is_cuda and frame.dtype == torch.float32 and frame.is_contiguous()
    Z, Y, X = (int(v) for v in frame.shape)
    shape_p = (Z + 2, Y + 2, X + 2)
    assert tuple(labels.shape) == shape_p and labels.dtype == torch.int32 and labels.is_contiguous()
    dev = frame.device
    if max_seeds is None:
        max_seeds = max(1 << 16, (Z * Y * X) // 8)
    p = _dog_params(min_sigma, max_sigma, threshold)
    nbytes = lib.isg_dog_workspace_bytes_layers(Z, Y, X, max_seeds, int(p.n_layers))
    wsb = ws._workspace('dog', (Z, Y, X, max_seeds, int(p.n_layers)), nbytes, dev)
    mask = torch.empty(shape_p, dtype=torch.uint8, device=dev)
    counts = torch.zeros(4, dtype=torch.int64, device=dev)
    import ctypes
    with torch.cuda.device(dev):
        rc = lib.isg_dog_blob_segment(frame.data_ptr(), Z, Y, X, ctypes.byref(p), labels.data_ptr(),
                                      mask.data_ptr(), distance.data_ptr() if distance is not None else None,
                                      max_seeds, counts.data_ptr(), wsb.data_ptr(), wsb.numel(),
                                      _lib.stream_ptr())
    _lib.check(rc, 'isg_dog_blob_segment')
    return mask, counts


def dog_blob_watershed_for_chunks(input_volume, current_output, chunk_size, margin, min_sigma,
                                  max_sigma, threshold, **kwargs):
    """The DoG processing function of the plug-in protocol (segmentation.py:592-650): labels of
    one frame written IN PLACE into the padded `current_output`.  chunk_size / margin are unused,
    as in the reference."""
    _lib.require_device()
    dev = current_output.device if isinstance(current_output, torch.Tensor) and current_output.is_cuda \
        else torch.device('cuda', torch.cuda.current_device())
    vol = input_volume if isinstance(input_volume, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(input_volume, dtype=np.float32))
    frame = vol.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    shape_p = tuple(int(s) + 2 for s in frame.shape)
    out_is_dev = isinstance(current_output, torch.Tensor) and current_output.is_cuda
    labels = current_output.view(shape_p) if out_is_dev else torch.zeros(shape_p, dtype=torch.int32, device=dev)
    dog_blob_segment_device(frame, labels, min_sigma, max_sigma, threshold)
    if not out_is_dev:
        flat = current_output.reshape(-1)
        if (isinstance(current_output, np.ndarray) and current_output.flags.c_contiguous
                and current_output.dtype in (np.uint32, np.int32) and current_output.size == labels.numel()):
            torch.from_numpy(flat.view(np.int32)).copy_(labels.view(-1))
        else:
            flat[...] = labels.cpu().numpy().view(np.uint32).reshape(-1).astype(flat.dtype, copy=False)


segmenters = {
    'affinity-unet-watershed': affinity_unet_watershed,
    'DoG-blob-watershed': dog_blob_watershed,
}
