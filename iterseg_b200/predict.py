"""Chunked prediction with margin overlap: the reference's predict module.

Mirrors src/iterseg/predict.py: `load_unet` (:25-35), `make_chunks` (:38-61),
`process_chunks` (:64-96), `predict_chunk_feature_map` (:100-126),
`get_device` (:130-135), `IGNORE_CUDA` (:19).

There is no overlap *blending* in the reference: `make_chunks` yields an exact
tiling, every output voxel is written by the cropped interior of exactly one
chunk.  Here all chunks of a frame run as ONE batch on the device (each chunk
keeps its own BatchNorm statistics, as in the reference where a chunk is a
batch of one) and the crop-and-place is done by the kernel that writes the
feature volume.
"""
import itertools
import os

import numpy as np
import torch

from . import _lib
from . import unet as unet_mod

IGNORE_CUDA = False      # kept for interface compatibility; this implementation is CUDA-only

DEFAULT_UNET_PATH = os.path.join(os.path.dirname(__file__), 'data', '232208_161159_plateseg.pt')

# chunks of one forward_chunks call: an upper bound; `chunks_per_batch` lowers it so that the
# activation workspace (~0.36 GB per (10,256,256) chunk) uses at most half of the free memory
MAX_CHUNKS_PER_BATCH = 72


def chunks_per_batch(unet, chunk_size, n_chunks):
    """How many chunks one forward_chunks call takes: all of them up to MAX_CHUNKS_PER_BATCH,
    fewer when the device is short of memory (the workspace the network already holds counts as
    available).  Raises when not even one chunk fits."""
    n = max(1, min(int(n_chunks), MAX_CHUNKS_PER_BATCH))
    held = unet._workspace.numel() if getattr(unet, '_workspace', None) is not None else 0
    if unet_mod.workspace_bytes(n, chunk_size) <= held:
        return n                                       # steady state: no driver query per frame
    free, _ = torch.cuda.mem_get_info(unet.device)
    budget = held + free // 2
    while n > 1 and unet_mod.workspace_bytes(n, chunk_size) > budget:
        n = (n + 1) // 2
    if unet_mod.workspace_bytes(n, chunk_size) > held + free:
        raise _lib.IsgError(f'not enough device memory for one chunk of {tuple(chunk_size)} '
                            f'({unet_mod.workspace_bytes(1, chunk_size) / 2**30:.2f} GiB of workspace)')
    return n


def get_device():
    _lib.require_device()
    return torch.device('cuda', torch.cuda.current_device())


def load_unet(u_state_fn=DEFAULT_UNET_PATH):
    """predict.py:25-35.  The network file is the reference's: a pickled state_dict."""
    if u_state_fn is None:
        u_state_fn = DEFAULT_UNET_PATH
    if not os.path.exists(u_state_fn):
        raise FileNotFoundError(
            f'no network file at {u_state_fn} (the reference ships its default network as a large '
            'blob that is not part of this repository: pass a .pt/.pth state_dict explicitly)')
    u = unet_mod.UNet(in_channels=1, out_channels=5)
    device = get_device()
    u.load_state_dict(torch.load(u_state_fn, map_location='cpu'))
    u.to(device)
    return u


def make_chunks(arr_shape, chunk_shape, margin):
    """predict.py:38-61: per axis, chunk starts step by chunk-2*margin with the last one
    clamped to arr-chunk; every chunk keeps [margin, chunk-margin), the first from 0, the
    last whatever is still uncovered.  Returns (starts, crops) as lists of tuples in
    itertools.product order (z-major)."""
    ndim = len(arr_shape)
    if isinstance(margin, (int, np.integer)):
        margin = [int(margin)] * ndim
    starts, crops = [], []
    for arr, chk, mrg in zip(arr_shape, chunk_shape, margin):
        arr, chk, mrg = int(arr), int(chk), int(mrg)
        if arr < chk:
            raise ValueError(f'array extent {arr} is smaller than the chunk extent {chk}')
        if chk - 2 * mrg <= 0:
            raise ValueError(f'margin {mrg} leaves nothing of a chunk of {chk}')
        start = list(range(0, arr - 2 * mrg, chk - 2 * mrg))
        start[-1] = arr - chk
        if len(start) > 1 and start[-1] == start[-2]:
            start.pop()
        crop = [[mrg, chk - mrg] for _ in start]
        crop[0][0] = 0
        crop[-1][0] = chk - (arr - sum(c[1] - c[0] for c in crop[:-1]))
        crop[-1][1] = chk
        starts.append(np.asarray(start))
        crops.append(np.asarray(crop))
    chunk_starts = list(itertools.product(*starts))
    chunk_crops = list(itertools.product(*crops))
    return chunk_starts, chunk_crops


def _chunk_tables(shape, chunk_size, margin):
    starts, crops = make_chunks(shape, chunk_size, margin)
    st = np.asarray(starts, dtype=np.int32).reshape(-1, 3)
    cr = np.asarray(crops, dtype=np.int32).reshape(-1, 3, 2)
    return st, np.ascontiguousarray(cr[:, :, 0]), np.ascontiguousarray(cr[:, :, 1])


def predict_frame_device(unet, frame, chunk_size, margin, out=None, tables=None, norm_max=None):
    """All chunks of one frame on the device: frame (Z,Y,X) float32 CUDA tensor ->
    (5,Z,Y,X) float32 CUDA tensor.  `tables` = (starts, crop_lo, crop_hi) restricts the
    work to a subset of the global chunk list (spatial sharding)."""
    st, lo, hi = tables if tables is not None else _chunk_tables(frame.shape, chunk_size, margin)
    if out is None:
        out = torch.zeros((5,) + tuple(frame.shape), dtype=torch.float32, device=unet.device)
    per = chunks_per_batch(unet, chunk_size, len(st))
    for b in range(0, len(st), per):
        sl = slice(b, b + per)
        unet.forward_chunks(frame, chunk_size, st[sl], lo[sl], hi[sl], out=out, norm_max=norm_max)
    return out


def predict_chunk_feature_map(input_volume, sl, unet=False, default_only_mask=False, **kwargs):
    """predict.py:100-126: one chunk through the network, returned as numpy (1,5,D,H,W)."""
    assert unet != False, 'Please ensure a unet is loaded and supplied'  # noqa: E712
    sl = sl[1:]
    chunk = input_volume[sl]
    if isinstance(chunk, torch.Tensor):
        tensor = chunk[None, None]
    else:
        tensor = torch.from_numpy(np.ascontiguousarray(chunk, dtype=np.float32)[np.newaxis, np.newaxis])
    predicted_array = unet(tensor).detach().cpu().numpy()
    if default_only_mask:
        predicted_array = predicted_array[3, ...]
    return predicted_array


def process_chunks(input_volume, chunk_size, output_volume, margin, process_data_function,
                   config=None):
    """predict.py:64-96.  With the stock `predict_chunk_feature_map` and an iterseg_b200 UNet
    in `config`, the whole frame is predicted in one batched device pass; any other
    processing function is driven chunk by chunk exactly like the reference."""
    if config is None:
        config = {}
    ndim = len(chunk_size)
    net = config.get('unet')
    if process_data_function is predict_chunk_feature_map and isinstance(net, unet_mod.UNet) \
            and ndim == 3 and not config.get('default_only_mask', False):
        dev = net.device
        vol = input_volume if isinstance(input_volume, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(input_volume, dtype=np.float32))
        frame = vol.to(device=dev, dtype=torch.float32).contiguous()
        if isinstance(output_volume, torch.Tensor) and output_volume.is_cuda:
            predict_frame_device(net, frame, tuple(chunk_size), margin, out=output_volume)
        else:
            feats = predict_frame_device(net, frame, tuple(chunk_size), margin)
            output_volume[...] = feats.cpu().numpy()
        return output_volume
    chunk_starts, chunk_crops = make_chunks(input_volume.shape[-ndim:], chunk_size, margin=margin)
    for start, crop in zip(chunk_starts, chunk_crops):
        sl = (slice(None),) + tuple(slice(int(s0), int(s0) + int(step)) for s0, step in zip(start, chunk_size))
        predicted_array = process_data_function(input_volume, sl, **config)
        p_dim = predicted_array.ndim
        o_dim = output_volume.ndim
        cr = (slice(None),) * (p_dim - o_dim) + tuple(slice(int(i), int(j)) for i, j in crop)
        output_volume[sl][cr] = predicted_array[(0,) + cr]
    return output_volume
