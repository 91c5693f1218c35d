"""The bundled 3-D anisotropic U-Net, executed by hand-written sm_100a kernels.

Mirrors `UNet(in_channels=1, out_channels=5)` of src/iterseg/unet.py:126-364:
the module holds parameters under the SAME state_dict keys (so the reference's
network files -- `torch.save(unet.state_dict(), path)`, train.py:414-420 -- load
unchanged, predict.py:34), and `unet(tensor)` returns what the reference's
forward returns.  The arithmetic does not go through torch: `forward` hands
the packed weights and the chunk(s) to the C-ABI (isg_unet_forward_chunks):
conv3d = implicit GEMM on tcgen05 tensor cores (fp16 operands, fp32
accumulate), BatchNorm with per-chunk batch statistics (the reference never
calls .eval(), predict.py:25-35), max-pools, depthwise transposed convolutions,
crops/concats and the sigmoid head fused into memory-bound companion kernels.

Differences that do not change outputs: BatchNorm running statistics are not
updated (the reference mutates them as a side effect of train mode), no
autograd graph is built.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib

ENCODER = [('c0', 1, 32), ('c1', 32, 64), ('c2', 64, 128), ('c3', 128, 256), ('c4', 256, 256)]
DECODER = [('c5_0', 512, 128), ('c6_0', 256, 64), ('c7_0', 128, 32), ('c8_0', 64, 5)]
UPS = [('up0', 256, (2, 2, 2)), ('up1', 128, (1, 2, 2)), ('up2', 64, (1, 2, 2)), ('up3', 32, (1, 2, 2))]


class ConvModule(nn.Module):
    """Parameter container with the reference's names (unet.py:25-88)."""

    def __init__(self, in_channels, out_channels, final='relu'):
        super().__init__()
        self.conv0 = nn.Conv3d(in_channels, out_channels, 3, padding=1)
        self.conv1 = nn.Conv3d(out_channels, out_channels, 3, padding=1)
        self.batch0 = nn.BatchNorm3d(out_channels)
        self.batch1 = nn.BatchNorm3d(out_channels)
        self.final = final


class _Plan:
    """isg_unet_plan: the geometry (frame extent, chunk extent, chunk count) over the U-Net's
    shared workspace.  The chunk tables are per-call data (`set_chunks`)."""

    def __init__(self, unet, frame_shape, chunk_shape, n_chunks, workspace):
        lib = _lib.load()
        self.n = int(n_chunks)
        self.frame_shape = tuple(int(s) for s in frame_shape)
        self.chunk_shape = tuple(int(s) for s in chunk_shape)
        cz, cy, cx = self.chunk_shape
        self.workspace = workspace                     # keeps the tensor alive as long as the plan
        self.packed = unet.packed_weights()
        with torch.cuda.device(unet.device):
            self.ptr = lib.isg_unet_plan_create(
                self.packed.data_ptr(), self.n, cz, cy, cx, *self.frame_shape, None, None, None,
                self.workspace.data_ptr(), self.workspace.numel())
        if not self.ptr:
            raise _lib.IsgError('isg_unet_plan_create: ' + lib.isg_last_error().decode())
        self.flops = lib.isg_unet_plan_flops(self.ptr)

    def set_chunks(self, starts, crop_lo, crop_hi):
        st = np.ascontiguousarray(starts, dtype=np.int32).reshape(-1, 3)
        lo = np.ascontiguousarray(crop_lo, dtype=np.int32).reshape(-1, 3)
        hi = np.ascontiguousarray(crop_hi, dtype=np.int32).reshape(-1, 3)
        if not (len(st) == len(lo) == len(hi) == self.n):
            raise ValueError(f'plan of {self.n} chunks got tables of {len(st)}, {len(lo)}, {len(hi)} rows')
        _lib.check(_lib.load().isg_unet_plan_set_chunks(self.ptr, st.ctypes.data, lo.ctypes.data, hi.ctypes.data),
                   'isg_unet_plan_set_chunks')

    def __del__(self):
        try:
            if getattr(self, 'ptr', None):
                _lib.load().isg_unet_plan_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def workspace_bytes(n_chunks, chunk_shape):
    """Activation workspace of one forward_chunks call (0: the chunk shape is not valid)."""
    cz, cy, cx = (int(c) for c in chunk_shape)
    return int(_lib.load().isg_unet_workspace_bytes(int(n_chunks), cz, cy, cx))


class UNet(nn.Module):

    def __init__(self, in_channels=1, out_channels=5, down_factors=(1, 2, 2), up='convolution',
                 downsample_1_at_bottom=True, chan_final_activations=None):
        super().__init__()
        if (in_channels, out_channels, tuple(down_factors), up, downsample_1_at_bottom,
                chan_final_activations) != (1, 5, (1, 2, 2), 'convolution', True, None):
            raise NotImplementedError(
                'iterseg_b200 implements the bundled architecture UNet(in_channels=1, out_channels=5) '
                'of the affinity-unet-watershed path only')
        self.out_channels = (out_channels,)
        self.forked = False
        for name, cin, cout in ENCODER:
            setattr(self, name, ConvModule(cin, cout))
        for name, cin, cout in DECODER:
            setattr(self, name, ConvModule(cin, cout, final='sigmoid' if name == 'c8_0' else 'relu'))
        for name, c, k in UPS:
            setattr(self, name, nn.ConvTranspose3d(c, c, kernel_size=k, stride=k, groups=c))
        self._packed = None
        self._plans = {}              # (frame shape, chunk shape, n chunks) -> _Plan, least recently used first
        self._workspace = None        # ONE activation workspace shared by all plans
        self._ws_event = None         # end of the last forward pass that used the workspace

    # ---- weights -------------------------------------------------------------------
    @property
    def device(self):
        return self.c0.conv0.weight.device

    def _invalidate(self):
        self._packed = None
        self._plans = {}
        self._workspace = None
        self._ws_event = None

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._invalidate()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._invalidate()
        return r

    def packed_weights(self):
        """fp16 tap-major conv weights + BN affine + transposed-conv weights in one device blob."""
        if self._packed is not None:
            return self._packed
        _lib.require_device()
        if self.device.type != 'cuda':
            raise _lib.IsgError('the U-Net must live on a CUDA device (call .cuda()); there is no CPU path')
        lib = _lib.load()
        sd = self.state_dict()
        tensors = []
        keep = []
        for key, t in sd.items():
            if t.dtype.is_floating_point:
                t = t.detach().to(torch.float32).contiguous()
                keep.append(t)
                tensors.append(t.data_ptr())
            else:
                tensors.append(0)
        assert len(tensors) == 134, len(tensors)
        arr = (ctypes.c_void_p * len(tensors))(*tensors)
        packed = torch.empty(int(lib.isg_unet_packed_weight_bytes()), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.isg_unet_weights_pack(arr, len(tensors), packed.data_ptr(), _lib.stream_ptr())
            _lib.check(rc, 'isg_unet_weights_pack')
            torch.cuda.current_stream().synchronize()
        self._packed = packed
        return packed

    # ---- execution -----------------------------------------------------------------
    MAX_PLANS = 8

    def plan(self, frame_shape, chunk_shape, n_chunks):
        """The plan for `n_chunks` chunks of `chunk_shape` cut from a frame of `frame_shape`.
        All plans of a network share one workspace, sized for the largest request so far; when it
        has to grow, the plans that point into the old block are dropped (the old block stays
        alive until the kernels already enqueued on the current stream are done: torch's caching
        allocator only hands it out again in stream order, and forward_chunks records every
        stream that used it)."""
        key = (tuple(int(v) for v in frame_shape), tuple(int(v) for v in chunk_shape), int(n_chunks))
        p = self._plans.pop(key, None)
        if p is None:
            need = workspace_bytes(n_chunks, chunk_shape)
            if need == 0:
                raise ValueError(
                    f'chunk shape {key[1]} is not valid for this U-Net (z must be even, y/x '
                    'must survive four poolings and the decoder crops, e.g. (10, 256, 256))')
            if self._workspace is None or self._workspace.numel() < need:
                held = self._workspace.numel() if self._workspace is not None else 0
                free, _ = torch.cuda.mem_get_info(self.device)
                if need > free + held:
                    torch.cuda.empty_cache()                       # blocks torch caches but does not use
                    free, _ = torch.cuda.mem_get_info(self.device)
                if need > free + held:
                    raise _lib.IsgError(
                        f'the U-Net workspace for {n_chunks} chunks of {key[1]} needs {need / 2**30:.1f} GiB, '
                        f'{(free + held) / 2**30:.1f} GiB are free on {self.device}: use fewer chunks per batch')
                self._plans.clear()
                self._workspace = None
                self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
            while len(self._plans) >= self.MAX_PLANS:
                self._plans.pop(next(iter(self._plans)))          # least recently used
            p = _Plan(self, key[0], key[1], key[2], self._workspace)
        self._plans[key] = p                                       # most recently used last
        return p

    def forward_chunks(self, frame, chunk_shape, starts, crop_lo, crop_hi, out=None, norm_max=None):
        """frame (Z,Y,X) float32 CUDA tensor -> (5,Z,Y,X) float32: the cropped interior of
        every chunk's prediction is placed into `out` (process_chunks, predict.py:64-96).
        norm_max: optional 1-element float32 CUDA tensor; the frame is divided by it inside the first
        kernel (vol /= max, segmentation.py:889) and may then be a PINNED HOST tensor read in place
        ("zero-copy" staging; `out` must be given)."""
        assert frame.dtype == torch.float32 and frame.is_contiguous()
        assert frame.is_cuda or (norm_max is not None and frame.is_pinned() and out is not None)
        p = self.plan(frame.shape, chunk_shape, len(starts))
        p.set_chunks(starts, crop_lo, crop_hi)
        if out is None:
            out = torch.zeros((5,) + tuple(frame.shape), dtype=torch.float32, device=frame.device)
        with torch.cuda.device(out.device):
            cur = torch.cuda.current_stream()
            # forward passes share the workspace: whatever stream this one is enqueued on, it starts
            # after the previous one has finished (a device-side wait, nothing blocks on the host)
            if self._ws_event is not None:
                cur.wait_event(self._ws_event)
            p.workspace.record_stream(cur)
            if norm_max is None:
                rc = _lib.load().isg_unet_forward_chunks(p.ptr, frame.data_ptr(), out.data_ptr(),
                                                         _lib.stream_ptr())
            else:
                rc = _lib.load().isg_unet_forward_chunks_norm(p.ptr, frame.data_ptr(), norm_max.data_ptr(),
                                                              out.data_ptr(), _lib.stream_ptr())
            _lib.check(rc, 'isg_unet_forward_chunks')
            if self._ws_event is None:
                self._ws_event = torch.cuda.Event()
            self._ws_event.record(cur)
        return out

    def check_overflow(self):
        """Raise IsgError (status ISG_ERR_OVERFLOW) if a COMPLETED forward pass left the fp16 range
        of the pre-BatchNorm activations (include/iterseg_b200.h, "fp16 range guard").  Does not
        synchronise: call it after waiting for the features or the labels."""
        lib = _lib.load()
        for p in self._plans.values():
            if lib.isg_unet_plan_overflowed(p.ptr):
                with torch.cuda.device(self.device):
                    lib.isg_unet_plan_clear_overflow(p.ptr, _lib.stream_ptr())
                err = _lib.IsgError(
                    'the U-Net overflowed the fp16 range of its pre-BatchNorm activations: the feature volume of '
                    'that forward pass is invalid (filters are prescaled per output channel, so this means '
                    'BatchNorm / transposed-convolution parameters of extreme magnitude)')
                err.status = _lib.ISG_ERR_OVERFLOW
                raise err

    def forward(self, x):
        """x: (1,1,D,H,W) -> (1,5,D,H,W) float32 on the device, like the reference's forward
        (one chunk == one batch-norm batch)."""
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f'expected input of shape (1, 1, D, H, W); got {tuple(x.shape)}')
        if x.shape[0] != 1:
            raise NotImplementedError('batch size > 1 would mix BatchNorm statistics across chunks; '
                                      'use forward_chunks for many chunks')
        frame = x[0, 0].to(device=self.device, dtype=torch.float32).contiguous()
        shape = tuple(frame.shape)
        zeros = np.zeros((1, 3), np.int32)
        out = self.forward_chunks(frame, shape, zeros, zeros, np.asarray([shape], np.int32))
        torch.cuda.current_stream(self.device).synchronize()
        self.check_overflow()
        return out[None]

    def debug_conv_output(self, frame, chunk_shape, starts, crop_lo, crop_hi, name, chunk=0):
        """Raw (pre-BatchNorm, bias-free) output of convolution `name` for one chunk as
        (Cout, D, H, W) float32 -- parity tests only."""
        lib = _lib.load()
        p = self.plan(frame.shape, chunk_shape, len(starts))
        p.set_chunks(starts, crop_lo, crop_hi)
        mod, conv = name.split('.')
        cout = getattr(getattr(self, mod), conv).out_channels
        level = {'c0': 0, 'c1': 1, 'c2': 2, 'c3': 3, 'c4': 4, 'c5_0': 3, 'c6_0': 2, 'c7_0': 1, 'c8_0': 0}[mod]
        d, h, w = chunk_shape
        for l in range(level):
            h, w = h // 2 + 1, w // 2 + 1
        if level == 4:
            d //= 2
        out = torch.empty((cout, d, h, w), dtype=torch.float32, device=frame.device)
        with torch.cuda.device(frame.device):
            rc = lib.isg_unet_debug_activation(p.ptr, frame.data_ptr(), name.encode(), chunk,
                                               out.data_ptr(), out.numel(), _lib.stream_ptr())
        _lib.check(rc, 'isg_unet_debug_activation')
        return out
