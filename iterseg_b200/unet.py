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
    """isg_unet_plan + the workspace it points into."""

    def __init__(self, unet, frame_shape, chunk_shape, starts, crop_lo, crop_hi):
        lib = _lib.load()
        self.n = len(starts)
        self.frame_shape = tuple(int(s) for s in frame_shape)
        self.chunk_shape = tuple(int(s) for s in chunk_shape)
        cz, cy, cx = self.chunk_shape
        nbytes = lib.isg_unet_workspace_bytes(self.n, cz, cy, cx)
        if nbytes == 0:
            raise ValueError(
                f'chunk shape {self.chunk_shape} is not valid for this U-Net (z must be even, y/x '
                'must survive four poolings and the decoder crops, e.g. (10, 256, 256))')
        dev = unet.device
        self.workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        st = np.ascontiguousarray(starts, dtype=np.int32).reshape(-1, 3)
        lo = np.ascontiguousarray(crop_lo, dtype=np.int32).reshape(-1, 3)
        hi = np.ascontiguousarray(crop_hi, dtype=np.int32).reshape(-1, 3)
        self.packed = unet.packed_weights()
        with torch.cuda.device(dev):
            self.ptr = lib.isg_unet_plan_create(
                self.packed.data_ptr(), self.n, cz, cy, cx, *self.frame_shape,
                st.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                self.workspace.data_ptr(), self.workspace.numel())
        if not self.ptr:
            raise _lib.IsgError('isg_unet_plan_create: ' + lib.isg_last_error().decode())
        self.flops = lib.isg_unet_plan_flops(self.ptr)

    def __del__(self):
        try:
            if getattr(self, 'ptr', None):
                _lib.load().isg_unet_plan_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


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
        self._plans = {}

    # ---- weights -------------------------------------------------------------------
    @property
    def device(self):
        return self.c0.conv0.weight.device

    def _invalidate(self):
        self._packed = None
        self._plans = {}

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
    def plan(self, frame_shape, chunk_shape, starts, crop_lo, crop_hi):
        key = (tuple(frame_shape), tuple(chunk_shape), np.asarray(starts).tobytes(),
               np.asarray(crop_lo).tobytes(), np.asarray(crop_hi).tobytes())
        p = self._plans.get(key)
        if p is None:
            if len(self._plans) >= 4:
                self._plans.clear()
            p = _Plan(self, frame_shape, chunk_shape, starts, crop_lo, crop_hi)
            self._plans[key] = p
        return p

    def forward_chunks(self, frame, chunk_shape, starts, crop_lo, crop_hi, out=None):
        """frame (Z,Y,X) float32 CUDA tensor -> (5,Z,Y,X) float32: the cropped interior of
        every chunk's prediction is placed into `out` (process_chunks, predict.py:64-96)."""
        assert frame.is_cuda and frame.dtype == torch.float32 and frame.is_contiguous()
        p = self.plan(frame.shape, chunk_shape, starts, crop_lo, crop_hi)
        if out is None:
            out = torch.zeros((5,) + tuple(frame.shape), dtype=torch.float32, device=frame.device)
        with torch.cuda.device(frame.device):
            rc = _lib.load().isg_unet_forward_chunks(p.ptr, frame.data_ptr(), out.data_ptr(),
                                                     _lib.stream_ptr())
        _lib.check(rc, 'isg_unet_forward_chunks')
        return out

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
        return out[None]

    def debug_conv_output(self, frame, chunk_shape, starts, crop_lo, crop_hi, name, chunk=0):
        """Raw (pre-BatchNorm, bias-free) output of convolution `name` for one chunk as
        (Cout, D, H, W) float32 -- parity tests only."""
        lib = _lib.load()
        p = self.plan(frame.shape, chunk_shape, starts, crop_lo, crop_hi)
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
