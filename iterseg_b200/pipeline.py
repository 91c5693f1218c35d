"""Two frames in flight on one GPU: the post stage (seeds / mask / components / ordered flood) of
frame i runs on one CUDA stream while the U-Net of frame i+1 runs on another.

The ordered flood is latency bound (one warp per object, iterseg_b200/csrc/flood.cuh) and
leaves the tensor cores idle; the U-Net is tensor bound.  Frames of a series are independent
(segmentation.py:873-882), so overlapping the two stages of consecutive frames changes nothing
in the results -- every frame still goes through exactly the same kernels with the same
inputs -- and hides the post stage behind the next frame's U-Net.
"""
import torch

from . import _lib, predict
from . import watershed as ws


class FramePipeline:
    def __init__(self, net, shape, chunk_size, margin, depth=2, post_sms=0, **post_kw):
        self.net, self.shape = net, tuple(int(s) for s in shape)
        self.chunk_size, self.margin = tuple(chunk_size), tuple(margin)
        self.dev = net.device
        self.post_kw = post_kw
        shape_p = tuple(s + 2 for s in self.shape)
        with torch.cuda.device(self.dev):
            self.s_unet, self.s_post = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
            self.slots = [{'feats': torch.zeros((5,) + self.shape, dtype=torch.float32, device=self.dev),
                           'labels': torch.zeros(shape_p, dtype=torch.int32, device=self.dev),
                           'ev_unet': torch.cuda.Event(), 'ev_post': torch.cuda.Event(), 'busy': False}
                          for _ in range(depth)]
        self.n_in = self.n_out = 0
        import os
        self.post_sms = int(os.environ.get('ISG_POST_SMS', post_sms))
        torch.cuda.synchronize(self.dev)

    def _reserve(self, on):
        # optional (off by default; measured neutral on B200, profiles/r01_notes.md): the persistent conv
        # kernels leave `post_sms` SMs to the flood while both are in flight
        _lib.check(_lib.load().isg_set_post_sm_reservation(self.post_sms if on else 0), 'isg_set_post_sm_reservation')

    def submit(self, frame, norm_max=None):
        """Enqueue the U-Net of one (Z,Y,X) float32 device frame; returns its slot index.
        norm_max: 1-element float32 device tensor -> the frame is divided by it inside the first
        kernel (and may then be a pinned HOST tensor, read in place)."""
        k = self.n_in % len(self.slots)
        slot = self.slots[k]
        assert not slot['busy'], 'pipeline full: collect() a frame first'
        cur = torch.cuda.current_stream(self.dev)
        self.s_unet.wait_stream(cur)                    # the frame was produced on the caller's stream
        self._reserve(self.n_in > self.n_out)          # a post stage will run beside this U-Net
        with torch.cuda.stream(self.s_unet):
            self.s_unet.wait_event(slot['ev_post'])     # the slot's previous post stage has read feats
            predict.predict_frame_device(self.net, frame, self.chunk_size, self.margin, out=slot['feats'],
                                         norm_max=norm_max)
            slot['ev_unet'].record(self.s_unet)
        if frame.is_cuda:
            frame.record_stream(self.s_unet)
        slot['busy'] = True
        self.n_in += 1
        return k

    def collect(self):
        """Run the post stage of the oldest submitted frame; returns (labels padded int32 tensor,
        counts int64[8] device tensor).  The tensors stay valid until the slot is submitted again."""
        k = self.n_out % len(self.slots)
        slot = self.slots[k]
        assert slot['busy'], 'nothing submitted'
        self._reserve(self.n_in > self.n_out + 1)      # a U-Net is in flight beside this post stage
        with torch.cuda.stream(self.s_post):
            self.s_post.wait_event(slot['ev_unet'])
            slot['labels'].zero_()
            seeds, counts, mask, otsu = ws.segment_features_device(slot['feats'], slot['labels'], **self.post_kw)
            slot['ev_post'].record(self.s_post)
        slot['busy'] = False
        slot['counts'] = counts
        self.last_post_event = slot['ev_post']      # wait on this to read the labels (not on the next U-Net)
        self.n_out += 1
        self._reserve(False)
        return slot['labels'], counts

    def drain_to(self, stream=None):
        """Make `stream` (default: the current one) wait for everything enqueued so far."""
        stream = stream or torch.cuda.current_stream(self.dev)
        stream.wait_stream(self.s_unet)
        stream.wait_stream(self.s_post)


class DogCore:
    """The submit / collect interface of FramePipeline for the DoG blob segmenter
    (segmentation.py:592-650): the whole frame is one device call, enqueued at submit time on one
    stream; two label slots so that the copy-out of frame i overlaps frame i+1."""

    def __init__(self, shape, dev, **dog_kw):
        self.shape = tuple(int(s) for s in shape)
        self.dev = dev
        self.dog_kw = dog_kw
        shape_p = tuple(s + 2 for s in self.shape)
        with torch.cuda.device(dev):
            self.s_unet = self.s_post = torch.cuda.Stream(dev)
            self.slots = [{'labels': torch.zeros(shape_p, dtype=torch.int32, device=dev), 'counts': None,
                           'ev_post': torch.cuda.Event(), 'busy': False} for _ in range(2)]
        self.n_in = self.n_out = 0

    def submit(self, frame):
        from . import segmentation
        k = self.n_in % 2
        slot = self.slots[k]
        assert not slot['busy'], 'pipeline full: collect() a frame first'
        self.s_post.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.s_post):
            slot['labels'].zero_()
            _, counts = segmentation.dog_blob_segment_device(frame, slot['labels'], **self.dog_kw)
            c8 = torch.zeros(8, dtype=torch.int64, device=self.dev)
            c8[:4] = counts
            slot['ev_post'].record(self.s_post)
        frame.record_stream(self.s_post)
        slot['counts'] = c8
        slot['busy'] = True
        self.n_in += 1
        return k

    def collect(self):
        slot = self.slots[self.n_out % 2]
        assert slot['busy'], 'nothing submitted'
        slot['busy'] = False
        self.last_post_event = slot['ev_post']
        self.n_out += 1
        return slot['labels'], slot['counts']


# Pinned host staging buffers survive a series loop: page-locking 35 MB takes ~10 ms, which is half a
# frame -- a second loop of the same shape (warm restart, the next series of a batch job) finds them here.
_PINNED_POOL = {}
_PINNED_POOL_MAX = 12


def _pinned_take(shape, dtype):
    lst = _PINNED_POOL.get((tuple(shape), dtype))
    if lst:
        return lst.pop()
    return torch.empty(tuple(shape), dtype=dtype).pin_memory()


def _pinned_give(t):
    lst = _PINNED_POOL.setdefault((tuple(t.shape), t.dtype), [])
    if len(lst) < _PINNED_POOL_MAX:
        lst.append(t)


class SeriesPipeline:
    """Host frames in, host labels out, several frames in flight (the frame loop of
    segmentation.py:873-882 without a host synchronisation on the compute streams):

      loader threads   data[t] -> float32 in a pinned ring buffer (or the caller's own array when it
                       is already pinned float32)                       [host, off the main thread]
      s_copy           H2D of frame t+1, its min / max (isg_frame_minmax) and the 8-byte read-back
                       that decides the reference's `min() == 0` branch -- the main thread waits
                       for THIS event only, which never waits for a U-Net
      s_unet           vol /= max (isg_frame_divide_by_max), chunked U-Net of frame t
      s_post           seeds / mask / components / flood of frame t-1, crop of the padded labels
                       (+ the device-resident global label offset) into a contiguous buffer
      s_d2h            one contiguous D2H per frame into a pinned ring buffer (or straight into
                       the caller's array when that is pinned)
      writer threads   wait for the D2H event, store into `output_labels` (numpy / zarr store)

    `core` is a FramePipeline (affinity U-Net watershed) or a DogCore."""

    N_IN = 4        # pinned input buffers / device frame buffers
    N_OUT = 4       # device crop buffers / pinned output buffers (= frames the writer threads may hold)

    def __init__(self, core, normalise=True, staging=None):
        import os
        import queue
        self.core, self.dev, self.shape = core, core.dev, core.shape
        self.normalise = normalise
        # input staging of the affinity path (BASELINE.json north_star (1)):
        #   'copy'     one H2D copy per frame on the copy engine, min/max, divide kernel, U-Net
        #   'fused'    the same copy, the division fused into the first U-Net kernel's loads   (default)
        #   'zerocopy' no copy: min/max and the first U-Net kernel read the pinned frame in place
        # measured on B200 in profiles/r02_notes.md; the copy engine wins (it runs beside the previous
        # frame's kernels, while zero-copy reads put PCIe latency inside the first kernel)
        self.staging = staging or os.environ.get('ISG_INPUT_STAGING', 'fused')
        if not isinstance(core, FramePipeline) or not normalise:
            self.staging = 'copy'
        self.lib = _lib.load()
        n = 1
        for s in self.shape:
            n *= s
        self.nvox = n
        with torch.cuda.device(self.dev):
            self.s_copy, self.s_d2h = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
            self.stage = [{'frame': torch.empty(self.shape, dtype=torch.float32, device=self.dev),
                           'mm': torch.zeros(2, dtype=torch.float32, device=self.dev),
                           'mm_host': torch.zeros(2, dtype=torch.float32).pin_memory(),
                           'scratch': torch.empty(64, dtype=torch.uint8, device=self.dev),
                           'ev_stage': torch.cuda.Event(), 'ev_consumed': torch.cuda.Event()}
                          for _ in range(self.N_IN)]
            self.outs = [{'crop': torch.empty(self.shape, dtype=torch.int32, device=self.dev),
                          'host': None, 'ev_post': torch.cuda.Event(), 'ev_d2h': torch.cuda.Event()}
                         for _ in range(self.N_OUT)]
        # pinned input ring: allocated on first use (page-locking ~35 MB takes milliseconds; a caller
        # whose frames already sit in pinned float32 memory never needs it)
        self.free_in = queue.Queue()
        self._n_in_alloc = 0
        self._in_lock = __import__('threading').Lock()
        self.free_out = queue.Queue()
        for o in range(self.N_OUT):
            self.free_out.put(o)
        self.n_staged = 0
        self._owned_in = []
        self.closed = False
        self.held = __import__('collections').deque()      # zero-copy: pinned buffers of the frames in flight
        torch.cuda.synchronize(self.dev)

    # ---- host side of the input (loader threads) -------------------------------------------
    def load(self, src):
        """-> ('direct', pinned float32 tensor viewing the caller's memory) or ('buf', pinned ring
        buffer holding np.asarray(src).astype(float32))."""
        import numpy as np
        if isinstance(src, np.ndarray) and src.dtype == np.float32 and src.flags.c_contiguous:
            t = torch.from_numpy(src)
            if t.is_pinned():
                return 'direct', t
        buf = None
        with self._in_lock:
            if self.free_in.empty() and self._n_in_alloc < self.N_IN:
                self._n_in_alloc += 1
                buf = _pinned_take(self.shape, torch.float32)
                self._owned_in.append(buf)
        import queue
        while buf is None:                                       # never block for ever: close() wakes us up
            if self.closed:
                raise RuntimeError('series pipeline closed')
            try:
                buf = self.free_in.get(timeout=0.2)
            except queue.Empty:
                pass
        np.copyto(buf.numpy(), np.asarray(src), casting='unsafe')
        return 'buf', buf

    def close(self):
        """Wake loader threads that wait for a buffer.  Call `recycle()` once every thread is done."""
        self.closed = True

    def recycle(self):
        """Hand the pinned staging buffers back to the module-level pool (all copies must have completed)."""
        for buf in self._owned_in:
            _pinned_give(buf)
        self._owned_in = []
        for out in self.outs:
            if out['host'] is not None:
                _pinned_give(out['host'])
                out['host'] = None

    # ---- main thread ---------------------------------------------------------------------------
    def h2d(self, loaded):
        """Enqueue the H2D copy + min / max of one loaded frame on the copy stream."""
        kind, host = loaded
        j = self.n_staged % self.N_IN
        st = self.stage[j]
        with torch.cuda.device(self.dev), torch.cuda.stream(self.s_copy):
            self.s_copy.wait_event(st['ev_consumed'])           # the U-Net that read this buffer last
            if self.staging == 'zerocopy':
                st['src'] = host
            else:
                st['frame'].copy_(host, non_blocking=True)
                st['src'] = st['frame']
            _lib.check(self.lib.isg_frame_minmax(st['src'].data_ptr(), self.nvox, st['mm'].data_ptr(),
                                                 st['scratch'].data_ptr(), st['scratch'].numel(),
                                                 _lib.stream_ptr()), 'isg_frame_minmax')
            st['mm_host'].copy_(st['mm'], non_blocking=True)
            st['ev_stage'].record(self.s_copy)
        st['host'] = (kind, host)
        self.n_staged += 1
        return j

    def submit(self, j):
        """Frame in staging slot j -> compute.  Returns False (nothing enqueued) when the frame
        contains zeros: the caller takes the reference's host route (zero-slice strip)."""
        st = self.stage[j]
        st['ev_stage'].synchronize()                            # copy stream only
        kind, host = st.pop('host')
        zero = self.normalise and float(st['mm_host'][0]) == 0.0
        if self.staging == 'zerocopy' and not zero:
            self.held.append((kind, host, st['ev_consumed']))   # read in place by the U-Net: released in collect()
        elif kind == 'buf':
            self.free_in.put(host)
        if zero:
            return False
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.device(self.dev):
            if self.staging in ('fused', 'zerocopy'):
                with torch.cuda.stream(self.core.s_unet):
                    self.core.s_unet.wait_stream(cur)
                    self.core.submit(st['src'], norm_max=st['mm'][1:2])
            elif self.normalise:
                with torch.cuda.stream(self.core.s_unet):
                    self.core.s_unet.wait_stream(cur)
                    _lib.check(self.lib.isg_frame_divide_by_max(st['frame'].data_ptr(), self.nvox,
                                                                st['mm'].data_ptr(), _lib.stream_ptr()),
                               'isg_frame_divide_by_max')
                with torch.cuda.stream(self.core.s_unet):       # submit() makes s_unet wait for "current"
                    self.core.submit(st['frame'])
            else:
                self.core.submit(st['frame'])
            st['ev_consumed'].record(self.core.s_unet)
        return True

    def collect(self, dst=None, offset_dev=None, on_counts=None):
        """Post stage of the oldest submitted frame, crop, D2H.  `dst`: optional pinned int32
        (Z,Y,X) tensor to copy into directly.  `on_counts(counts)` runs on the post stream right
        after the post stage and may return the device int64 offset tensor for this frame.
        Returns (o, host tensor, counts): wait for self.outs[o]['ev_d2h'], then release(o)."""
        o = self.free_out.get()                                  # back-pressure from the writers
        out = self.outs[o]
        lab, counts = self.core.collect()
        if self.held:                                            # zero-copy: this frame's U-Net has read its input
            kind, host, ev = self.held.popleft()
            ev.synchronize()
            if kind == 'buf':
                self.free_in.put(host)
        Z, Y, X = self.shape
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.core.s_post):
                if on_counts is not None:
                    offset_dev = on_counts(counts)
                self.core.s_post.wait_event(out['ev_d2h'])       # the previous copy out of this buffer
                _lib.check(self.lib.isg_crop_labels(lab.data_ptr(), Z, Y, X, out['crop'].data_ptr(),
                                                    offset_dev.data_ptr() if offset_dev is not None else None,
                                                    _lib.stream_ptr()), 'isg_crop_labels')
                out['ev_post'].record(self.core.s_post)
                if offset_dev is not None:
                    offset_dev.record_stream(self.core.s_post)
            if dst is None:
                if out['host'] is None:
                    out['host'] = _pinned_take(self.shape, torch.int32)
                dst = out['host']
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(out['ev_post'])
                dst.copy_(out['crop'], non_blocking=True)
                out['ev_d2h'].record(self.s_d2h)
        return o, dst, counts

    def release(self, o):
        self.free_out.put(o)
