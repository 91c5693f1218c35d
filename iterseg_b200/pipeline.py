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

    def submit(self, frame):
        """Enqueue the U-Net of one (Z,Y,X) float32 device frame; returns its slot index."""
        k = self.n_in % len(self.slots)
        slot = self.slots[k]
        assert not slot['busy'], 'pipeline full: collect() a frame first'
        cur = torch.cuda.current_stream(self.dev)
        self.s_unet.wait_stream(cur)                    # the frame was produced on the caller's stream
        self._reserve(self.n_in > self.n_out)          # a post stage will run beside this U-Net
        with torch.cuda.stream(self.s_unet):
            self.s_unet.wait_event(slot['ev_post'])     # the slot's previous post stage has read feats
            predict.predict_frame_device(self.net, frame, self.chunk_size, self.margin, out=slot['feats'])
            slot['ev_unet'].record(self.s_unet)
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
