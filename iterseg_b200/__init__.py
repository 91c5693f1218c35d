"""iterseg_b200 -- the affinity U-Net watershed `segment_data` path of
AbigailMcGovern/iterseg, rebuilt for NVIDIA B200 (sm_100a).

The Python modules mirror the reference's names (predict, unet, watershed,
segmentation, _dock_widgets, _io); the arithmetic lives in
libiterseg_b200.so (hand-written CUDA behind the C-ABI of
include/iterseg_b200.h).  There is no CPU fallback.
"""
__version__ = '0.1.0'
