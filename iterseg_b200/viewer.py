"""Headless stand-ins for the napari objects the segmentation path touches.

The reference reads `layer.data / .scale / .translate / .metadata` and calls
`viewer.add_labels(...)` and `viewer.dims.current_step` (segmentation.py:770-797).
napari is not a dependency of this package: real napari viewers and layers
work (duck typing), and these minimal classes allow the same entry points to
run in scripts, tests and benchmarks.
"""
import numpy as np


class Layer:
    def __init__(self, data, name='layer', scale=None, translate=None, metadata=None):
        self.data = data
        self.name = name
        ndim = getattr(data, 'ndim', len(getattr(data, 'shape', ())))
        self.scale = np.ones(ndim) if scale is None else np.asarray(scale, dtype=float)
        self.translate = np.zeros(ndim) if translate is None else np.asarray(translate, dtype=float)
        self.metadata = {} if metadata is None else metadata


class Image(Layer):
    pass


class Labels(Layer):
    pass


class _Dims:
    def __init__(self):
        self.current_step = (0, 0, 0, 0)


class HeadlessViewer:
    def __init__(self):
        self.layers = {}
        self.dims = _Dims()

    def add_labels(self, data, name='labels', scale=None, translate=None, **kwargs):
        layer = Labels(data, name=name, scale=scale, translate=translate)
        self.layers[name] = layer
        return layer

    def add_image(self, data, name='image', scale=None, translate=None, **kwargs):
        layer = Image(data, name=name, scale=scale, translate=translate)
        self.layers[name] = layer
        return layer
