"""`segment_data`: the widget-level entry point of the path.

Mirrors src/iterseg/_dock_widgets.py:537-612.  In the reference this is a
magicgui factory; magicgui is optional here: when it is importable the same
decorator is applied (so napari can dock it), otherwise `segment_data` is the
plain function with the same signature and defaults.
"""
from typing import Union

from .segmentation import segmenters


def _segment_data(napari_viewer, input_volume_layer, save_dir: Union[str, None] = None,
                  name: str = 'labels-prediction', segmenter: str = 'affinity-unet-watershed',
                  network_or_config_file: Union[str, None] = None,
                  layer_reference: Union[str, None] = None, chunk_size: tuple = (10, 256, 256),
                  margin: tuple = (1, 64, 64), debug: bool = True):
    """Segment an image (3-D zyx or 4-D tzyx layer) with one of the registered segmenters."""
    segment_func = segmenters[segmenter]
    return segment_func(napari_viewer, input_volume_layer, save_dir, name, network_or_config_file,
                        layer_reference, chunk_size, margin, debug)


try:                                               # pragma: no cover - GUI only
    from magicgui import magic_factory

    segment_data = magic_factory(
        _segment_data, call_button='Segment',
        segmenter={'widget_type': 'ComboBox', 'choices': list(segmenters.keys())},
        save_dir={'widget_type': 'FileEdit', 'mode': 'd'})
except Exception:
    segment_data = _segment_data
