"""Import the reference's own modules VERBATIM (no source is copied).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable where the reference
checkout exists (the build container: /root/reference); never on the GPU box.
Used by scripts/make_golden.py to mint tests/golden/ and by the `-m "not gpu"`
tests that cross-check the restatement against the real thing when available.

The reference imports napari / magicgui / zarr / dask / toolz / ome_zarr /
tifffile / umetrix / matplotlib / seaborn / tensorstore at module level
(watershed.py:7-10, predict.py:7-10, segmentation.py:2-13) but does not use them
on the hot path; permissive stubs stand in for them.  scikit-image is served by
oracle/skimage_shim.py (parity unpinned for those six functions).
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_SRC = os.environ.get('ISG_REFERENCE_SRC', '/root/reference/src')

_STUB_ROOTS = ('napari', 'magicgui', 'zarr', 'dask', 'toolz', 'ome_zarr',
               'tifffile', 'umetrix', 'matplotlib', 'seaborn', 'tensorstore',
               'qtpy', 'napari_plugin_engine', 'numcodecs')


class _Anything:
    """Callable as decorator and decorator factory, attribute-able, subscriptable."""

    def __init__(self, name='stub'):
        self._name = name

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _Anything):
            return a[0]
        return _Anything(self._name + '()')

    def __getattr__(self, item):
        if item.startswith('__') and item.endswith('__'):
            raise AttributeError(item)
        return _Anything(self._name + '.' + item)

    def __getitem__(self, item):
        return _Anything(self._name + '[]')

    def __iter__(self):
        return iter(())

    def __or__(self, other):
        return self

    __ror__ = __or__

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, item):
        if item.startswith('__') and item.endswith('__'):
            raise AttributeError(item)
        return _Anything(self.__name__ + '.' + item)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def _install_skimage_shim():
    from . import skimage_shim as shim
    if 'skimage' in sys.modules and not getattr(sys.modules['skimage'], '_isg_shim', False):
        return  # a real scikit-image is present: use it
    def mod(name, **attrs):
        m = _StubModule(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    sk = mod('skimage', _isg_shim=True)
    sk.filters = mod('skimage.filters', gaussian=shim.gaussian,
                     threshold_otsu=shim.threshold_otsu)
    sk.feature = mod('skimage.feature', peak_local_max=shim.peak_local_max)
    sk.morphology = mod('skimage.morphology',
                        remove_small_objects=shim.remove_small_objects)
    sk.morphology._util = mod(
        'skimage.morphology._util',
        _offsets_to_raveled_neighbors=shim._offsets_to_raveled_neighbors,
        _validate_connectivity=shim._validate_connectivity)
    for extra in ('segmentation', 'exposure', 'measure', 'metrics', 'io',
                  'util', 'data', 'transform', 'draw'):
        setattr(sk, extra, mod('skimage.' + extra))


_loaded = {}


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'iterseg'))


def load():
    """Return a namespace with the reference modules: unet, watershed, predict,
    segmentation.  Raises RuntimeError when the checkout is absent."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError('reference checkout not found at ' + REFERENCE_SRC)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    _install_skimage_shim()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    # numba cache=True: the reference tree is read-only, keep the cache elsewhere
    os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/isg_numba_cache')
    # iterseg/__init__.py pulls in the widgets; bypass it with a bare package
    pkg = types.ModuleType('iterseg')
    pkg.__path__ = [os.path.join(REFERENCE_SRC, 'iterseg')]
    sys.modules.setdefault('iterseg', pkg)
    for name in ('unet', 'watershed', 'predict', 'segmentation'):
        _loaded[name] = importlib.import_module('iterseg.' + name)
    return types.SimpleNamespace(**_loaded)
