"""The C-ABI library loads and exports every symbol include/iterseg_b200.h declares
(no compute calls: CPU only)."""
import ctypes
import os
import re

from iterseg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'iterseg_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(isg_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_version_and_error_text():
    lib = _lib.load()
    assert lib.isg_version() >= 100
    assert isinstance(lib.isg_last_error(), bytes)
    # workspace queries are pure host arithmetic
    assert lib.isg_post_workspace_bytes(33, 512, 512, 1 << 20) > 35 * 514 * 514 * 12
    assert lib.isg_flood_workspace_bytes(35, 514, 514, 1000) > 35 * 514 * 514 * 13
    assert lib.isg_post_workspace_bytes(0, 1, 1, 1) == 0


def test_argument_validation_returns_status_codes():
    """Bad arguments are refused with a status code and an error text BEFORE anything touches the
    device (so this runs on a CPU-only box): the reference's Python exceptions are raised from
    these codes by the mirror modules."""
    lib = _lib.load()
    n = None
    assert lib.isg_affinity_flood(n, 0, 0, n, n, n, 0, n, 5, 5, 5, n, n, 0, n) != 0
    assert b'null pointer' in lib.isg_last_error()
    p = _lib.PostParams()
    assert lib.isg_segment_features(n, 5, 4, 4, 4, ctypes.byref(p), n, n, n, n, n, 1, n, n, n, 0, n) != 0
    assert lib.isg_slab_stats(n, 5, 4, 4, 4, ctypes.byref(p), n, 0, n, n, n, n, 0, n) != 0
    d = _lib.DogParams()
    assert lib.isg_dog_blob_segment(n, 4, 4, 4, ctypes.byref(d), n, n, n, 1, n, n, 0, n) != 0
    assert lib.isg_label_metrics(n, n, 10, 3, 0.5, n, n, 0, n) != 0
    assert lib.isg_relabel_by_keys(n, 0, n, 0, n, 0, n, n, n) != 0
    assert lib.isg_sort_keys_u64(n, 5, n, 0, n) != 0
    assert lib.isg_unet_plan_create(n, 1, 10, 256, 256, 33, 512, 512, n, n, n, n, 0) is None
    # size queries: host arithmetic, monotone in the volume
    assert lib.isg_dog_workspace_bytes(33, 512, 512, 1 << 16) > lib.isg_dog_workspace_bytes(10, 128, 128, 1 << 16) > 0
    assert lib.isg_metrics_workspace_bytes(1000, 10) > 0 and lib.isg_metrics_workspace_bytes(0, 10) == 0
    assert lib.isg_sort_tmp_bytes(1000) > 8000
    assert lib.isg_unet_workspace_bytes(36, 10, 256, 256) > 10 * (1 << 30)
