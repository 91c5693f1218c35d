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
