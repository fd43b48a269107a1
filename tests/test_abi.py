"""CPU: the C-ABI library loads and exports exactly what include/distar_b200.h declares (no compute calls)."""
import os
import re

from distar_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'distar_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dsb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    handle = lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(handle, n), 'missing export ' + n
    assert sorted(lib.EXPORTS) == names, 'ctypes prototypes and header out of sync'
    assert handle.dsb_version() >= 1
    assert handle.dsb_launch_count() == 0
