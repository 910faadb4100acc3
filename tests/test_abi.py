"""The C-ABI library loads and exports every symbol `include/mpqe_b200.h` declares (no compute: runs without a GPU)."""
import ctypes
import os
import re

from mpqe_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mpqe_b200.h')).read()
    return sorted(set(re.findall(r'MPQE_API\s+[\w\s\*]+?\b(mpqe_\w+)\s*\(', text)))


def test_library_builds_and_loads():
    build.build()
    lib = _lib.load()
    assert lib.mpqe_b200_version() >= 100
    assert lib.mpqe_b200_last_error() == b''


def test_every_declared_symbol_is_exported_and_bound():
    build.build()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(raw, name), 'header declares %s but the library does not export it' % name
        assert name in _lib.SIGNATURES, 'no ctypes signature for %s' % name
    for name in _lib.SIGNATURES:
        assert name in names, 'ctypes binds %s which the header does not declare' % name


def test_struct_sizes_match():
    build.build()
    lib = _lib.load()
    for which, struct in enumerate(_lib.ABI_STRUCTS):
        assert lib.mpqe_b200_sizeof(which) == ctypes.sizeof(struct)


def test_argument_errors_are_reported_not_crashing():
    build.build()
    lib = _lib.load()
    assert lib.mpqe_layer_forward(None, 0, 0, None) != 0
    assert b'num_groups' in lib.mpqe_b200_last_error()
    assert lib.mpqe_transpose(None, None, 1, 1, 1, None) != 0
