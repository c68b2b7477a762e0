"""CPU: the C-ABI library loads and exports every symbol include/vpd_b200.h declares."""
import ctypes
import os

from vpd_b200 import _lib, build


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    protos = _lib.parse_header()
    assert len(protos) >= 10
    dll = ctypes.CDLL(path)
    for name in protos:
        assert hasattr(dll, name), name
    dll.vpd_abi_version.restype = ctypes.c_int
    assert dll.vpd_abi_version() >= 1
    dll.vpd_last_error.restype = ctypes.c_char_p
    assert isinstance(dll.vpd_last_error(), bytes)


def test_binding_has_no_fallback(monkeypatch):
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libvpd_b200.so')
    monkeypatch.setattr(_lib, '_lib', None)
    try:
        _lib.lib()
    except _lib.VpdError as e:
        assert 'no CPU or PyTorch fallback' in str(e)
    else:
        raise AssertionError('missing library must raise')
