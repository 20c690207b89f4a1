"""Without a GPU the product must fail loudly, never fall back to a CPU path."""
import ctypes

import pytest


def test_engine_refuses_without_cuda(pkg, make_model):
    lib = pkg.load_library()
    if lib.rwkv_b200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(pkg.EngineError, match="no CUDA device|no CPU fallback"):
        pkg.Engine(make_model(1, 64))
    h = ctypes.c_void_p()
    rc = lib.rwkv_b200_load(make_model(1, 64).encode(), 1, 0, 1, ctypes.byref(h), None, None)
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.rwkv_b200_last_error()


def test_host_alloc_works_without_driver(pkg):
    lib = pkg.load_library()
    p = lib.rwkv_b200_host_alloc(1 << 16)
    assert p
    ctypes.memset(p, 0xAB, 1 << 16)
    lib.rwkv_b200_host_free(p)
