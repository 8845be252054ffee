"""The C-ABI library loads and exports every symbol include/*.h declares; without a GPU it fails
loudly instead of falling back to the CPU.  No compute calls here."""
import ctypes as C
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(kdbgpu_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for required in ("kdbgpu_index_create", "kdbgpu_upload_vectors", "kdbgpu_set_graph", "kdbgpu_search_batch",
                     "kdbgpu_distance_batch", "kdbgpu_flat_search_batch", "kdbgpu_merge_topk_device",
                     "kdbgpu_last_error"):
        assert required in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    from kektordb_b200 import build, ffi
    build.build()
    lib = C.CDLL(ffi.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert set(ffi.SIGNATURES) == set(_declared_symbols())
    assert b"sm_100a" in ffi.lib().kdbgpu_version()


def test_no_cpu_fallback_without_a_device():
    from kektordb_b200 import ffi
    lib = ffi.lib()
    if lib.kdbgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.kdbgpu_index_create(0, 128, ffi.METRIC_COSINE, 16, 1000, C.byref(h))
    assert rc == ffi.ERR_CUDA and not h
    assert b"no CPU fallback" in lib.kdbgpu_last_error()
    from kektordb_b200 import GpuIndex
    with pytest.raises(ffi.GpuError):
        GpuIndex(128, "cosine", 16, 1000)


def test_argument_validation_happens_before_any_device_work():
    from kektordb_b200 import ffi
    lib = ffi.lib()
    h = C.c_void_p()
    assert lib.kdbgpu_index_create(0, 0, ffi.METRIC_COSINE, 16, 1000, C.byref(h)) == ffi.ERR_INVALID
    assert lib.kdbgpu_index_create(0, 128, 7, 16, 1000, C.byref(h)) == ffi.ERR_INVALID
    assert lib.kdbgpu_index_create(0, 128, ffi.METRIC_L2, 16, 0, C.byref(h)) == ffi.ERR_INVALID
    assert lib.kdbgpu_search_batch(None, None, 1, 10, 0, None, 0, None, None, None, None) == ffi.ERR_INVALID
    assert lib.kdbgpu_index_destroy(None) == ffi.OK


def test_product_package_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "kektordb_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
            text = open(path, errors="replace").read()
            assert "kdb_oracle" not in text and "from oracle" not in text and "import oracle" not in text, path


def test_effective_ef_mirror():
    from kektordb_b200 import effective_ef
    from oracle import oracle as O
    for ef in (0, 5, 10, 39, 40, 64, 100, 128, 250):
        for nr in (False, True):
            assert effective_ef(ef, nr) == O.effective_ef(ef, nr)
