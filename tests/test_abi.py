"""The C-ABI libraries load without a GPU, export every symbol the headers declare, and fail loudly
(no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import pytest

from dune_sculpt_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix_re):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(%s\w*)\s*\(" % prefix_re, txt))
    return {n for n in names if not n.isupper()}


def test_cuda_library_exports_every_declared_symbol():
    L = capi.cuda_lib()
    declared = _declared("dune_sculpt_cuda.h", "dsc_")
    assert declared == set(capi.CUDA_SYMBOLS), declared ^ set(capi.CUDA_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.dsc_abi_version() == 1


def test_host_library_exports_every_declared_symbol():
    L = capi.host_lib()
    declared = _declared("dune_pbvh.h", "(?:BKE_|DUNE_|SCULPT_|MEM_)")
    declared -= {"MEM_SAFE_FREE", "MEM_mallocN", "MEM_callocN"}
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    assert declared - {"BKE_pbvh_SearchCallback"} <= set(capi.HOST_SYMBOLS) | {"BKE_pbvh_SearchCallback"}


def test_dab_descriptor_layout():
    # DscDab is 4 ints + 20 floats, no padding: the kernels take it by value
    assert C.sizeof(capi.DscDab) == 120
    assert capi.DscDab.radius.offset == 28 and capi.DscDab.grab_delta.offset == 80 and capi.DscDab.radius_scale.offset == 92


def _have_gpu():
    ctx = C.c_void_p()
    r = capi.cuda_lib().dsc_ctx_create(0, C.byref(ctx))
    if r == 0:
        capi.cuda_lib().dsc_ctx_destroy(ctx)
    return r == 0


@pytest.mark.skipif(_have_gpu(), reason="a device is present")
def test_no_device_fails_loudly():
    L = capi.cuda_lib()
    ctx = C.c_void_p()
    r = L.dsc_ctx_create(0, C.byref(ctx))
    assert r == -1 and not ctx  # DSC_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.dsc_last_error(None)
    from dune_sculpt_b200 import meshgen
    with pytest.raises(capi.DeviceError):
        capi.SculptSession(meshgen.grid(9), device=0)


def test_missing_library_message(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_LIB_DIR", str(tmp_path))
    with pytest.raises(capi.NativeLibraryMissing):
        capi._load("libdune_sculpt_cuda.so")
