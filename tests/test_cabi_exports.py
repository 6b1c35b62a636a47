"""CPU: the C-ABI library builds for sm_100a, loads without a GPU/driver, and exports every symbol include/*.h declares.
(No compute call is made here.)"""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "denseclip_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dclip_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    from denseclip_vit_multimodal_b200 import _lib
    path = _lib.build()          # no-op when up to date; nvcc cross-compiles without a GPU
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == syms   # the ctypes binding covers the whole header
    lib.dclip_abi_version.restype = ctypes.c_int
    assert lib.dclip_abi_version() == 1


def test_no_driver_dependency_and_loud_failure_without_gpu():
    import subprocess
    from denseclip_vit_multimodal_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out   # plain C ABI, driver entry points resolved at run time
    import torch
    if not torch.cuda.is_available():
        lib = _lib.lib()
        h = ctypes.c_void_p()
        assert lib.dclip_create(0, ctypes.byref(h)) != 0        # fails loudly: no CPU fallback
        assert b"CUDA" in lib.dclip_last_error(None) or b"device" in lib.dclip_last_error(None)


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import shutil
    import subprocess
    from denseclip_vit_multimodal_b200 import _lib
    if not shutil.which("cuobjdump"):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):   # tcgen05.mma, TMA load, tcgen05.ld / st
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")           # no legacy mma.sync tensor path
