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


def test_gemm_args_mirrors_match_the_library():
    """a binding whose struct mirror is shorter than dclip_gemm_args makes dclip_gemm read past it.  The
    library exports sizeof(dclip_gemm_args); the package's mirror and the one printed in INTEGRATION.md must match it."""
    from denseclip_vit_multimodal_b200 import _lib
    lib = _lib.lib()                      # (lib() itself refuses to load on a mismatch)
    assert ctypes.sizeof(_lib.GemmArgs) == lib.dclip_sizeof_gemm_args()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("class GemmArgs(C.Structure)"):]
    block = block[:block.index("]\n") + 1]
    doc_fields = [(n, getattr(ctypes, t)) for n, t in re.findall(r'\("(\w+)", C\.(c_\w+)\)', block)]
    assert doc_fields == list(_lib.GemmArgs._fields_)      # (ctypes aliases c_longlong to c_long on LP64: compare the types)
    hdr = open(os.path.join(ROOT, "include", "denseclip_b200.h")).read()
    body = hdr[hdr.index("typedef struct {"):hdr.index("} dclip_gemm_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.replace("typedef struct {", "").split(";"):
        decl = decl.strip()
        if decl:
            names += [re.sub(r"^.*[\s\*]", "", part.strip()) for part in decl.split(",")]
    assert names == [n for n, _ in _lib.GemmArgs._fields_]
