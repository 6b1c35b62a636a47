"""ctypes binding of libdenseclip_b200.so (C ABI declared in include/denseclip_b200.h).

The shared library is built in-tree by ``build()`` (nvcc, sm_100a only) and loaded lazily.  There is no CPU or
PyTorch fallback: if the library is missing, or the device is not a Blackwell (sm_100) GPU, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libdenseclip_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"]


class DclipError(RuntimeError):
    pass


def _sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")) and not f.startswith("selftest")
    ) + [os.path.join(INCLUDE, "denseclip_b200.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/dclip_api.cu into libdenseclip_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise DclipError("nvcc not found: cannot build libdenseclip_b200.so")
    tmp = LIB_PATH + ".tmp%d" % os.getpid()
    cmd = [nvcc, *NVCC_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", tmp, os.path.join(CSRC, "dclip_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise DclipError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print("built", LIB_PATH)
    return LIB_PATH


# ---- struct mirrors (keep in sync with include/denseclip_b200.h) ----------------------------------------------
class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_longlong),
        ("W", C.c_void_p), ("ldw", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("split_in", C.c_int),
        ("bias", C.c_void_p),
        ("act", C.c_int),
        ("out_scale", C.c_float),
        ("residual", C.c_void_p), ("ldr", C.c_longlong),
        ("res_mod", C.c_int),
        ("remap_P", C.c_int), ("remap_Nt", C.c_int),
        ("out_f32", C.c_void_p), ("ldc", C.c_longlong),
        ("out_bf16", C.c_void_p), ("ldcb", C.c_longlong),
        ("split_out", C.c_int), ("split_out_off", C.c_longlong),
        ("block_n", C.c_int),
        ("conv_C", C.c_int), ("conv_gw", C.c_int), ("conv_gh", C.c_int), ("conv_B", C.c_int),
        ("a_bs", C.c_longlong),
        ("conv_G", C.c_int), ("a_gs", C.c_longlong),
        ("wg_C", C.c_int), ("wg_pitch", C.c_int), ("wg_grouped", C.c_int), ("wg_rows", C.c_int),
    ]


class ColReduceArgs(C.Structure):   # mirrors dclip_col_reduce_args
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_longlong), ("x", C.c_void_p), ("ldx", C.c_longlong),
        ("mean", C.c_void_p), ("rstd", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("mask", C.c_void_p), ("ldm", C.c_longlong), ("mask_scale", C.c_float),
        ("relu", C.c_int), ("M", C.c_int), ("N", C.c_int), ("mode", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("out0", C.c_void_p), ("out1", C.c_void_p), ("out2", C.c_void_p), ("eps", C.c_float),
        ("run_mean", C.c_void_p), ("run_var", C.c_void_p), ("momentum", C.c_float),
    ]


class BnApplyArgs(C.Structure):   # mirrors dclip_bn_apply_args
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_longlong), ("x", C.c_void_p), ("ldx", C.c_longlong),
        ("mean", C.c_void_p), ("rstd", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("sum_g", C.c_void_p), ("sum_gx", C.c_void_p),
        ("mask", C.c_void_p), ("ldm", C.c_longlong), ("mask_scale", C.c_float),
        ("relu", C.c_int), ("M", C.c_int), ("N", C.c_int), ("mode", C.c_int),
        ("out_f32", C.c_void_p), ("ldo", C.c_longlong), ("out_bf16", C.c_void_p), ("ldb", C.c_longlong),
    ]


class VitConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("layers", C.c_int), ("heads", C.c_int), ("patch_size", C.c_int), ("grid0", C.c_int),
                ("precise", C.c_int), ("ln_fold", C.c_int)]


_PP = C.POINTER(C.c_void_p)


class VitWeights(C.Structure):
    _fields_ = [
        ("conv1_w", C.c_void_p), ("class_embedding", C.c_void_p), ("positional_embedding", C.c_void_p),
        ("ln_pre_g", C.c_void_p), ("ln_pre_b", C.c_void_p), ("ln_post_g", C.c_void_p), ("ln_post_b", C.c_void_p),
        ("ln1_g", _PP), ("ln1_b", _PP), ("ln2_g", _PP), ("ln2_b", _PP),
        ("in_proj_w", _PP), ("in_proj_b", _PP), ("out_proj_w", _PP), ("out_proj_b", _PP),
        ("fc_w", _PP), ("fc_b", _PP), ("proj_w", _PP), ("proj_b", _PP),
        ("ln1_c", _PP), ("ln2_c", _PP),
    ]


class VitOutputs(C.Structure):
    _fields_ = [("n_taps", C.c_int), ("tap_layers", C.POINTER(C.c_int)), ("taps_nchw", _PP), ("taps_tokens_bf16", _PP),
                ("last_tokens_f32", C.c_void_p)]


EXPORTS = [
    "dclip_abi_version", "dclip_create", "dclip_destroy", "dclip_last_error", "dclip_launch_count",
    "dclip_reset_launch_count", "dclip_gemm", "dclip_sizeof_gemm_args", "dclip_layernorm", "dclip_cast_bf16", "dclip_attention",
    "dclip_attention_split",
    "dclip_attention_small", "dclip_im2col_patches", "dclip_posemb_interp", "dclip_tap_nchw", "dclip_nchw_to_tokens",
    "dclip_token_mean", "dclip_score_map", "dclip_upsample_bilinear", "dclip_upsample_argmax", "dclip_eval_stats", "dclip_gamma_residual", "dclip_conv3x3_gather",
    "dclip_col_reduce_workspace", "dclip_col_reduce", "dclip_bn_apply", "dclip_transpose_pad", "dclip_upsample_bilinear_bwd",
    "dclip_loss_workspace", "dclip_ce_loss", "dclip_ce_loss_bwd", "dclip_silog_loss", "dclip_silog_loss_bwd",
    "dclip_vit_create", "dclip_vit_destroy", "dclip_vit_set_weights", "dclip_vit_workspace_bytes", "dclip_vit_forward",
]

_lib = None
_lock = threading.RLock()
_handles = {}


def _declare(lib):
    vp, ll, i, f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    lib.dclip_abi_version.restype = i
    lib.dclip_create.argtypes = [i, C.POINTER(vp)]
    lib.dclip_destroy.argtypes = [vp]
    lib.dclip_last_error.argtypes = [vp]
    lib.dclip_last_error.restype = C.c_char_p
    lib.dclip_launch_count.argtypes = [vp]
    lib.dclip_launch_count.restype = ll
    lib.dclip_reset_launch_count.argtypes = [vp]
    lib.dclip_gemm.argtypes = [vp, C.POINTER(GemmArgs), vp]
    lib.dclip_layernorm.argtypes = [vp, vp, ll, vp, vp, f, i, i, vp, ll, vp, ll, i, ll, vp]
    lib.dclip_cast_bf16.argtypes = [vp, vp, ll, vp, ll, i, i, i, ll, f, vp]
    lib.dclip_attention.argtypes = [vp, vp, vp, vp, ll, ll, ll, ll, ll, ll, i, i, i, i, i, i, i, i, f, vp, ll, ll, vp]
    lib.dclip_attention_split.argtypes = [vp, vp, vp, vp, ll, ll, ll, ll, ll, ll, i, i, i, ll, i, i, i, i, f, vp, ll, ll, ll, vp]
    lib.dclip_sizeof_gemm_args.argtypes = []
    lib.dclip_attention_small.argtypes = [vp, vp, vp, vp, i, ll, ll, ll, ll, ll, ll, i, i, i, i, i, i, i, i, f, i, vp, i, ll, ll,
                                          ll, vp]
    lib.dclip_im2col_patches.argtypes = [vp, vp, i, i, i, i, vp, ll, i, ll, vp]
    lib.dclip_posemb_interp.argtypes = [vp, vp, i, i, i, i, vp, vp]
    lib.dclip_tap_nchw.argtypes = [vp, vp, i, i, i, vp, vp]
    lib.dclip_nchw_to_tokens.argtypes = [vp, vp, i, i, i, vp, vp, ll, ll, i, vp]
    lib.dclip_token_mean.argtypes = [vp, vp, i, i, i, ll, ll, i, vp, vp]
    lib.dclip_score_map.argtypes = [vp, vp, ll, ll, i, vp, i, i, i, i, f, vp, vp]
    lib.dclip_upsample_bilinear.argtypes = [vp, vp, i, ll, ll, i, i, i, i, i, i, vp, vp]
    lib.dclip_upsample_argmax.argtypes = [vp, vp, ll, ll, i, i, i, i, i, i, vp, vp]
    lib.dclip_eval_stats.argtypes = [vp, vp, vp, i, ll, i, i, vp, vp, vp, ll, vp, vp, vp]
    lib.dclip_gamma_residual.argtypes = [vp, vp, vp, vp, vp, ll, i, vp]
    lib.dclip_conv3x3_gather.argtypes = [vp, vp, i, ll, ll, i, i, i, i, i, vp, ll, vp]
    lib.dclip_col_reduce_workspace.argtypes = [i, i]
    lib.dclip_col_reduce.argtypes = [vp, C.POINTER(ColReduceArgs), vp]
    lib.dclip_bn_apply.argtypes = [vp, C.POINTER(BnApplyArgs), vp]
    lib.dclip_transpose_pad.argtypes = [vp, vp, i, ll, ll, i, i, i, i, i, i, i, i, i, ll, vp, ll, vp]
    lib.dclip_upsample_bilinear_bwd.argtypes = [vp, vp, i, i, i, i, i, i, vp, ll, vp]
    lib.dclip_loss_workspace.argtypes = []
    lib.dclip_ce_loss.argtypes = [vp, vp, vp, i, i, ll, i, vp, vp, vp]
    lib.dclip_ce_loss_bwd.argtypes = [vp, vp, vp, i, i, ll, i, vp, vp, vp, vp]
    lib.dclip_silog_loss.argtypes = [vp, vp, vp, vp, ll, f, f, vp, vp, vp]
    lib.dclip_silog_loss_bwd.argtypes = [vp, vp, vp, vp, ll, f, f, vp, vp, vp, vp]
    lib.dclip_vit_create.argtypes = [vp, C.POINTER(VitConfig), C.POINTER(vp)]
    lib.dclip_vit_destroy.argtypes = [vp]
    lib.dclip_vit_set_weights.argtypes = [vp, C.POINTER(VitWeights)]
    lib.dclip_vit_workspace_bytes.argtypes = [vp, i, i, i, C.POINTER(C.c_size_t)]
    lib.dclip_vit_forward.argtypes = [vp, vp, i, i, i, vp, C.c_size_t, C.POINTER(VitOutputs), vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("dclip_last_error", "dclip_launch_count", "dclip_sizeof_gemm_args", "dclip_col_reduce_workspace",
                        "dclip_loss_workspace"):
            fn.restype = i
    lib.dclip_sizeof_gemm_args.restype = C.c_size_t
    lib.dclip_col_reduce_workspace.restype = C.c_size_t
    lib.dclip_loss_workspace.restype = C.c_size_t


def lib():
    """Load (building first if the sources are newer) and return the ctypes library."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if needs_build():
                    build()
                try:
                    loaded = C.CDLL(LIB_PATH)
                except OSError as e:  # pragma: no cover
                    raise DclipError(f"cannot load {LIB_PATH}: {e}") from e
                _declare(loaded)
                if loaded.dclip_abi_version() != 1:
                    raise DclipError("libdenseclip_b200.so ABI version mismatch; rebuild")
                if loaded.dclip_sizeof_gemm_args() != C.sizeof(GemmArgs):
                    raise DclipError("dclip_gemm_args layout mismatch between _lib.GemmArgs and the library; rebuild")
                _lib = loaded
    return _lib


def handle(device_index: int):
    """One native handle per CUDA device (created on first use). Raises if there is no sm_100 GPU."""
    h = _handles.get(device_index)
    if h is None:
        with _lock:
            h = _handles.get(device_index)
            if h is None:
                out = C.c_void_p()
                if lib().dclip_create(int(device_index), C.byref(out)) != 0:
                    raise DclipError(lib().dclip_last_error(None).decode())
                h = out
                _handles[device_index] = h
    return h


def check(h, rc: int):
    if rc != 0:
        raise DclipError(lib().dclip_last_error(h).decode())


def launch_count(device_index: int = 0) -> int:
    return int(lib().dclip_launch_count(handle(device_index)))


def reset_launch_count(device_index: int = 0) -> None:
    lib().dclip_reset_launch_count(handle(device_index))
