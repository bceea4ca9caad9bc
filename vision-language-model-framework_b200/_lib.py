"""ctypes binding of ``libvlmfw_b200.so`` — the C-ABI in ``include/vlmfw_b200.h``.

There is deliberately no fallback here: if the shared library is missing or a call fails the caller gets an
exception (``VlmfwError``), never a silent PyTorch/CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# VLMFW_B200_LIB: kernel-tuning experiments load an alternative build of the same sources (tools/ only)
LIB_PATH = os.environ.get("VLMFW_B200_LIB") or os.path.join(_HERE, "libvlmfw_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

OK, E_INVALID, E_NODEVICE, E_CUDA, E_STATE, E_NOTFOUND = 0, -1, -2, -3, -4, -5
E_RANGE = -6
DT_BF16, DT_F16, DT_F32 = 0, 1, 2
CDT_F16, CDT_BF16 = 0, 1


class VlmfwError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"vlmfw_b200 error {code}: {message}")
        self.code = code


class TowerCfg(C.Structure):
    _fields_ = [
        ("arch", C.c_int32), ("image_size", C.c_int32), ("patch_size", C.c_int32), ("hidden", C.c_int32),
        ("heads", C.c_int32), ("mlp", C.c_int32), ("layers", C.c_int32), ("has_cls", C.c_int32),
        ("has_pre_ln", C.c_int32), ("conv_bias", C.c_int32), ("act", C.c_int32), ("pool", C.c_int32),
        ("ln_eps", C.c_float), ("pix_mean", C.c_float * 3), ("pix_std", C.c_float * 3), ("compute_dtype", C.c_int32),
    ]


class Head(C.Structure):
    _fields_ = [
        ("bn_weight", C.c_void_p), ("bn_bias", C.c_void_p), ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("classes", C.c_int32), ("hidden", C.c_int32), ("bn_eps", C.c_float), ("logit_offset", C.c_int32),
    ]


_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class GemmDesc(C.Structure):
    """``vlmfw_gemm_desc``: the GEMM with LayerNorm folded around it (producer / consumer sides, see the header)."""
    _fields_ = [
        ("a", _vp), ("lda", _i32), ("w", _vp), ("ldw", _i32), ("dtype", _i32), ("m", _i32), ("n", _i32), ("k", _i32),
        ("bias", _vp), ("epi", _i32), ("act", _i32), ("out", _vp), ("ldo", _i32), ("resid", _vp), ("ldr", _i32),
        ("block_n", _i32), ("ln_stats", _vp), ("ln_np", _i32), ("ln_dim", _i32), ("ln_eps", _f32),
        ("h16_out", _vp), ("ld_h16", _i32), ("stats_out", _vp), ("acc_scale", _f32), ("sat_count", _vp),
    ]


class ProbeStep(C.Structure):
    """``vlmfw_probe_step``: every pointer / scalar of one fused data-parallel optimisation step."""
    _fields_ = [
        ("x", _vp), ("targets", _vp), ("batch", _i32), ("dim", _i32), ("classes", _i32), ("n_global", _i32), ("row0", _i64),
        ("stats", _vp), ("params", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("bucket", _vp), ("logits", _vp),
        ("scratch", _vp), ("keep", _vp), ("class_weight", _vp), ("running_mean", _vp), ("running_var", _vp),
        ("steps_done", _vp), ("mask_key", C.c_uint64), ("mask_step_stride", C.c_uint64),
        ("dropout_p", _f32), ("bn_eps", _f32), ("bn_momentum", _f32), ("lr", _f32), ("beta1", _f32), ("beta2", _f32),
        ("eps", _f32), ("weight_decay", _f32),
    ]


# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against include/vlmfw_b200.h
SIGNATURES = {
    "vlmfw_abi_version": (C.c_int, []),
    "vlmfw_last_error": (C.c_char_p, []),
    "vlmfw_device_ok": (C.c_int, []),
    "vlmfw_tower_create": (C.c_int, [C.POINTER(TowerCfg), C.POINTER(_vp)]),
    "vlmfw_tower_destroy": (None, [_vp]),
    "vlmfw_tower_set_weight": (C.c_int, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "vlmfw_tower_ready": (C.c_int, [_vp]),
    "vlmfw_tower_prepare": (C.c_int, [_vp, _vp]),
    "vlmfw_tower_weight_scale": (C.c_int, [_vp, C.c_char_p, C.POINTER(_f32)]),
    "vlmfw_tower_set_cls_tail": (C.c_int, [_vp, _i32]),
    "vlmfw_tower_cls_tail": (_i32, [_vp]),
    "vlmfw_tower_saturation": (C.c_int, [_vp, C.POINTER(_i64), _i32, _vp]),
    "vlmfw_tower_workspace_bytes": (_sz, [_vp, _i32]),
    "vlmfw_tower_tokens": (_i32, [_vp]),
    "vlmfw_tower_forward": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _sz, _vp, _vp]),
    "vlmfw_tower_hidden_ptr": (_vp, [_vp, _vp]),
    "vlmfw_tower_hidden16_ptr": (_vp, [_vp, _vp, _i32]),
    "vlmfw_tower_last_launches": (_i64, [_vp]),
    "vlmfw_tower_forward_host": (C.c_int, [_vp, _vp, _i32, _vp]),
    "vlmfw_preproc_create": (C.c_int, [_i32, _i32, C.POINTER(_vp)]),
    "vlmfw_preproc_destroy": (None, [_vp]),
    "vlmfw_preproc_workspace_bytes": (_sz, [_vp, _vp, _vp, _i32]),
    "vlmfw_preproc_run": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "vlmfw_preproc_last_launches": (_i64, [_vp]),
    "vlmfw_resize_weights": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "vlmfw_probe_forward": (C.c_int, [_vp, _i32, _i32, C.POINTER(Head), _i32, _vp, _i32, _vp, _vp]),
    "vlmfw_confusion_update": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "vlmfw_probe_bn_stats": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "vlmfw_probe_linear_grads": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _f32, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _f32, _vp, _vp, _vp, _vp]),
    "vlmfw_probe_adamw": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "vlmfw_probe_deep_scratch_floats": (_sz, [_i32, _i32, _i32, _i32]),
    "vlmfw_probe_deep_grads": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _f32, _vp, _vp, _vp, _vp, _f32,
                                         _vp, _vp, _vp, _vp]),
    "vlmfw_probe_adamw_ex": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _i32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "vlmfw_probe_adamw_ex_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _f32, _f32, _f32, _f32, _f32, _vp]),
    "vlmfw_probe_bn_running_update": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "vlmfw_dropout_keep_mask": (C.c_int, [_vp, _i32, _i32, _i64, C.c_uint64, _f32, _vp]),
    "vlmfw_probe_adamw_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _f32, _f32, _f32, _f32, _f32, _vp]),
    "vlmfw_dropout_keep_mask_dev": (C.c_int, [_vp, _i32, _i32, _i64, C.c_uint64, _vp, C.c_uint64, _f32, _vp]),
    "vlmfw_counter_add": (C.c_int, [_vp, _i64, _vp]),
    "vlmfw_comm_create": (C.c_int, [_i32, _i32, _i64, _i64, C.POINTER(_vp)]),
    "vlmfw_comm_handle_bytes": (_i32, []),
    "vlmfw_comm_handle": (C.c_int, [_vp, _vp]),
    "vlmfw_comm_connect": (C.c_int, [_vp, _vp]),
    "vlmfw_comm_destroy": (None, [_vp]),
    "vlmfw_comm_error": (C.c_int, [_vp, C.POINTER(_i32)]),
    "vlmfw_probe_linear_step_dp": (C.c_int, [C.POINTER(ProbeStep), _vp, _vp]),
    "vlmfw_gemm_tn": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _i32,
                                _i32, _i32, _i32, _i32, _i32, _vp]),
    "vlmfw_gemm_tn_ex": (C.c_int, [C.POINTER(GemmDesc), _vp]),
    "vlmfw_gemm_stats_parts": (_i32, [_i32, _i32]),
    "vlmfw_row_stats_cast": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _vp, _i32, _vp]),
    "vlmfw_ln_fold_weight": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "vlmfw_layernorm": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _f32, _vp, _i32, _vp]),
    "vlmfw_attention": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "vlmfw_attention_cls": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "vlmfw_im2col": (C.c_int, [_vp, _i32, _i32, _i32, _i32, C.POINTER(_f32), C.POINTER(_f32), _vp, _i32, _i32, _vp]),
    "vlmfw_pool": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _f32, _vp, _vp]),
    "vlmfw_pack_16": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i32, _vp]),
}

_lib: Optional[C.CDLL] = None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/*.cu`` for sm_100a into the in-tree ``libvlmfw_b200.so`` (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC_DIR, f) for f in sorted(os.listdir(CSRC_DIR)) if f.endswith((".cu", ".cuh", ".sh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "vlmfw_b200.h"))
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in srcs if os.path.exists(s))
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    proc = subprocess.run(["bash", os.path.join(CSRC_DIR, "build.sh")], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libvlmfw_b200.so failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout.strip())
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library and attach signatures. Raises if it has not been built — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "vlmfw_b200 has no PyTorch/CPU fallback for the vision-tower path.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().vlmfw_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    if rc != OK:
        raise VlmfwError(rc, last_error())


def device_ok() -> bool:
    return load().vlmfw_device_ok() == OK
