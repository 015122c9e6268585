"""ctypes binding of libspeechclip_b200.so (the C ABI declared in include/speechclip_b200.h).

The product path has no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised.  ``ensure_built()`` compiles the library in-tree with nvcc when the
sources are newer than the binary (nvcc cross-compiles sm_100a without a GPU).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB_PATH = os.path.join(_HERE, "libspeechclip_b200.so")

F32, F16, BF16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_QUICK_GELU = 0, 1, 2

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_INCLUDE, "speechclip_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def ensure_built(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libspeechclip_b200.so for sm_100a (in-tree, so it travels to the GPU box)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs, procs = [], []
    for src in _sources():  # one nvcc per translation unit, in parallel
        obj = src[:-3] + ".o"
        objs.append(obj)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else []) + os.environ.get("SCB_NVCC_EXTRA", "").split()
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: " + r.stdout)
    return LIB_PATH


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p),
        ("a_inner", ctypes.c_int64), ("a_rows", ctypes.c_int64), ("a_row_stride", ctypes.c_int64), ("a_batch_stride", ctypes.c_int64),
        ("batch", ctypes.c_int32), ("m_per_batch", ctypes.c_int32),
        ("kb_per_tap", ctypes.c_int32), ("tap_row_shift", ctypes.c_int32), ("a_col0", ctypes.c_int32), ("a_group_cols", ctypes.c_int32),
        ("b", ctypes.c_void_p),
        ("b_row_stride", ctypes.c_int64), ("b_group_stride", ctypes.c_int64),
        ("n", ctypes.c_int32), ("k", ctypes.c_int32), ("groups", ctypes.c_int32),
        ("out", ctypes.c_void_p),
        ("out_dtype", ctypes.c_int32), ("out_group_cols", ctypes.c_int32),
        ("ldc", ctypes.c_int64), ("out_batch_stride", ctypes.c_int64),
        ("out2", ctypes.c_void_p),
        ("out2_dtype", ctypes.c_int32), ("ab_format", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("residual", ctypes.c_void_p),
        ("residual_dtype", ctypes.c_int32), ("act", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("residual_ld", ctypes.c_int64), ("residual_batch_stride", ctypes.c_int64),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64),
    ]


_lib = None
_lock = threading.Lock()


def _declare(lib):
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    lib.scb_abi_version.restype = i32
    lib.scb_last_error.restype = ctypes.c_char_p
    lib.scb_launch_count.restype = i64
    lib.scb_conv0_scratch_bytes.restype = i64
    lib.scb_conv0_scratch_bytes.argtypes = [i32]
    lib.scb_gemm_workspace_bytes.restype = i64
    lib.scb_gemm_workspace_bytes.argtypes = []
    lib.scb_infonce_scratch_bytes.restype = i64
    lib.scb_infonce_scratch_bytes.argtypes = [i32]
    sig = {
        "scb_gemm": [ctypes.POINTER(GemmArgs), vp],
        "scb_sgemm": [vp, i64, i64, vp, i64, i64, vp, i64, i32, i32, i32, f32, f32, vp],
        "scb_attention_fwd": [vp, vp, vp, vp, i32] + [i64] * 8 + [vp, i32, i32, i32, i32, i32, f32, i32, vp],
        "scb_cls_attention_fwd": [vp, vp, i32, i64, i64, i32, i32, vp, i32, i32, i32, i32, f32, vp, vp, vp, i32, f32, vp, i32, vp],
        "scb_cls_attention_bwd": [vp, vp, i32, i64, i64, i32, i32, vp, i32, i32, i32, i32, f32, vp, vp, vp, i32, vp, f32, vp, i32, vp],
        "scb_rng_advance": [vp, vp],
        "scb_image_normalize": [vp, i32, i32, i32, vp, vp, vp, vp],
        "scb_pad_rows": [vp, vp, vp, i32, i64, vp, vp],
        "scb_masked_mean_fwd": [vp, vp, i32, i32, i32, vp, vp],
        "scb_masked_mean_bwd": [vp, vp, i32, i32, i32, vp, vp],
        "scb_attentive_pool_fwd": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp],
        "scb_tanh_softmax_dim1": [vp, vp, i32, i32, i32, vp, vp],
        "scb_relu_fwd": [vp, vp, i64, vp],
        "scb_relu_bwd": [vp, vp, vp, i64, vp],
        "scb_dropout_mask": [vp, i32, f32, vp, i64, vp],
        "scb_dropout_rows": [vp, vp, i64, f32, vp, i32, vp],
        "scb_frame_lengths": [vp, i32, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp],
        "scb_lengths_to_i32": [vp, i32, i32, i32, vp, vp],
        "scb_wav_prepare": [vp, i64, i32, vp, vp, i64, i32, vp, vp, i64, vp],
        "scb_conv0_groupnorm_gelu": [vp, i64, i32, i32, vp, vp, vp, vp, f32, vp, i32, i64, vp, i64, vp],
        "scb_conv0_layernorm_gelu": [vp, i64, i32, i32, vp, vp, vp, vp, f32, vp, i32, i64, vp, i64, vp],
        "scb_posconv_pack": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "scb_patchify": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "scb_broadcast_row": [vp, vp, vp, i32, i64, i32, i32, vp],
        "scb_cast_rows": [vp, i32, i64, vp, i32, i64, i64, i32, vp],
        "scb_transpose": [vp, i32, i64, vp, i32, i64, i32, i32, vp],
        "scb_layernorm_fwd": [vp, i32, vp, vp, vp, vp, i32, vp, i64, i32, i64, i64, f32, i32, vp],
        "scb_layernorm_bwd": [vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
        "scb_l2norm_fwd": [vp, vp, vp, i32, i32, vp],
        "scb_l2norm_bwd": [vp, vp, vp, vp, i32, i32, vp],
        "scb_weighted_sum_fwd": [vp, i32, i64, vp, i32, i32, vp, vp, i32, i64, i32, i32, i64, i64, vp],
        "scb_weighted_sum_bwd": [vp, i32, i64, vp, i32, i32, vp, i64, i32, i32, i64, i64, vp, vp, f32, vp],
        "scb_rows_bias_act": [vp, i64, vp, vp, i64, i32, vp, vp, i64, i64, i32, vp],
        "scb_gelu_bwd": [vp, vp, vp, i64, vp],
        "scb_column_sum": [vp, i32, i64, i64, i32, vp, f32, vp],
        "scb_infonce": [vp, vp, vp, i32, i32, vp, f32, f32, i32, i32, i32, i32, vp, vp, f32, vp, vp, vp, vp, vp, i64, vp],
        "scb_adam_step": [vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, f32, f32, i32, vp, vp, vp],
        "scb_retrieval_rank": [vp, i64, i32, i32, vp, vp, vp, vp, vp],
        "scb_mq_attention_fwd": [vp, vp, i32, i64, i64, i32, i32, vp, i32, i32, i32, i32, i32, f32, vp, vp, f32, vp, i32, vp],
        "scb_mq_attention_bwd": [vp, vp, i32, i64, i64, i32, i32, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, i32, vp, f32, vp, i32, vp],
        "scb_batchnorm_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, i32, vp],
        "scb_batchnorm_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "scb_vq_forward": [vp, vp, vp, i32, i32, i32, i64, vp, i32, f32, vp, vp, vp],
        "scb_vq_backward": [vp, vp, i32, i32, i64, vp, f32, vp, vp],
        "scb_cosine_bwd_rows": [vp, vp, vp, vp, vp, i32, i32, vp],
        "scb_vq_diagnostics": [vp, i32, i32, i64, vp, vp, vp, vp, vp, vp],
        "scb_keyword_embed": [vp, vp, vp, i64, i64, i32, i32, i32, vp, vp, vp],
        "scb_attention_small_bwd": [vp, i32, vp, vp, i32, i32, i32, i32, f32, i32, vp],
        "scb_token_embed": [vp, vp, vp, i32, i32, i32, i64, vp, vp],
        "scb_gather_rows": [vp, vp, i32, i32, i32, vp, vp],
        "scb_softmax_rows": [vp, i64, i64, i32, vp, i32, vp, i32, i64, i32, vp],
        "scb_split_tf32": [vp, i64, vp, i64, i32, i32, vp],
        "scb_act16_fwd": [vp, i32, i32, vp, i64, vp],
        "scb_act_bwd": [vp, vp, i32, i32, vp, i64, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = i32


EXPORTS = ["scb_abi_version", "scb_last_error", "scb_launch_count", "scb_conv0_scratch_bytes", "scb_gemm_workspace_bytes", "scb_infonce_scratch_bytes",
           "scb_gemm", "scb_sgemm", "scb_attention_fwd", "scb_cls_attention_fwd", "scb_cls_attention_bwd", "scb_frame_lengths",
           "scb_lengths_to_i32", "scb_wav_prepare", "scb_conv0_groupnorm_gelu", "scb_conv0_layernorm_gelu", "scb_posconv_pack", "scb_patchify",
           "scb_broadcast_row", "scb_cast_rows", "scb_transpose", "scb_layernorm_fwd", "scb_layernorm_bwd", "scb_l2norm_fwd",
           "scb_l2norm_bwd", "scb_weighted_sum_fwd", "scb_weighted_sum_bwd", "scb_rows_bias_act", "scb_gelu_bwd", "scb_column_sum",
           "scb_infonce", "scb_adam_step", "scb_retrieval_rank", "scb_mq_attention_fwd", "scb_mq_attention_bwd", "scb_batchnorm_fwd",
           "scb_batchnorm_bwd", "scb_vq_forward", "scb_vq_backward", "scb_cosine_bwd_rows", "scb_vq_diagnostics", "scb_keyword_embed",
           "scb_rng_advance", "scb_dropout_mask", "scb_dropout_rows", "scb_image_normalize", "scb_pad_rows", "scb_masked_mean_fwd",
           "scb_masked_mean_bwd", "scb_attentive_pool_fwd", "scb_tanh_softmax_dim1", "scb_relu_fwd", "scb_relu_bwd",
           "scb_attention_small_bwd", "scb_token_embed", "scb_gather_rows", "scb_softmax_rows", "scb_split_tf32", "scb_act16_fwd", "scb_act_bwd"]


def load():
    """Return the loaded library; raise loudly if it is absent (no CPU / eager fallback exists)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the CUDA extension is the only implementation of the hot path. "
                    "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).")
            lib = ctypes.CDLL(LIB_PATH)
            _declare(lib)
            _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {load().scb_last_error().decode()}")


def launch_count() -> int:
    return int(load().scb_launch_count())
