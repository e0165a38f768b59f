"""ctypes loader for libllark_b200.so (the C-ABI declared in include/llark_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing the import of any op fails loudly.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LLARK_LIB_PATH: A/B perf experiments against another build of the same C-ABI (tools/); never a fallback
LIB_PATH = os.environ.get("LLARK_LIB_PATH") or os.path.join(_HERE, "libllark_b200.so")

_ERR = {-1: "LLARK_EINVAL (bad shape/alignment/argument)", -2: "LLARK_ECUDA", -3: "LLARK_EDRIVER", -4: "LLARK_ENCCL"}

# every symbol include/llark_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "llark_last_cuda_error", "llark_gemm", "llark_gemm_ex", "llark_gemm_clear_cache", "llark_gemm_set_mode", "llark_gemm_get_mode", "llark_splice_scan", "llark_embed_splice_fwd",
    "llark_embed_splice_bwd", "llark_cast_f32_to_bf16", "llark_cast_bf16_to_f32", "llark_rmsnorm_fwd",
    "llark_rmsnorm_bwd_workspace", "llark_rmsnorm_bwd", "llark_rope", "llark_swiglu_fwd", "llark_swiglu_bwd",
    "llark_attn_fwd", "llark_attn_decode", "llark_gemv", "llark_attn_alibi_fwd", "llark_attn_alibi_bwd", "llark_ln_fwd",
    "llark_ln_bwd_workspace", "llark_ln_bwd", "llark_gelu_fwd", "llark_gelu_bwd", "llark_attn_bwd_workspace", "llark_attn_bwd", "llark_ce_fwd", "llark_ce_bwd",
    "llark_sumsq_workspace", "llark_grad_sumsq", "llark_adamw_step", "llark_comm_unique_id", "llark_comm_init",
    "llark_comm_destroy", "llark_grad_bucket_allreduce", "llark_jb_conv_down", "llark_jb_resblock",
    "llark_jb_conv_k3", "llark_jb_vq_encode", "llark_layernorm_fwd", "llark_jb_embed_tokens", "llark_jb_sparse_attn",
    "llark_jb_sparse_attn_tc",
    "llark_quick_gelu", "llark_add3", "llark_jb_pool", "llark_jb_pool_batched", "llark_colsum",
    "llark_init", "llark_attn_bwd_rope", "llark_jb_label_cond", "llark_v32_split_tf32", "llark_v32_embed_splice_fwd", "llark_v32_rmsnorm_fwd", "llark_v32_rope",
    "llark_v32_swiglu_fwd", "llark_v32_attn_fwd", "llark_v32_ce_fwd",
]

_lib = None


class GemmEpilogue(ctypes.Structure):
    """LlarkGemmEpilogue of include/llark_b200.h."""
    _fields_ = [("kind", ctypes.c_int), ("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("ld", ctypes.c_longlong),
                ("i0", ctypes.c_int), ("i1", ctypes.c_int)]


EPI_NONE, EPI_QUICK_GELU, EPI_ROPE, EPI_SWIGLU_BWD = 0, 1, 2, 3


class LlarkError(RuntimeError):
    pass


def lib():
    """Return the loaded CDLL; raises if the CUDA extension has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LlarkError(
                f"{LIB_PATH} is missing: build the sm_100a extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'). llark_b200 has no CPU/PyTorch fallback."
            )
        import torch  # noqa: F401  (loads the NCCL / CUDA runtime libraries the extension resolves against)

        _lib = ctypes.CDLL(LIB_PATH)
        c = ctypes
        P, LL, I, F = c.c_void_p, c.c_longlong, c.c_int, c.c_float
        sig = {
            "llark_last_cuda_error": ([], I),
            "llark_gemm": ([P, I, LL, P, I, LL, P, LL, I, P, I, P, LL, I, I, I, I, I, P], I),
            "llark_gemm_ex": ([P, I, LL, P, I, LL, P, LL, I, P, I, P, LL, I, I, I, I, I, P, P], I),
            "llark_gemm_clear_cache": ([], None),
            "llark_gemm_set_mode": ([I], None),
            "llark_gemm_get_mode": ([], I),
            "llark_splice_scan": ([P, I, I, LL, LL, P, P, P, P], I),
            "llark_embed_splice_fwd": ([P, P, P, P, P, P, P, I, I, I, LL, P], I),
            "llark_embed_splice_bwd": ([P, P, P, P, P, P, P, I, I, I, LL, I, P], I),
            "llark_cast_f32_to_bf16": ([P, P, LL, P], I),
            "llark_cast_bf16_to_f32": ([P, P, LL, P], I),
            "llark_rmsnorm_fwd": ([P, P, P, P, LL, I, F, P], I),
            "llark_rmsnorm_bwd_workspace": ([LL, I], LL),
            "llark_rmsnorm_bwd": ([P, P, P, P, P, P, P, I, LL, I, P, P], I),
            "llark_rope": ([P, P, P, LL, I, I, I, LL, I, P], I),
            "llark_swiglu_fwd": ([P, P, LL, I, P], I),
            "llark_swiglu_bwd": ([P, P, P, LL, I, P], I),
            "llark_attn_fwd": ([P, P, P, P, I, I, I, I, F, P], I),
            "llark_attn_decode": ([P, P, P, P, P, I, I, I, I, F, P], I),
            "llark_gemv": ([P, LL, P, LL, P, LL, P, P, LL, I, I, I, P], I),
            "llark_attn_alibi_fwd": ([P, P, P, P, P, I, I, I, I, F, P], I),
            "llark_attn_alibi_bwd": ([P, P, P, P, P, P, P, I, I, I, I, F, P, P], I),
            "llark_ln_fwd": ([P, LL, P, P, LL, P, P, LL, I, F, P], I),
            "llark_ln_bwd_workspace": ([LL, I], LL),
            "llark_ln_bwd": ([P, LL, P, P, LL, P, P, P, P, LL, P, I, LL, I, P, P], I),
            "llark_gelu_fwd": ([P, P, LL, P], I),
            "llark_gelu_bwd": ([P, P, P, LL, P], I),
            "llark_attn_bwd_workspace": ([I, I, I], LL),
            "llark_attn_bwd": ([P, P, P, P, P, P, I, I, I, I, F, P, P], I),
            "llark_ce_fwd": ([P, LL, I, P, I, I, P, P, P, P], I),
            "llark_ce_bwd": ([P, LL, I, P, I, I, P, P, P, F, P, P], I),
            "llark_sumsq_workspace": ([], LL),
            "llark_grad_sumsq": ([P, LL, P, I, P, P], I),
            "llark_adamw_step": ([P, P, P, P, LL, P, F, F, F, F, F, I, F, F, I, P], I),
            "llark_comm_unique_id": ([P], I),
            "llark_comm_init": ([P, P, I, I], I),
            "llark_comm_destroy": ([P], I),
            "llark_grad_bucket_allreduce": ([P, P, LL, P], I),
            "llark_jb_conv_down": ([P, P, P, P, I, I, I, I, P], I),
            "llark_jb_resblock": ([P, P, P, P, P, P, I, I, I, P], I),
            "llark_jb_conv_k3": ([P, P, P, P, I, I, I, I, P], I),
            "llark_jb_vq_encode": ([P, P, P, I, I, I, P], I),
            "llark_layernorm_fwd": ([P, P, P, P, I, LL, I, F, P], I),
            "llark_jb_embed_tokens": ([P, P, P, P, P, P, I, I, I, P], I),
            "llark_jb_sparse_attn": ([P, P, I, I, I, I, I, I, P], I),
            "llark_jb_sparse_attn_tc": ([P, P, I, I, I, I, I, I, I, P], I),
            "llark_quick_gelu": ([P, P, LL, P], I),
            "llark_add3": ([P, P, P, P, LL, P], I),
            "llark_jb_pool": ([P, P, I, I, I, P], I),
            "llark_jb_pool_batched": ([P, P, I, LL, I, I, I, P], I),
            "llark_colsum": ([P, P, LL, I, I, P], I),
            "llark_init": ([], I),
            "llark_attn_bwd_rope": ([P, P, P, P, P, P, I, I, I, I, F, P, P, P, P], I),
            "llark_jb_label_cond": ([P, P, P, P, P, P, I, I, P], I),
            "llark_v32_split_tf32": ([P, P, P, LL, P], I),
            "llark_v32_embed_splice_fwd": ([P, P, P, P, P, P, P, I, I, I, LL, P], I),
            "llark_v32_rmsnorm_fwd": ([P, P, P, LL, I, F, P], I),
            "llark_v32_rope": ([P, P, P, LL, I, I, I, LL, P], I),
            "llark_v32_swiglu_fwd": ([P, P, LL, I, P], I),
            "llark_v32_attn_fwd": ([P, P, P, I, I, I, I, F, P], I),
            "llark_v32_ce_fwd": ([P, LL, I, P, I, I, P, P, P], I),
        }
        for name, (argtypes, restype) in sig.items():
            if os.environ.get("LLARK_LIB_PATH") and not hasattr(_lib, name):
                continue  # A/B run against an older build of the C-ABI: newer entry points are simply absent
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
    return _lib


def check(rc, what):
    if rc != 0:
        extra = ""
        if rc in (-2, -3, -4):
            extra = f" (code {lib().llark_last_cuda_error()})"
        raise LlarkError(f"{what} failed: {_ERR.get(rc, rc)}{extra}")
