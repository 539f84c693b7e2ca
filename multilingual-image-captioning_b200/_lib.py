"""ctypes binding of the C-ABI library `libmic_b200.so` (include/mic_b200.h).

The product path has no fallback: if the shared library is missing or a symbol is absent, importing
callers get a loud error.  `build()` compiles it in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmic_b200.so")
CSRC = os.path.join(_HERE, "csrc")

P, I, L, F = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argument ctypes (return type is always int unless listed in _RET)
SIGNATURES = {
    "mic_abi_version": [],
    "mic_launch_options": [I, I, I],
    "mic_attention_impl": [I],
    "mic_decoder_plan_bytes": [I],
    "mic_decoder_packed_bytes": [I, I, I],
    "mic_decoder_pack_weights": [P, P, I, I, I, P],
    "mic_decoder_plan_init": [P, P, P, I, P, P, I, I, I, I, I, I, I, L, I, F],
    "mic_decoder_step": [P, P, I, I, I, P, P, P, I],
    "mic_barrier_bench": [P, P, I, I],
    "mic_decoder_cross_kv_tiles_bytes": [I, I, I],
    "mic_decoder_pack_cross_kv": [P, P, L, I, I, I, I, I, P],
    "mic_f32_gemm": [P, P, L, P, L, I, I, I, I, P, I, P, L, P, L],
    "mic_f32_layernorm": [P, P, P, P, F, P, I, I],
    "mic_f32_attention": [P, P, L, P, L, P, L, P, L, P, I, I, I, I, I, I, F],
    "mic_f32_embed": [P, P, P, F, P, I, I, P, I, I],
    "mic_f32_patchify": [P, P, P, I, I, I, I, I],
    "mic_f32_vit_embed": [P, P, P, P, P, P, I, I, I],
    "mic_f32_ce_rows": [P, P, L, P, I, I, F, P, P],
    "mic_gemm_bf16": [P, I, I, P, L, P, L, I, I, I, P, L, I, I, P, I, P, P, L, I, I, I, P, I, F],
    "mic_lm_head_num_partials": [I],
    "mic_lm_head_ce_stats": [P, P, L, P, L, P, P, I, I, I, P, P, P, P, P, L],
    "mic_ce_softmax_bwd_workspace_floats": [I, L],
    "mic_ce_softmax_bwd": [P, P, L, P, P, P, F, F, I, I, P, P, P],
    "mic_ce_finalize": [P, P, P, P, P, P, I, I, I, F, P, P, P, P],
    "mic_lm_head_ce_grad": [P, P, L, P, L, P, P, P, P, F, F, I, I, I, P, L],
    "mic_lm_head_search_num_partials": [I],
    "mic_lm_head_search": [P, P, L, P, L, P, I, I, I, I, P, P, P, P, P, P, P, P],
    "mic_pack_kmajor_tiles_bytes": [L, I, I],
    "mic_pack_kmajor_tiles": [P, P, L, L, I, I, P],
    "mic_lm_head_search_packed": [P, P, P, P, I, I, I, I, P, P, P, P, P, P, P, P],
    "mic_layernorm_fwd": [P, P, P, P, F, P, P, P, I, I],
    "mic_residual_ln_fwd": [P, P, P, P, P, P, F, P, I, I],
    "mic_layernorm_bwd_workspace_floats": [I, I],
    "mic_layernorm_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, I, I],
    "mic_colsum_workspace_floats": [I, I],
    "mic_act_bwd_colsum": [P, P, L, P, L, I, P, L, P, I, P, P, I, I, P, I, F],
    "mic_embed_ln_fwd": [P, P, P, I, I, P, P, F, P, P, F, P, P, P, P, I, I, P, I, F],
    "mic_embed_bwd": [P, P, P, F, P, P, I, I, I, I],
    "mic_batch_sum": [P, P, I, I, I, P, L],
    "mic_patchify": [P, P, P, I, I, I, I, I],
    "mic_patchify_u8": [P, P, P, I, I, I, I, I, P, P],
    "mic_resize_crop_u8": [P, P, P, I, I, I, P],
    "mic_vit_embed_ln_fwd": [P, P, P, P, P, P, P, F, I, P, P, P, P, I, I, I],
    "mic_drop_cls_rows": [P, P, P, I, I, I],
    "mic_adamw": [P, P, P, P, P, P, L, F, F, F, F, F, F, F, F],
    "mic_cast_f32_to_bf16": [P, P, P, L],
    "mic_attention_fwd": [P, P, L, P, L, P, L, P, L, P, P, I, I, I, I, I, I, F],
    "mic_attention_bwd": [P, P, L, P, L, P, L, P, L, P, L, P, P, I, P, L, P, L, P, L, I, I, I, I, I, F],
    "mic_decode_attention": [P, P, L, P, P, L, P, I, I, I, P, L, I, I, I, F],
    "mic_search_merge": [P, P, P, P, P, I, I, P, P, P, I, I, I, P, P],
    "mic_beam_step": [P, P, P, I, I, I, I, I, I, I, I, F, P, P, P, P, P, P, P, P, I],
    "mic_beam_cond": [P, P, P, P, I, I, I, I, F, I, P],
    "mic_beam_finalize": [P, P, P, P, P, P, I, I, I, P, P],
    "mic_greedy_step": [P, P, I, I, I, I, I, I, P, P, P, P],
    "mic_greedy_cond": [P, P, I, I, I, P],
}
EXPORTED = sorted(list(SIGNATURES) + ["mic_last_error"])


class MicError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libmic_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise MicError("building libmic_b200.so failed")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the library once; missing file or symbols are fatal (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MicError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"(or `make -C {CSRC}`); there is no fallback path")
    l = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(l, name)       # AttributeError if the symbol is not exported -> loud
        fn.argtypes = args
        fn.restype = L if name.endswith(("_workspace_floats", "_plan_bytes", "_packed_bytes", "_tiles_bytes")) else I
    l.mic_last_error.argtypes = []
    l.mic_last_error.restype = C.c_char_p
    if l.mic_abi_version() != 2:
        raise MicError("libmic_b200.so ABI version mismatch")
    _lib = l
    return l


def check(rc: int, what: str):
    if rc != 0:
        raise MicError(f"{what} failed (status {rc}): {lib().mic_last_error().decode()}")
