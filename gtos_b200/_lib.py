"""ctypes binding of libgtos_b200.so (C ABI declared in include/gtos_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Import succeeds without a GPU
(so host-side logic and the symbol table can be tested on CPU), but every kernel call raises if
the shared library is missing or the device is not sm_100.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgtos_b200.so")

vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float


class AttnDesc(C.Structure):
    """mirror of gtos_attn_desc"""
    _fields_ = [
        ("T", i32), ("S", i32), ("B", i32), ("H", i32), ("hd", i32), ("bwd_part", i32),
        ("q", vp), ("ldq", i64), ("k", vp), ("ldk", i64), ("v", vp), ("ldv", i64),
        ("scale", f32), ("p_drop", f32),
        ("scores_jt", vp), ("key_pad", vp), ("attn_mask", vp),
        ("seed_ptr", vp), ("seed_off", u64),
        ("probs", vp), ("probs_dropped", vp),
        ("out", vp), ("ldo", i64), ("out_bf16", vp),
        ("dout", vp), ("lddo", i64), ("dprobs_extra", vp),
        ("dscores_jt", vp), ("dscores_ts", vp),
        ("dq", vp), ("lddq", i64), ("dk", vp), ("lddk", i64), ("dv", vp), ("lddv", i64),
        ("dq_bf16", vp), ("dk_bf16", vp), ("dv_bf16", vp),
        ("precise", i32),
    ]


# name -> (restype, argtypes); must list EVERY symbol declared in include/gtos_b200.h
SIGNATURES = {
    "gtos_last_error": (C.c_char_p, []),
    "gtos_abi_version": (i32, []),
    "gtos_launch_count": (u64, []),
    "gtos_set_sm_reserve": (i32, [i32]),
    "gtos_device_check": (i32, []),
    "gtos_debug_read_trace": (i32, [vp, i32]),
    "gtos_debug_attn_trace": (i32, [vp, i32]),
    "gtos_cast_bf16": (i32, [vp, i64, vp, i64, i64, i32, vp]),
    "gtos_cast_colsum": (i32, [vp, i64, vp, i64, vp, i64, i32, vp]),
    "gtos_weight_prep": (i32, [vp, i32, i32, vp, i64, vp, i64, i32, vp]),
    "gtos_gemm_tn": (i32, [vp, i64, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, vp]),
    "gtos_gemm_tn_add": (i32, [vp, i64, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, vp]),
    "gtos_gemm_nn_workspace": (i64, [i32, i32, i32]),
    "gtos_gemm_nn": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, i64, vp]),
    "gtos_rel_tiling": (i32, [i32, i32, i32, i32, C.POINTER(i32)]),
    "gtos_rel_score": (i32, [vp, vp, vp, vp, i64, vp, i32, i32, i32, i32, vp]),
    "gtos_rel_attn_fusable": (i32, [i32, i32, i32, i32]),
    "gtos_rel_attn_fwd": (i32, [vp, vp, vp, vp, i64, vp, i64, vp, f32, vp, u64, vp, vp, vp, i64, vp, i32, i32, i32, i32, vp]),
    "gtos_rel_grad": (i32, [vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, vp]),
    "gtos_rel_drel": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "gtos_rel_dw_workspace": (i64, [i32, i32, i32, i32]),
    "gtos_rel_dw": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "gtos_graph_all_paths": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "gtos_graph_bfs": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "gtos_graph_paths": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, u64, vp, vp, vp]),
    "gtos_rel_dqk": (i32, [vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, vp]),
    "gtos_rel_attn_banked_fwd": (i32, [vp, i64, vp, vp, vp, i64, vp, i64, vp, vp, f32, vp, u64, vp, vp, vp, i64, vp, i32, i32,
                                       i32, i32, i32, vp]),
    "gtos_rel_grad_banked": (i32, [vp, i64, vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, i32, vp]),
    "gtos_rel_pair_keys": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "gtos_rel_segsum": (i32, [vp, vp, vp, i64, i32, vp, i64, vp, vp]),
    "gtos_rel_dw_bank": (i32, [vp, i64, vp, vp, i32, i32, i32, vp]),
    "gtos_zero_regions": (i32, [vp, vp, i32, vp]),
    "gtos_split3": (i32, [vp, i64, i64, i64, i32, vp, i64, i32, i32, vp]),
    "gtos_rel_score_f32": (i32, [vp, i64, vp, vp, i64, vp, i32, i32, i32, i32, vp]),
    "gtos_rel_grad_f32": (i32, [vp, i64, vp, vp, i64, vp, vp, i64, i32, i32, i32, i32, vp]),
    "gtos_rel_dqk_f32": (i32, [vp, i64, vp, vp, i64, i32, i32, i32, vp]),
    "gtos_relu_drop_bwd_f32": (i32, [vp, vp, vp, i64, f32, vp]),
    "gtos_gru_gate_fwd_f32": (i32, [vp, i64, vp, i64, vp, vp, i32, vp, vp, i64, vp, i64, i32, vp]),
    "gtos_gru_gate_bwd_f32": (i32, [vp, vp, i64, vp, vp, vp, i32, vp, vp, i64, vp, i64, i64, i32, vp]),
    "gtos_attn_fwd": (i32, [C.POINTER(AttnDesc), vp]),
    "gtos_attn_bwd": (i32, [C.POINTER(AttnDesc), vp]),
    "gtos_attn_bwd_dk_on_query_side": (i32, [C.POINTER(AttnDesc)]),
    "gtos_add_ln_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, vp, u64, vp]),
    "gtos_add_ln_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, vp, u64, vp]),
    "gtos_ln_param_grad": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, vp]),
    "gtos_colsum": (i32, [vp, i64, vp, i64, i32, vp]),
    "gtos_colsum_bf16": (i32, [vp, i64, vp, i64, i32, vp]),
    "gtos_dropout_bf16": (i32, [vp, i64, f32, vp, u64, vp]),
    "gtos_dropout_f32": (i32, [vp, vp, i64, f32, vp, u64, vp]),
    "gtos_relu_drop_bwd": (i32, [vp, vp, vp, vp, i64, f32, vp]),
    "gtos_token_nll_fwd": (i32, [vp, i64, i32, vp, vp, i32, vp, vp, i64, i32, i64, vp, vp, vp]),
    "gtos_token_nll_bwd": (i32, [vp, vp, i64, i32, vp, i32, vp, vp, i64, i32, i64, vp, vp, i64, vp, vp, vp, i64, vp]),
    "gtos_bank_gather": (i32, [vp, vp, i64, i32, vp, vp, vp]),
    "gtos_bank_scatter_add": (i32, [vp, vp, i64, i32, vp, i64, vp]),
    "gtos_bank_segsum": (i32, [vp, vp, vp, i64, i32, vp, i64, vp]),
    "gtos_bank_gather_mean": (i32, [vp, vp, i64, i32, i32, vp, vp, vp]),
    "gtos_embed_gather": (i32, [vp, vp, i64, i32, vp, vp, i64, f32, vp, u64, vp]),
    "gtos_embed_scatter_add": (i32, [vp, vp, i64, i32, vp, f32, vp, u64, vp]),
    "gtos_gru_weight_prep": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, i64, vp, vp]),
    "gtos_gru_step_fwd": (i32, [vp, i64, i32, vp, i64, vp, vp, i64, i32, vp, vp, i32, vp, vp, i64, vp, i64, vp, i64, i64,
                                i32, vp]),
    "gtos_gru_gate_bwd": (i32, [vp, vp, i64, vp, vp, vp, i32, vp, vp, i64, vp, i64, vp, vp, i64, i32, vp]),
    "gtos_attn_decode": (i32, [i32, i32, i32, i32, vp, i64, vp, i64, i32, i64, vp, i64, vp, i64, f32, vp, i64, vp, i64, vp,
                               vp]),
    "gtos_beam_ancestry": (i32, [vp, vp, i64, vp, i32, i32, vp]),
    "gtos_token_logprob": (i32, [vp, i64, i32, vp, vp, i32, vp, i32, vp, i64, i32, vp, i64, i32, vp]),
    "gtos_token_topk": (i32, [vp, i64, i32, vp, vp, i32, vp, i32, vp, i64, i32, i32, i32, vp, vp, vp, i64, vp]),
    "gtos_beam_update": (i32, [i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "gtos_grad_sumsq_workspace": (i64, []),
    "gtos_grad_sumsq": (i32, [vp, i64, vp, vp, vp]),
    "gtos_adam_step": (i32, [vp, vp, vp, vp, i64, i64, vp, f32, f32, f32, f32, vp, f32, vp]),
}

_lib = None


class GtosLibraryError(RuntimeError):
    pass


def load():
    """dlopen the library and bind every symbol; raises loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GtosLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m gtos_b200.build` (or __graft_entry__.build()). "
            "gtos_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().gtos_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise GtosLibraryError(f"gtos_b200 {what} failed (code {rc}): {last_error()}")
