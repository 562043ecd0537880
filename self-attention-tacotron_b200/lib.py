"""ctypes binding of libsatk.so (the C ABI declared in include/satk.h).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C csrc``) and must be present:
every op below raises ``SatkError`` if the shared object is missing or a call fails — there is no
CPU / PyTorch fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SATK_LIB_PATH") or os.path.join(_HERE, "libsatk.so")   # override: developer builds (make pt)


class SatkError(RuntimeError):
    pass


fp = C.c_void_p
i32 = C.c_int
i64 = C.c_longlong
f32 = C.c_float


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32), ("transA", i32), ("transB", i32),
        ("A", fp), ("lda", i64), ("B", fp), ("ldb", i64), ("C", fp), ("ldc", i64),
        ("alpha", f32), ("beta", f32), ("bias", fp), ("act", i32), ("residual", fp), ("ldres", i64),
        ("keep_mask", fp), ("keep_scale", f32),
        ("batch1", i32), ("batch2", i32),
        ("sA1", i64), ("sA2", i64), ("sB1", i64), ("sB2", i64), ("sC1", i64), ("sC2", i64),
        ("taps", i32), ("shift0", i32), ("tap_dir", i32), ("seq_len", i32), ("sBtap", i64),
        ("shift_per_batch1", i32), ("split_k", i32), ("causal_skip", i32), ("kshift0", i32), ("kshift_per_batch1", i32),
        ("bank_widths", i32), ("bank_a_kstep", i32), ("bank_c_nstep", i32),
        ("zcoord", i32), ("za_row", i32), ("za_k", i32), ("zb_row", i32), ("zb_k", i32), ("zc_col", i32),
        ("a_rows", i64), ("a_cols", i64), ("b_rows", i64), ("b_cols", i64), ("c_cols", i64),
        ("precision", i32),
    ]


class LstmFwdDesc(C.Structure):
    _fields_ = [
        ("T", i32), ("B", i32), ("H", i32), ("reverse", i32),
        ("xg", fp), ("Wh", fp), ("lengths", fp), ("mask_c", fp), ("mask_h", fp),
        ("zc", f32), ("zh", f32), ("forget_bias", f32),
        ("out", fp), ("ld_out", i64), ("gates", fp), ("c_prev", fp), ("h_prev", fp),
    ]


class LstmBwdDesc(C.Structure):
    _fields_ = [
        ("T", i32), ("B", i32), ("H", i32), ("reverse", i32),
        ("Wh", fp), ("lengths", fp), ("mask_c", fp), ("mask_h", fp),
        ("zc", f32), ("zh", f32),
        ("gates", fp), ("c_prev", fp), ("dout", fp), ("ld_dout", i64), ("dgates", fp), ("step_end", fp),
    ]


class AttnRnnFwdDesc(C.Structure):
    _fields_ = [
        ("Td", i32), ("B", i32), ("Tt", i32), ("H", i32), ("A1", i32), ("A2", i32), ("M1", i32), ("M2", i32),
        ("att_kernel", i32), ("mode", i32), ("cumulative", i32),
        ("xg", fp), ("Wrec", fp), ("mask_c", fp), ("mask_h", fp),
        ("zc", f32), ("zh", f32), ("forget_bias", f32),
        ("lengths", fp), ("keys1", fp), ("values1", fp), ("Wq1", fp), ("v1", fp), ("b1", fp),
        ("loc_conv_w", fp), ("loc_conv_b", fp), ("loc_layer_w", fp), ("att_filters", i32),
        ("keys2", fp), ("values2", fp), ("Wq2", fp), ("v2", fp),
        ("x2", fp), ("align1", fp), ("align2", fp),
        ("gates", fp), ("c_prev", fp), ("h_prev", fp), ("soft1", fp), ("q_save", fp),
        ("agent_w", fp), ("agent_b", fp), ("u_save", fp), ("state_final", fp),
    ]


class AttnRnnBwdDesc(C.Structure):
    _fields_ = [
        ("f", AttnRnnFwdDesc),
        ("dx2", fp), ("dgates", fp), ("dq", fp), ("dkeys1", fp), ("dkeys2", fp),
        ("dv1", fp), ("dv2", fp), ("dloc_conv_w", fp), ("dloc_conv_b", fp), ("dloc_layer_w", fp),
        ("dagent_w", fp), ("dagent_b", fp), ("step_end", fp), ("de_ws", fp), ("sync_ws", fp),
    ]


class RowGemmDesc(C.Structure):
    _fields_ = [
        ("M", i32), ("K", i32), ("A", fp), ("lda", i64), ("a_tstride", i64), ("t_ptr", fp), ("nmat", i32),
        ("W", fp * 3), ("bias", fp * 3), ("C", fp * 3), ("ldc", i64 * 3), ("c_tstride", i64 * 3),
        ("N", i32 * 3), ("act", i32 * 3), ("residual", fp * 3), ("ldres", i64 * 3), ("res_tstride", i64 * 3),
        ("a_pstride", i64), ("c_pstride", i64 * 3),
        ("lstm_H", i32), ("lstm_c", fp), ("lstm_h", fp), ("zc", f32), ("zh", f32), ("forget_bias", f32),
        ("lstm_out", fp), ("ld_out", i64), ("out_pstride", i64), ("lstm_hdst", fp), ("ld_hdst", i64), ("hdst_pstride", i64),
    ]


class AttnStepDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("Tt", i32), ("A1", i32), ("A2", i32), ("M1", i32), ("M2", i32),
        ("att_kernel", i32), ("att_filters", i32), ("mode", i32), ("cumulative", i32), ("use_agent", i32),
        ("t_ptr", fp), ("lengths", fp), ("q", fp), ("ldq", i64),
        ("Wq1", fp), ("Wq2", fp), ("q_x", fp), ("q_x_ld", i64), ("q_x_pstride", i64), ("q_in", i32),
        ("keys1", fp), ("values1", fp), ("v1", fp), ("b1", fp),
        ("loc_conv_w", fp), ("loc_conv_b", fp), ("loc_layer_w", fp),
        ("keys2", fp), ("values2", fp), ("v2", fp), ("agent_w", fp), ("agent_b", fp),
        ("aprev", fp), ("alpha", fp), ("u", fp),
        ("ctx_dst0", fp), ("ld0", i64), ("pstride0", i64), ("ctx_dst1", fp), ("ld1", i64), ("pstride1", i64),
        ("align1", fp), ("align2", fp), ("forced1", fp), ("forced2", fp),
    ]


class SaStepDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("D", i32), ("heads", i32), ("Tmax", i32), ("t_ptr", fp), ("q", fp), ("ldq", i64),
        ("Kc", fp), ("Vc", fp), ("out", fp), ("ldo", i64), ("probs", fp),
    ]


class MlpChainDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("K0", i32), ("nlayers", i32), ("t_ptr", fp), ("x", fp), ("x_ld", i64), ("x_tstride", i64),
        ("W", fp * 3), ("bias", fp * 3), ("N", i32 * 3), ("act", i32 * 3), ("residual", fp * 3), ("ldres", i64 * 3),
        ("out", fp), ("out_ld", i64), ("out_pstride", i64),
    ]


class SaTailDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("D", i32), ("heads", i32), ("Tmax", i32), ("hops", i32), ("t_ptr", fp), ("x", fp), ("ldx", i64),
        ("Wk", fp * 4), ("bk", fp * 4), ("Wv", fp * 4), ("bv", fp * 4), ("Wq", fp * 4), ("bq", fp * 4),
        ("Wo", fp * 4), ("bo", fp * 4), ("Wt", fp * 4), ("bt", fp * 4), ("Kc", fp * 4), ("Vc", fp * 4), ("probs", fp * 4),
        ("W_out", fp), ("b_out", fp), ("n_out", i32), ("W_stop", fp), ("b_stop", fp),
        ("mel_dst", fp), ("mel_tstride", i64), ("stop_dst", fp),
        ("tick_counter", fp), ("tick_t", fp), ("done_step", fp), ("min_iters", i32), ("use_stop", i32),
    ]


ACT = {"none": 0, None: 0, "relu": 1, "tanh": 2, "sigmoid": 3}

_lib: Optional[C.CDLL] = None

# every symbol include/satk.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "satk_last_error", "satk_version", "satk_device_info", "satk_struct_sizes", "satk_gemm",
    "satk_embedding_fwd", "satk_embedding_bwd", "satk_bn_stats", "satk_bn_apply", "satk_bn_bwd",
    "satk_highway_fwd", "satk_highway_bwd", "satk_act_bwd", "satk_mask_scale", "satk_colsum_acc", "satk_add", "satk_axpy",
    "satk_transpose", "satk_transpose_batched", "satk_transpose_strided", "satk_mask_rows", "satk_softsign_fwd", "satk_softsign_bwd", "satk_add_rowvec_tb",
    "satk_sum_over_t", "satk_bernoulli_mask", "satk_bernoulli_mask_dev", "satk_softmax_fwd", "satk_softmax_bwd", "satk_teacher_inputs",
    "satk_losses", "satk_grad_sumsq", "satk_adam_clip", "satk_l2_reg", "satk_lstm_seq_fwd", "satk_lstm_seq_bwd",
    "satk_attn_rnn_fwd", "satk_attn_rnn_bwd", "satk_attn_rnn_bwd_recurrence", "satk_attn_energy_grad", "satk_attn_energy_grad_parts", "satk_attn_rnn_bwd_overlapped", "satk_attn_energy_grad_prepare", "satk_debug_phase_cycles",
    "satk_struct_sizes_decode", "satk_rowgemm", "satk_attn_step", "satk_sa_step", "satk_sa_tail", "satk_mlp_chain", "satk_decode_tick",
]


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SatkError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.satk_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().satk_last_error().decode(errors="replace")
        raise SatkError(f"{what} failed (rc={rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL).  The tensor must be contiguous from that pointer on."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def device_info():
    out = (C.c_int * 5)()
    check(load().satk_device_info(out), "satk_device_info")
    return dict(sms=out[0], cc=(out[1], out[2]), attn_rnn_clusters=out[3], lstm_clusters=out[4])
