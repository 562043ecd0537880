"""Tensor-level wrappers over the C ABI (one per entry point of include/satk.h).

Tensors are CUDA fp32 (masks uint8, ids/lengths int64).  These wrappers only translate tensors to
raw device pointers + the current CUDA stream; all arithmetic happens in libsatk.so.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from typing import Optional

import torch

from . import lib as L
from .lib import (ACT, AttnRnnBwdDesc, AttnRnnFwdDesc, AttnStepDesc, GemmDesc, LstmBwdDesc, LstmFwdDesc, MlpChainDesc, RowGemmDesc, SaStepDesc, SaTailDesc,
                  check, load, ptr, stream_ptr)

# GEMM engine: 0 auto (tcgen05 tile when the shape allows, else SIMT), 1 SIMT fp32, 2 tcgen05 only
GEMM_ENGINE = 0
_launches = 0


def add_launches(n: int) -> None:
    """Kernel launches replayed from a CUDA graph (counted once at capture time)."""
    global _launches
    _launches += n


def launches() -> int:
    """Number of libsatk kernel-launching calls issued so far (bench.py reports the per-step delta)."""
    return _launches


def _count(n: int = 1) -> None:
    global _launches
    _launches += n


def _req(t: torch.Tensor, dtype=torch.float32) -> None:
    if not t.is_cuda:
        raise L.SatkError("satk ops need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise L.SatkError(f"expected dtype {dtype}, got {t.dtype}")


# satk_gemm_desc.precision of the products issued inside `tf32_group(name)`: 1 (single TF32 pass) for the groups named in
# SATK_TF32_1X (comma separated, or "all"), else 0 (3xTF32).  A/B instrument of tools/ab_tf32.py; the default is 3xTF32 everywhere.
_PRECISION = 0
_TF32_1X = set(x for x in os.environ.get("SATK_TF32_1X", "").split(",") if x)


@contextlib.contextmanager
def tf32_group(name: str):
    global _PRECISION
    prev = _PRECISION
    _PRECISION = 1 if (name in _TF32_1X or "all" in _TF32_1X) else prev
    try:
        yield
    finally:
        _PRECISION = prev


_PREC_STACK = []


def tf32_push(name: str) -> None:
    """Open a precision group without a `with` block (engine code marks long regions with push / pop pairs)."""
    global _PRECISION
    _PREC_STACK.append(_PRECISION)
    if name in _TF32_1X or "all" in _TF32_1X:
        _PRECISION = 1


def tf32_pop() -> None:
    global _PRECISION
    _PRECISION = _PREC_STACK.pop()


def set_tf32_1x(groups) -> None:
    global _TF32_1X
    _TF32_1X = set(groups)


def gemm(A: torch.Tensor, B: torch.Tensor, C_: torch.Tensor, M: int, N: int, K: int, *, lda: int, ldb: int, ldc: int,
         transA=False, transB=False, alpha=1.0, beta=0.0, bias=None, act=None, residual=None, ldres=0,
         keep_mask=None, keep_scale=1.0, batch1=1, batch2=1, sA=(0, 0), sB=(0, 0), sC=(0, 0),
         taps=1, shift0=0, tap_dir=1, seq_len=0, sBtap=0, shift_per_batch1=0, split_k=1, causal_skip=0,
         a_off=0, b_off=0, c_off=0, engine=None, kshift0=0, kshift_per_batch1=0,
         bank_widths=0, bank_a_kstep=0, bank_c_nstep=0, zcoord=None) -> None:
    """C = epi(alpha * sum_taps op(A) op(B)) (+beta*C).  ``*_off`` are element offsets into the tensors."""
    _req(A); _req(B); _req(C_)
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.transA, d.transB = int(transA), int(transB)
    d.A, d.lda = A.data_ptr() + 4 * a_off, lda
    d.B, d.ldb = B.data_ptr() + 4 * b_off, ldb
    d.C, d.ldc = C_.data_ptr() + 4 * c_off, ldc
    d.alpha, d.beta = alpha, beta
    d.bias = ptr(bias)
    d.act = ACT[act]
    d.residual, d.ldres = ptr(residual), ldres
    d.keep_mask, d.keep_scale = ptr(keep_mask), keep_scale
    d.batch1, d.batch2 = batch1, batch2
    d.sA1, d.sA2 = sA
    d.sB1, d.sB2 = sB
    d.sC1, d.sC2 = sC
    d.taps, d.shift0, d.tap_dir, d.seq_len, d.sBtap = taps, shift0, tap_dir, seq_len, sBtap
    d.shift_per_batch1, d.split_k, d.causal_skip = shift_per_batch1, split_k, causal_skip
    d.kshift0, d.kshift_per_batch1 = kshift0, kshift_per_batch1
    d.bank_widths, d.bank_a_kstep, d.bank_c_nstep = bank_widths, bank_a_kstep, bank_c_nstep
    if zcoord is not None:   # batched self-attention products on the tcgen05 tile (satk_gemm_desc.zcoord)
        d.zcoord = 1
        d.za_row, d.za_k, d.zb_row, d.zb_k, d.zc_col = (zcoord.get(k, 0) for k in ("za_row", "za_k", "zb_row", "zb_k", "zc_col"))
        d.a_rows, d.a_cols, d.b_rows, d.b_cols, d.c_cols = (zcoord.get(k, 0) for k in ("a_rows", "a_cols", "b_rows", "b_cols", "c_cols"))
    d.precision = _PRECISION
    check(load().satk_gemm(C.byref(d), GEMM_ENGINE if engine is None else engine, C.c_void_p(stream_ptr())), "satk_gemm")
    _count()


def linear(x: torch.Tensor, W: torch.Tensor, out: torch.Tensor, bias=None, act=None, residual=None, keep_mask=None,
           keep_scale=1.0, ldc=None, c_off=0, lda=None, a_off=0) -> torch.Tensor:
    """out[rows, N] = epi(x[rows, K] @ W[K, N] + bias); rows = x.numel() // K unless lda is given."""
    K, N = W.shape
    lda = lda or K
    rows = (x.numel() - a_off) // lda if lda != K else x.numel() // K
    gemm(x, W, out, rows, N, K, lda=lda, ldb=N, ldc=ldc or N, bias=bias, act=act, residual=residual,
         ldres=(residual.shape[-1] if residual is not None else 0), keep_mask=keep_mask, keep_scale=keep_scale,
         c_off=c_off, a_off=a_off)
    return out


def linear_dx(dy: torch.Tensor, W: torch.Tensor, dx: torch.Tensor, rows: int, beta=0.0, ldy=None, y_off=0, ldx=None,
              x_off=0, w_off=0, K=None, N=None, ldw=None) -> None:
    """dx[rows, K] (+)= dy[rows, N] @ W[K, N]^T"""
    Kw, Nw = W.shape[-2], W.shape[-1]
    K = K or Kw
    N = N or Nw
    gemm(dy, W, dx, rows, K, N, lda=ldy or N, ldb=ldw or Nw, ldc=ldx or K, transB=True, beta=beta, a_off=y_off, c_off=x_off,
         b_off=w_off)


def attn_tc_ok(T: int, dh: int) -> bool:
    """Shapes of the batched self-attention products that the tcgen05 tile takes (``attn_scores_tc`` / ``attn_apply_tc`` /
    ``attn_apply_t_tc``): a head is a k-shift (or, for an MN-major operand, a column shift) of the time-major activation, so d_head
    must be whole 32-float blocks."""
    return T >= 128 and T % 4 == 0 and dh % 32 == 0 and 32 <= dh <= 128 and os.environ.get("SATK_ATTN_TC", "1") != "0"


def attn_scores_tc(X: torch.Tensor, Y: torch.Tensor, S: torch.Tensor, T: int, nz: int, dh: int, alpha=1.0, causal=False) -> None:
    """S[z] = alpha * X_z @ Y_z^T for the nz = B*heads (utterance, head) pairs of time-major X, Y [T, nz*dh] (head z = columns
    z*dh..): QK^T of ScaledDotProductAttentionMechanism (self_attention.py:52-53) and dP = dO V^T.  S is [nz, T, T]."""
    W = nz * dh
    gemm(X, Y, S, T, T, dh, lda=W, ldb=W, ldc=T, transB=True, alpha=alpha, batch1=nz, sC=(T * T, 0),
         causal_skip=1 if causal else 0, engine=2,
         zcoord=dict(za_k=dh, zb_k=dh, a_rows=T, a_cols=W, b_rows=T, b_cols=W))


def attn_apply_tc(P: torch.Tensor, Y: torch.Tensor, Out: torch.Tensor, T: int, nz: int, dh: int, alpha=1.0, causal=False) -> None:
    """Out[:, z*dh..] = alpha * P[z] @ Y_z with P [nz, T, T] and Y time-major [T, nz*dh], read as it is (the reduction index of Y is its
    row: MN-major B operand, entry z = columns z*dh..): P.V (self_attention.py:63-65) and dQ = dS K."""
    W = nz * dh
    gemm(P, Y, Out, T, dh, T, lda=T, ldb=W, ldc=W, transB=False, alpha=alpha, batch1=nz,
         causal_skip=2 if causal else 0, engine=2,
         zcoord=dict(za_row=T, zb_row=dh, zc_col=dh, a_rows=nz * T, a_cols=T, b_rows=T, b_cols=W, c_cols=W))


def attn_apply_t_tc(P: torch.Tensor, Y: torch.Tensor, Out: torch.Tensor, T: int, nz: int, dh: int, alpha=1.0, causal=False) -> None:
    """Out[:, z*dh..] = alpha * P[z]^T @ Y_z with P the stacked [nz*T, T] matrices and Y time-major [T, nz*dh], both read as they are
    (both reduce over their row index: MN-major A and B; entry z = rows z*T.. of P, columns z*dh.. of Y): dV = P^T dO and
    dK = dS^T Q.  The last k-block of an entry runs into the next entry's rows of P; Y has no such rows and reads as zero there."""
    W = nz * dh
    gemm(P, Y, Out, T, dh, T, lda=T, ldb=W, ldc=W, transA=True, transB=False, alpha=alpha, batch1=nz,
         causal_skip=3 if causal else 0, engine=2,
         zcoord=dict(za_k=T, zb_row=dh, zc_col=dh, a_rows=nz * T, a_cols=T, b_rows=T, b_cols=W, c_cols=W))


def transposed_rows(x: torch.Tensor, rows: int, cols: int, ldx=None, x_off=0, front=0) -> torch.Tensor:
    """[cols, ld] copy of x[rows, cols] with the ROW index contiguous (ld = front + rows rounded up to 4); `front` leading and the
    trailing pad columns are zero.  Operand layout of the tcgen05 tile for products that reduce over rows (weight gradients)."""
    ld = (front + rows + 3) // 4 * 4
    pad = ld != rows
    y = (torch.zeros if pad else torch.empty)((cols, ld), device=x.device, dtype=torch.float32)
    check(load().satk_transpose_strided(C.c_void_p(x.data_ptr() + 4 * x_off), C.c_longlong(ldx or cols), rows, cols,
                                        C.c_void_p(y.data_ptr() + 4 * front), C.c_longlong(ld), C.c_void_p(stream_ptr())),
          "satk_transpose_strided")
    _count()
    return y


def conv_dw_tc(xT: torch.Tensor, drawT: torch.Tensor, dW: torch.Tensor, rows: int, cin: int, cout: int, taps: int, B: int,
               drawT_row0: int = 0) -> None:
    """dW[tap, cin, cout] += sum_r x[r + (tap - (taps-1)//2) * B, :]^T draw[r, :] for every tap in ONE tcgen05 launch.
    ``xT`` [cin, ld] and ``drawT`` [*, ld] are row-contiguous transposes (``transposed_rows``); the tap shift is a shift of the
    reduction coordinate of xT (TMA zero-fills outside the sequence), the taps are the z-batches of the launch, split-K partial
    sums are added by the TMA unit.  (tf Conv1D SAME gradient, module.py:46-68)"""
    ld = xT.shape[1]
    assert drawT.shape[1] == ld
    pl = (taps - 1) // 2
    tiles = ((cin + 127) // 128) * ((cout + 127) // 128) * taps
    kblocks = (rows + 31) // 32
    sk = max(1, min(kblocks // 4, 148 // tiles))
    gemm(xT, drawT, dW, cin, cout, rows, lda=ld, ldb=ld, ldc=cout, transB=True, b_off=drawT_row0 * ld, batch1=taps,
         sC=(cin * cout, 0), kshift0=-pl * B, kshift_per_batch1=B, split_k=sk, beta=1.0, engine=2)


DW_TC_MIN_ROWS = 2048      # below this the two transposes cost more than the SIMT split-K product saves
# weight-gradient products straight from the row-major operands (MN-major tcgen05 descriptors, gemm_tc.cu) instead of from
# transposed copies; SATK_DW_MN=0 restores the transposes
DW_MN = os.environ.get("SATK_DW_MN", "1") != "0"


def _mn_ok(x, ld, off) -> bool:
    return ld % 4 == 0 and off % 4 == 0 and x.data_ptr() % 16 == 0


def conv_dw_mn(x: torch.Tensor, draw: torch.Tensor, dW: torch.Tensor, rows: int, cin: int, cout: int, taps: int, B: int, x_ld=None,
               draw_ld=None, draw_off=0) -> None:
    """dW[tap, cin, cout] += sum_r x[r + (tap - (taps-1)//2) * B, :]^T draw[r, :] for every tap in ONE tcgen05 launch, both operands
    read as they are (row-major, the reduction index is the row: MN-major descriptors); the tap is a shift of the reduction
    coordinate of x (TMA zero-fills outside the sequence), the taps are the z-batches of the launch.  (module.py:46-68)"""
    pl = (taps - 1) // 2
    tiles = ((cin + 127) // 128) * ((cout + 127) // 128) * taps
    kblocks = (rows + 31) // 32
    sk = max(1, min(kblocks // 4, 148 // tiles))
    gemm(x, draw, dW, cin, cout, rows, lda=x_ld or cin, ldb=draw_ld or cout, ldc=cout, transA=True, b_off=draw_off, batch1=taps,
         sC=(cin * cout, 0), kshift0=-pl * B, kshift_per_batch1=B, split_k=sk, beta=1.0, engine=2)


def linear_dw(x: torch.Tensor, dy: torch.Tensor, dW: torch.Tensor, rows: int, K: int, N: int, ldx=None, x_off=0, ldy=None,
              y_off=0, ldw=None, w_off=0, shift0=0, split_k=None, xT=None, yT=None) -> None:
    """dW[K, N] += x[rows, K]^T @ dy[rows, N]  (split-K over rows, atomically accumulated).

    Large products run on the tcgen05 tile: both operands are first transposed so that the reduction (row) index is contiguous
    (``xT`` / ``yT`` may be passed in when the caller already holds a transposed copy, see ``transposed_rows``); a negative
    ``shift0`` (x delayed by -shift0 rows, zeros before the start) becomes leading zero columns of xT."""
    if (DW_MN and rows >= DW_TC_MIN_ROWS and K >= 64 and N >= 48 and split_k is None and _mn_ok(x, ldx or K, x_off)
            and _mn_ok(dy, ldy or N, y_off) and (ldw or N) % 4 == 0 and w_off % 4 == 0):
        tiles = ((K + 127) // 128) * ((N + 127) // 128)
        sk = max(1, min(rows // 128, (148 + tiles // 2) // tiles))
        gemm(x, dy, dW, K, N, rows, lda=ldx or K, ldb=ldy or N, ldc=ldw or N, transA=True, a_off=x_off, b_off=y_off, c_off=w_off,
             shift0=shift0, split_k=sk, beta=1.0, engine=2)
        return
    if rows >= DW_TC_MIN_ROWS and K >= 64 and N >= 48 and shift0 <= 0 and split_k is None:
        front = -shift0
        if xT is None:
            xT = transposed_rows(x, rows, K, ldx, x_off, front)
        if yT is None:
            yT = transposed_rows(dy, rows, N, ldy, y_off, 0)
        kk = (rows + 3) // 4 * 4
        tiles = ((K + 127) // 128) * ((N + 127) // 128)
        sk = max(1, min(kk // 128, (148 + tiles // 2) // tiles))
        gemm(xT, yT, dW, K, N, kk, lda=xT.shape[1], ldb=yT.shape[1], ldc=ldw or N, transB=True, c_off=w_off, split_k=sk, beta=1.0)
        return
    if split_k is None:
        split_k = max(1, min(64, rows // 256))
    gemm(x, dy, dW, K, N, rows, lda=ldx or K, ldb=ldy or N, ldc=ldw or N, transA=True, a_off=x_off, b_off=y_off,
         c_off=w_off, shift0=shift0, split_k=split_k, beta=1.0)
    # split_k == 1 goes through the plain epilogue with beta = 1 (accumulate), >1 through atomics: same result


def colsum_acc(x: torch.Tensor, rows: int, Ccols: int, out: torch.Tensor, ldx=None, x_off=0) -> None:
    check(load().satk_colsum_acc(C.c_void_p(x.data_ptr() + 4 * x_off), C.c_longlong(ldx or Ccols), rows, Ccols,
                                 C.c_void_p(out.data_ptr()), C.c_void_p(stream_ptr())), "satk_colsum_acc")
    _count()


def embedding_fwd(ids, table, out, offset=0):
    check(load().satk_embedding_fwd(C.c_void_p(ids.data_ptr()), ids.numel(), offset, C.c_void_p(table.data_ptr()),
                                    table.shape[1], C.c_void_p(out.data_ptr()), C.c_void_p(stream_ptr())), "satk_embedding_fwd")
    _count()


def embedding_bwd(ids, dout, dtable, offset=0):
    check(load().satk_embedding_bwd(C.c_void_p(ids.data_ptr()), ids.numel(), offset, C.c_void_p(dout.data_ptr()),
                                    dtable.shape[1], C.c_void_p(dtable.data_ptr()), C.c_void_p(stream_ptr())), "satk_embedding_bwd")
    _count()


def bn_stats(x, rows, Cc, mean, var, ldx=None, x_off=0, mov_mean=None, mov_var=None, momentum=0.99, bessel=True):
    check(load().satk_bn_stats(C.c_void_p(x.data_ptr() + 4 * x_off), C.c_longlong(ldx or Cc), rows, Cc,
                               C.c_void_p(mean.data_ptr()), C.c_void_p(var.data_ptr()), C.c_void_p(ptr(mov_mean)),
                               C.c_void_p(ptr(mov_var)), C.c_float(momentum), int(bessel), C.c_void_p(stream_ptr())), "satk_bn_stats")
    _count(4)


def bn_apply(x, rows, Cc, mean, var, gamma, beta, y, eps=1e-3, act=None, residual=None, maxpool_seq_len=0, pos_stride=1, ldx=None,
             x_off=0, ldy=None, y_off=0):
    check(load().satk_bn_apply(C.c_void_p(x.data_ptr() + 4 * x_off), C.c_longlong(ldx or Cc), rows, Cc,
                               C.c_void_p(mean.data_ptr()), C.c_void_p(var.data_ptr()), C.c_void_p(gamma.data_ptr()),
                               C.c_void_p(beta.data_ptr()), C.c_float(eps), ACT[act], C.c_void_p(ptr(residual)),
                               maxpool_seq_len, pos_stride, C.c_void_p(y.data_ptr() + 4 * y_off), C.c_longlong(ldy or Cc),
                               C.c_void_p(stream_ptr())), "satk_bn_apply")
    _count()


def bn_bwd(x, rows, Cc, mean, var, gamma, beta, dy, dx, dgamma, dbeta, scratch, eps=1e-3, act=None, maxpool_seq_len=0,
           pos_stride=1, use_batch_stats=True, ldx=None, x_off=0, lddy=None, dy_off=0, lddx=None, dx_off=0):
    check(load().satk_bn_bwd(C.c_void_p(x.data_ptr() + 4 * x_off), C.c_longlong(ldx or Cc), rows, Cc,
                             C.c_void_p(mean.data_ptr()), C.c_void_p(var.data_ptr()), C.c_void_p(gamma.data_ptr()),
                             C.c_void_p(beta.data_ptr()), C.c_float(eps), ACT[act], maxpool_seq_len, pos_stride, int(use_batch_stats),
                             C.c_void_p(dy.data_ptr() + 4 * dy_off), C.c_longlong(lddy or Cc),
                             C.c_void_p(dx.data_ptr() + 4 * dx_off), C.c_longlong(lddx or Cc),
                             C.c_void_p(dgamma.data_ptr()), C.c_void_p(dbeta.data_ptr()), C.c_void_p(scratch.data_ptr()),
                             C.c_void_p(stream_ptr())), "satk_bn_bwd")
    _count(3)


def highway_fwd(Hh, T, x, y):
    check(load().satk_highway_fwd(C.c_void_p(Hh.data_ptr()), C.c_void_p(T.data_ptr()), C.c_void_p(x.data_ptr()),
                                  C.c_void_p(y.data_ptr()), C.c_longlong(x.numel()), C.c_void_p(stream_ptr())), "satk_highway_fwd")
    _count()


def highway_bwd(Hh, T, x, dy, dH, dT, dx):
    check(load().satk_highway_bwd(C.c_void_p(Hh.data_ptr()), C.c_void_p(T.data_ptr()), C.c_void_p(x.data_ptr()),
                                  C.c_void_p(dy.data_ptr()), C.c_void_p(dH.data_ptr()), C.c_void_p(dT.data_ptr()),
                                  C.c_void_p(dx.data_ptr()), C.c_longlong(x.numel()), C.c_void_p(stream_ptr())), "satk_highway_bwd")
    _count()


def act_bwd(y, dy, dz, act, keep_mask=None, keep_scale=1.0):
    check(load().satk_act_bwd(C.c_void_p(y.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dz.data_ptr()),
                              C.c_longlong(y.numel()), ACT[act], C.c_void_p(ptr(keep_mask)), C.c_float(keep_scale),
                              C.c_void_p(stream_ptr())), "satk_act_bwd")
    _count()


def mask_scale(x, keep_mask, keep_scale, y):
    """y = keep_mask ? x * keep_scale : 0 (dropout with an explicit mask, forward and backward; include/satk.h)."""
    _req(x); _req(y)
    check(load().satk_mask_scale(C.c_void_p(x.data_ptr()), C.c_void_p(keep_mask.data_ptr()), C.c_float(keep_scale),
                                 C.c_void_p(y.data_ptr()), C.c_longlong(y.numel()), C.c_void_p(stream_ptr())), "satk_mask_scale")
    _count()


def add(a, b, out):
    check(load().satk_add(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
                          C.c_longlong(out.numel()), C.c_void_p(stream_ptr())), "satk_add")
    _count()


def axpy(alpha, x, y):
    check(load().satk_axpy(C.c_float(alpha), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_longlong(y.numel()),
                           C.c_void_p(stream_ptr())), "satk_axpy")
    _count()


def transpose(x, rows, cols, y):
    check(load().satk_transpose(C.c_void_p(x.data_ptr()), rows, cols, C.c_void_p(y.data_ptr()), C.c_void_p(stream_ptr())),
          "satk_transpose")
    _count()


def transpose_batched(src, dst, desc, n):
    check(load().satk_transpose_batched(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), C.c_void_p(desc.data_ptr()), n,
                                        C.c_void_p(stream_ptr())), "satk_transpose_batched")
    _count()


def linear_t(x: torch.Tensor, Wt: torch.Tensor, out: torch.Tensor, rows: int, K: int, N: int, *, ldw: int, k_off=0, lda=None,
             a_off=0, bias=None, act=None, residual=None, keep_mask=None, keep_scale=1.0, ldc=None, c_off=0) -> torch.Tensor:
    """out[rows, N] = epi(x[rows, K] @ W[k_off:k_off+K, :N] + bias) with the weight given TRANSPOSED: Wt is [N, ldw]
    (K-contiguous) — the operand layout of the tcgen05 tile."""
    gemm(x, Wt, out, rows, N, K, lda=lda or K, ldb=ldw, ldc=ldc or N, transB=True, b_off=k_off, a_off=a_off, c_off=c_off,
         bias=bias, act=act, residual=residual, ldres=(residual.shape[-1] if residual is not None else 0),
         keep_mask=keep_mask, keep_scale=keep_scale)
    return out


def mask_rows(x, lengths, B, T, Cc, time_major, y):
    check(load().satk_mask_rows(C.c_void_p(x.data_ptr()), C.c_void_p(lengths.data_ptr()), B, T, Cc, int(time_major),
                                C.c_void_p(y.data_ptr()), C.c_void_p(stream_ptr())), "satk_mask_rows")
    _count()


def softsign_fwd(x, y):
    check(load().satk_softsign_fwd(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_longlong(x.numel()),
                                   C.c_void_p(stream_ptr())), "satk_softsign_fwd")
    _count()


def softsign_bwd(x, dy, dx):
    check(load().satk_softsign_bwd(C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dx.data_ptr()),
                                   C.c_longlong(x.numel()), C.c_void_p(stream_ptr())), "satk_softsign_bwd")
    _count()


def add_rowvec_tb(y, v, T, B, Cc):
    check(load().satk_add_rowvec_tb(C.c_void_p(y.data_ptr()), C.c_void_p(v.data_ptr()), T, B, Cc, C.c_void_p(stream_ptr())),
          "satk_add_rowvec_tb")
    _count()


def sum_over_t(dy, T, B, Cc, dv):
    check(load().satk_sum_over_t(C.c_void_p(dy.data_ptr()), T, B, Cc, C.c_void_p(dv.data_ptr()), C.c_void_p(stream_ptr())),
          "satk_sum_over_t")
    _count()


def bernoulli_mask(out, keep_prob, seed, seed_dev=None):
    """Keep mask with P(1) = keep_prob.  With ``seed_dev`` (a 1-element int64 CUDA tensor) the seed is seed_dev[0] * 1000003 + seed, read
    on the device at run time (graph-capturable, include/satk.h)."""
    if seed_dev is not None:
        check(load().satk_bernoulli_mask_dev(C.c_void_p(out.data_ptr()), C.c_longlong(out.numel()), C.c_float(keep_prob),
                                             C.c_void_p(seed_dev.data_ptr()), C.c_ulonglong(seed), C.c_void_p(stream_ptr())),
              "satk_bernoulli_mask_dev")
        _count()
        return
    check(load().satk_bernoulli_mask(C.c_void_p(out.data_ptr()), C.c_longlong(out.numel()), C.c_float(keep_prob),
                                     C.c_ulonglong(seed), C.c_void_p(stream_ptr())), "satk_bernoulli_mask")
    _count()


def softmax_fwd(S, nmat, T, causal, keep_mask=None, keep_scale=1.0, Pd=None):
    check(load().satk_softmax_fwd(C.c_void_p(S.data_ptr()), nmat, T, int(causal), C.c_void_p(ptr(keep_mask)),
                                  C.c_float(keep_scale), C.c_void_p(ptr(Pd)), C.c_void_p(stream_ptr())), "satk_softmax_fwd")
    _count()


def softmax_bwd(P, dPd, nmat, T, causal, dS, keep_mask=None, keep_scale=1.0):
    check(load().satk_softmax_bwd(C.c_void_p(P.data_ptr()), C.c_void_p(dPd.data_ptr()), nmat, T, int(causal),
                                  C.c_void_p(ptr(keep_mask)), C.c_float(keep_scale), C.c_void_p(dS.data_ptr()),
                                  C.c_void_p(stream_ptr())), "satk_softmax_bwd")
    _count()


def teacher_inputs(mel, B, Tm, n_mels, r, n_feed, out):
    check(load().satk_teacher_inputs(C.c_void_p(mel.data_ptr()), B, Tm, n_mels, r, n_feed, C.c_void_p(out.data_ptr()),
                                     C.c_void_p(stream_ptr())), "satk_teacher_inputs")
    _count()


def losses(pred_tm, stop_tm, mel, done, spec_mask, bin_mask, B, Tm, n_mels, r, out3, dpred, dstop, scratch4):
    check(load().satk_losses(C.c_void_p(pred_tm.data_ptr()), C.c_void_p(stop_tm.data_ptr()), C.c_void_p(mel.data_ptr()),
                             C.c_void_p(done.data_ptr()), C.c_void_p(spec_mask.data_ptr()), C.c_void_p(bin_mask.data_ptr()),
                             B, Tm, n_mels, r, C.c_void_p(out3.data_ptr()), C.c_void_p(dpred.data_ptr()),
                             C.c_void_p(dstop.data_ptr()), C.c_void_p(scratch4.data_ptr()), C.c_void_p(stream_ptr())), "satk_losses")
    _count(2)


SUMSQ_SCRATCH = 640     # include/satk.h: SATK_SUMSQ_SCRATCH


def grad_sumsq(g, sumsq):
    if sumsq.numel() < SUMSQ_SCRATCH:
        raise L.SatkError(f"grad_sumsq: the scratch needs {SUMSQ_SCRATCH} floats")
    check(load().satk_grad_sumsq(C.c_void_p(g.data_ptr()), C.c_longlong(g.numel()), C.c_void_p(sumsq.data_ptr()),
                                 C.c_void_p(stream_ptr())), "satk_grad_sumsq")
    _count()


def l2_reg(p, mask, scale, g=None, loss_acc=None):
    """l2_regularization_loss (regularizers.py:11-18): loss_acc[0] += scale * sum(mask * p^2) / 2 and / or g += scale * mask * p."""
    _req(p); _req(mask)
    check(load().satk_l2_reg(C.c_void_p(p.data_ptr()), C.c_void_p(mask.data_ptr()), C.c_longlong(p.numel()), C.c_float(scale),
                             C.c_void_p(ptr(g)), C.c_void_p(ptr(loss_acc)), C.c_void_p(stream_ptr())), "satk_l2_reg")
    _count()


def adam_clip(p, g, m, v, sumsq, grad_scale, clip_norm, lr, b1, b2, eps, step):
    check(load().satk_adam_clip(C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(m.data_ptr()),
                                C.c_void_p(v.data_ptr()), C.c_longlong(p.numel()), C.c_void_p(sumsq.data_ptr()),
                                C.c_float(grad_scale), C.c_float(clip_norm), C.c_float(lr), C.c_float(b1), C.c_float(b2),
                                C.c_float(eps), int(step), C.c_void_p(stream_ptr())), "satk_adam_clip")
    _count()


def lstm_seq_fwd(xg, Wh, out, T, B, H, *, reverse=False, lengths=None, mask_c=None, mask_h=None, zc=0.0, zh=0.0,
                 forget_bias=1.0, gates=None, c_prev=None, h_prev=None, wh_off=0, ld_out=None, out_off=0):
    d = LstmFwdDesc()
    d.T, d.B, d.H, d.reverse = T, B, H, int(reverse)
    d.xg, d.Wh = xg.data_ptr(), Wh.data_ptr() + 4 * wh_off
    d.lengths, d.mask_c, d.mask_h = ptr(lengths), ptr(mask_c), ptr(mask_h)
    d.zc, d.zh, d.forget_bias = zc, zh, forget_bias
    d.out, d.gates, d.c_prev, d.h_prev = out.data_ptr() + 4 * out_off, ptr(gates), ptr(c_prev), ptr(h_prev)
    d.ld_out = ld_out or H
    check(load().satk_lstm_seq_fwd(C.byref(d), C.c_void_p(stream_ptr())), "satk_lstm_seq_fwd")
    _count()


def lstm_seq_bwd(Wh, gates, c_prev, dout, dgates, T, B, H, *, reverse=False, lengths=None, mask_c=None, mask_h=None,
                 zc=0.0, zh=0.0, wh_off=0, ld_dout=None, dout_off=0, step_end=None):
    d = LstmBwdDesc()
    d.T, d.B, d.H, d.reverse = T, B, H, int(reverse)
    d.Wh = Wh.data_ptr() + 4 * wh_off
    d.lengths, d.mask_c, d.mask_h = ptr(lengths), ptr(mask_c), ptr(mask_h)
    d.zc, d.zh = zc, zh
    d.gates, d.c_prev, d.dout, d.dgates = gates.data_ptr(), c_prev.data_ptr(), dout.data_ptr() + 4 * dout_off, dgates.data_ptr()
    d.ld_dout = ld_dout or H
    d.step_end = ptr(step_end) if (step_end is not None and lengths is None and not reverse) else None
    check(load().satk_lstm_seq_bwd(C.byref(d), C.c_void_p(stream_ptr())), "satk_lstm_seq_bwd")
    _count()


def attn_rnn_desc(**kw) -> AttnRnnFwdDesc:
    d = AttnRnnFwdDesc()
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        setattr(d, k, v)
    return d


def attn_rnn_fwd(d: AttnRnnFwdDesc):
    check(load().satk_attn_rnn_fwd(C.byref(d), C.c_void_p(stream_ptr())), "satk_attn_rnn_fwd")
    _count()


SATK_ERR_UNSUPPORTED = -3


def attn_rnn_bwd_desc(f: AttnRnnFwdDesc, **kw) -> AttnRnnBwdDesc:
    d = AttnRnnBwdDesc()
    d.f = f
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        setattr(d, k, v)
    return d


def attn_rnn_bwd(f: AttnRnnFwdDesc, **kw):
    d = attn_rnn_bwd_desc(f, **kw)
    check(load().satk_attn_rnn_bwd(C.byref(d), C.c_void_p(stream_ptr())), "satk_attn_rnn_bwd")
    _count()


def attn_rnn_bwd_launch(d: AttnRnnBwdDesc):
    check(load().satk_attn_rnn_bwd(C.byref(d), C.c_void_p(stream_ptr())), "satk_attn_rnn_bwd")
    _count()


def attn_rnn_bwd_recurrence(d: AttnRnnBwdDesc) -> bool:
    """Sequential half of the second-generation backward (include/satk.h).  False: configuration not covered, nothing launched."""
    rc = load().satk_attn_rnn_bwd_recurrence(C.byref(d), C.c_void_p(stream_ptr()))
    if rc == SATK_ERR_UNSUPPORTED:
        return False
    check(rc, "satk_attn_rnn_bwd_recurrence")
    _count()
    return True


EG_FEATURES, EG_GRADIENTS = 1, 2


def de_ws_floats(Td: int, B: int, Tt: int) -> int:
    """Size of satk_attn_rnn_bwd_desc.de_ws (include/satk.h: SATK_DE_ROW)."""
    return Td * B * (2 * ((Tt + 31) // 32 * 32) + 8 * Tt)


def eg_sync_ints(B: int) -> int:
    """Size of satk_attn_rnn_bwd_desc.sync_ws in int32 (include/satk.h: SATK_EG_SYNC_INTS)."""
    return 4 + 2 * B


EG_PREPARED = 4


def attn_energy_grad_prepare(d: AttnRnnBwdDesc) -> bool:
    """Zeroes the dkeys accumulators, the work queue and the progress flags of the overlapped pair ahead of time (include/satk.h).
    False: configuration not covered by the second-generation kernels."""
    rc = load().satk_attn_energy_grad_prepare(C.byref(d), C.c_void_p(stream_ptr()))
    if rc == SATK_ERR_UNSUPPORTED:
        return False
    check(rc, "satk_attn_energy_grad_prepare")
    return True


def attn_rnn_bwd_overlapped(d: AttnRnnBwdDesc, features: bool, prepared: bool = False) -> bool:
    """Recurrence + streaming energy gradients as a programmatic-dependent pair (include/satk.h).  False: configuration not covered
    by the second-generation kernels, nothing launched."""
    flags = (EG_FEATURES if features else 0) | (EG_PREPARED if prepared else 0)
    rc = load().satk_attn_rnn_bwd_overlapped(C.byref(d), flags, C.c_void_p(stream_ptr()))
    if rc == SATK_ERR_UNSUPPORTED:
        return False
    check(rc, "satk_attn_rnn_bwd_overlapped")
    _count(2 + int(features))
    return True


def attn_energy_grad(d: AttnRnnBwdDesc, parts: int = EG_FEATURES | EG_GRADIENTS, optional: bool = False) -> bool:
    """Parallel half of the second-generation backward; ``parts`` selects the location-feature precomputation (needs the forward pass
    only) and / or the gradient launch (include/satk.h).  ``optional``: False is returned instead of an error when the
    configuration is not covered by the second-generation kernels."""
    rc = load().satk_attn_energy_grad_parts(C.byref(d), parts, C.c_void_p(stream_ptr()))
    if optional and rc == SATK_ERR_UNSUPPORTED:
        return False
    check(rc, "satk_attn_energy_grad")
    _count(bin(parts).count("1"))
    return True


# ---------------------------------------------------------------------------------------------- free-running decode step
def rowgemm_desc(A, M, K, mats, *, lda=None, a_off=0, a_tstride=0, a_pstride=0, t_ptr=None, lstm=None) -> RowGemmDesc:
    """Descriptor of up to three skinny dense layers sharing the input rows ``A`` [M,K].

    ``mats``: list of dicts with W [K,N] (TF layout), optional bias / act / residual (+ ldres, res_off, res_tstride),
    C (+ ldc, c_off, c_tstride, c_pstride).  Offsets are in elements.  ``lstm``: dict(H, c, h, zc, zh, forget_bias, out, ld_out,
    out_off, out_pstride, hdst, ld_hdst, hdst_off, hdst_pstride) turns the epilogue of the single [K,4H] matrix into the
    ZoneoutLSTMCell pointwise update.  The tensors must stay alive while the descriptor is used."""
    _req(A)
    d = RowGemmDesc()
    d.M, d.K = M, K
    d.A, d.lda, d.a_tstride, d.a_pstride = A.data_ptr() + 4 * a_off, lda or K, a_tstride, a_pstride
    d.t_ptr = ptr(t_ptr)
    d.nmat = len(mats)
    for i, m in enumerate(mats):
        W = m["W"]
        _req(W)
        if lstm is None:
            _req(m["C"])
        if W.shape[0] != K:
            raise L.SatkError(f"rowgemm: weight {tuple(W.shape)} does not match K={K}")
        N = W.shape[1]
        d.W[i] = W.data_ptr()
        d.bias[i] = ptr(m.get("bias"))
        d.C[i] = (m["C"].data_ptr() + 4 * m.get("c_off", 0)) if m.get("C") is not None else None
        d.c_pstride[i] = m.get("c_pstride", 0)
        d.ldc[i] = m.get("ldc", N)
        d.c_tstride[i] = m.get("c_tstride", 0)
        d.N[i] = N
        d.act[i] = ACT[m.get("act")]
        res = m.get("residual")
        d.residual[i] = (res.data_ptr() + 4 * m.get("res_off", 0)) if res is not None else None
        d.ldres[i] = m.get("ldres", N)
        d.res_tstride[i] = m.get("res_tstride", 0)
    if lstm is not None:
        _req(lstm["c"]); _req(lstm["h"])
        d.lstm_H, d.lstm_c, d.lstm_h = lstm["H"], lstm["c"].data_ptr(), lstm["h"].data_ptr()
        d.zc, d.zh, d.forget_bias = lstm["zc"], lstm["zh"], lstm["forget_bias"]
        if lstm.get("out") is not None:
            d.lstm_out = lstm["out"].data_ptr() + 4 * lstm.get("out_off", 0)
            d.ld_out, d.out_pstride = lstm["ld_out"], lstm.get("out_pstride", 0)
        if lstm.get("hdst") is not None:
            d.lstm_hdst = lstm["hdst"].data_ptr() + 4 * lstm.get("hdst_off", 0)
            d.ld_hdst, d.hdst_pstride = lstm["ld_hdst"], lstm.get("hdst_pstride", 0)
    return d


def rowgemm(d: RowGemmDesc) -> None:
    check(load().satk_rowgemm(C.byref(d), C.c_void_p(stream_ptr())), "satk_rowgemm")
    _count()


def attn_step_desc(**kw) -> AttnStepDesc:
    d = AttnStepDesc()
    for k, v in kw.items():
        if torch.is_tensor(v):
            if not v.is_cuda:
                raise L.SatkError("satk ops need CUDA tensors (no CPU fallback)")
            v = v.data_ptr()
        setattr(d, k, v)
    return d


def attn_step(d: AttnStepDesc) -> None:
    check(load().satk_attn_step(C.byref(d), C.c_void_p(stream_ptr())), "satk_attn_step")
    _count()


def sa_step_desc(**kw) -> SaStepDesc:
    d = SaStepDesc()
    for k, v in kw.items():
        if torch.is_tensor(v):
            if not v.is_cuda:
                raise L.SatkError("satk ops need CUDA tensors (no CPU fallback)")
            v = v.data_ptr()
        setattr(d, k, v)
    return d


def sa_step(d: SaStepDesc) -> None:
    check(load().satk_sa_step(C.byref(d), C.c_void_p(stream_ptr())), "satk_sa_step")
    _count()


def mlp_chain_desc(x, B, K0, layers, out, *, x_ld, x_off=0, x_tstride=0, out_ld, out_off=0, out_pstride=0, t_ptr=None) -> MlpChainDesc:
    """``layers``: list of dicts W [K,N], bias, act, residual ([B,N], optional).  Offsets in elements."""
    _req(x); _req(out)
    d = MlpChainDesc()
    d.B, d.K0, d.nlayers = B, K0, len(layers)
    d.t_ptr = ptr(t_ptr)
    d.x, d.x_ld, d.x_tstride = x.data_ptr() + 4 * x_off, x_ld, x_tstride
    K = K0
    for i, m in enumerate(layers):
        W = m["W"]
        _req(W)
        if W.shape[0] != K:
            raise L.SatkError(f"mlp_chain: layer {i} weight {tuple(W.shape)} does not match input width {K}")
        d.W[i], d.bias[i], d.N[i], d.act[i] = W.data_ptr(), ptr(m.get("bias")), W.shape[1], ACT[m.get("act")]
        res = m.get("residual")
        d.residual[i], d.ldres[i] = ptr(res), (res.shape[-1] if res is not None else 0)
        K = W.shape[1]
    d.out, d.out_ld, d.out_pstride = out.data_ptr() + 4 * out_off, out_ld, out_pstride
    return d


def mlp_chain(d: MlpChainDesc) -> None:
    check(load().satk_mlp_chain(C.byref(d), C.c_void_p(stream_ptr())), "satk_mlp_chain")
    _count()


def sa_tail_desc(*, B, D, heads, Tmax, t_ptr, x, ldx, hops, W_out, b_out, W_stop, b_stop, mel_dst, mel_tstride, stop_dst,
                 tick=None) -> SaTailDesc:
    """``hops``: list of dicts with Wk,bk,Wv,bv,Wq,bq,Wo,bo,Wt,bt (weights), Kc,Vc (caches) and probs (or None)."""
    d = SaTailDesc()
    d.B, d.D, d.heads, d.Tmax, d.hops = B, D, heads, Tmax, len(hops)
    d.t_ptr, d.x, d.ldx = t_ptr.data_ptr(), x.data_ptr(), ldx
    for i, hp_ in enumerate(hops):
        for k in ("Wk", "bk", "Wv", "bv", "Wq", "bq", "Wo", "bo", "Wt", "bt", "Kc", "Vc", "probs"):
            v = hp_.get(k)
            if v is not None:
                _req(v)
            getattr(d, k)[i] = ptr(v)
    for t_ in (W_out, b_out, W_stop, b_stop, mel_dst, stop_dst):
        _req(t_)
    d.W_out, d.b_out, d.n_out = W_out.data_ptr(), b_out.data_ptr(), W_out.shape[1]
    d.W_stop, d.b_stop = W_stop.data_ptr(), b_stop.data_ptr()
    d.mel_dst, d.mel_tstride, d.stop_dst = mel_dst.data_ptr(), mel_tstride, stop_dst.data_ptr()
    if tick is not None:      # dict(counter=int32[1] zeroed, done_step=int32[1], min_iters, use_stop)
        d.tick_counter, d.tick_t, d.done_step = tick["counter"].data_ptr(), t_ptr.data_ptr(), tick["done_step"].data_ptr()
        d.min_iters, d.use_stop = int(tick["min_iters"]), int(bool(tick["use_stop"]))
    return d


def sa_tail(d: SaTailDesc) -> None:
    check(load().satk_sa_tail(C.byref(d), C.c_void_p(stream_ptr())), "satk_sa_tail")
    _count()


def decode_tick(t_dev, stop, B, min_iters, done_step) -> None:
    check(load().satk_decode_tick(C.c_void_p(ptr(t_dev)), C.c_void_p(ptr(stop)), B, min_iters, C.c_void_p(ptr(done_step)),
                                  C.c_void_p(stream_ptr())), "satk_decode_tick")
    _count()
