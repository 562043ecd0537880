"""CPU dry run of the engine's host logic: every ops.* call is replaced by a stub that checks the
extents the kernel would touch against the tensors' sizes (no arithmetic).  Catches name / shape /
offset mistakes without a GPU."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path  # noqa: E402

satk = satk_path.load()
from importlib import import_module  # noqa: E402

O = import_module("self-attention-tacotron_b200.ops")
E = import_module("self-attention-tacotron_b200.engine")
L = import_module("self-attention-tacotron_b200.lib")


def _extent(t, off, need, what):
    # view tensors: count elements available from this view's start to the end of its storage
    avail = t.untyped_storage().nbytes() // t.element_size() - t.storage_offset()
    assert off >= 0 and off + need <= avail, f"{what}: needs {off}+{need} elements, tensor view has {avail} (shape {tuple(t.shape)})"


def gemm(A, B, C_, M, N, K, *, lda, ldb, ldc, transA=False, transB=False, alpha=1.0, beta=0.0, bias=None, act=None,
         residual=None, ldres=0, keep_mask=None, keep_scale=1.0, batch1=1, batch2=1, sA=(0, 0), sB=(0, 0), sC=(0, 0), taps=1,
         shift0=0, tap_dir=1, seq_len=0, sBtap=0, shift_per_batch1=0, split_k=1, causal_skip=0, a_off=0, b_off=0, c_off=0, engine=None,
         kshift0=0, kshift_per_batch1=0, bank_widths=0, bank_a_kstep=0, bank_c_nstep=0, zcoord=None):
    if zcoord is not None:
        # batched self-attention products: shared 2-D operand views, per-entry coordinate shifts that must stay inside them
        z = dict(za_row=0, za_k=0, zb_row=0, zb_k=0, zc_col=0, c_cols=0)
        z.update(zcoord)
        nz = batch1 * batch2
        assert taps == 1 and beta == 0.0 and bias is None and residual is None and keep_mask is None
        _extent(A, a_off, (z["a_rows"] - 1) * lda + z["a_cols"], "zcoord A")
        _extent(B, b_off, (z["b_rows"] - 1) * ldb + z["b_cols"], "zcoord B")
        # an operand is stored [index rows, reduction columns] (K-major: A with transA = 0, B with transB = 1) or
        # [reduction rows, index columns] (MN-major); the shifts keep their meaning either way
        a_idx, a_red = (z["a_cols"], z["a_rows"]) if transA else (z["a_rows"], z["a_cols"])
        b_idx, b_red = (z["b_rows"], z["b_cols"]) if transB else (z["b_cols"], z["b_rows"])
        assert (nz - 1) * z["za_row"] + M <= a_idx and (nz - 1) * z["za_k"] + K <= a_red, "zcoord A shifts"
        assert (nz - 1) * z["zb_row"] + N <= b_idx and (nz - 1) * z["zb_k"] + K <= b_red, "zcoord B shifts"
        assert lda >= z["a_cols"] and ldb >= z["b_cols"] and lda % 4 == 0 and ldb % 4 == 0 and ldc % 4 == 0
        if sC[0]:
            _extent(C_, c_off, (nz - 1) * sC[0] + (M - 1) * ldc + N, "zcoord C")
        else:
            assert (nz - 1) * z["zc_col"] + N <= z["c_cols"] <= ldc, "zcoord C columns"
            _extent(C_, c_off, (M - 1) * ldc + z["c_cols"], "zcoord C")
        return
    if bank_widths:
        taps = bank_widths * (bank_widths + 1) // 2      # all kernels of the bank, back to back
    zA = (batch1 - 1) * sA[0] + (batch2 - 1) * sA[1]
    zB = (batch1 - 1) * sB[0] + (batch2 - 1) * sB[1] + (taps - 1) * sBtap
    zC = (batch1 - 1) * sC[0] + (batch2 - 1) * sC[1]
    ea = ((K - 1) * lda + M) if transA else ((M - 1) * lda + K)
    eb = ((N - 1) * ldb + K) if transB else ((K - 1) * ldb + N)
    ec = (M - 1) * ldc + N
    _extent(A, a_off, zA + ea, "gemm A")
    _extent(B, b_off, zB + eb, "gemm B")
    _extent(C_, c_off, zC + ec, "gemm C")
    if bias is not None:
        assert bias.numel() >= N, "bias"
    if residual is not None:
        _extent(residual, 0, (M - 1) * ldres + N, "gemm residual")
    if keep_mask is not None:
        assert keep_mask.numel() == M * N and keep_mask.dtype == torch.uint8, f"keep_mask {keep_mask.shape} vs {M}x{N}"
    assert (not transA and lda >= K) or (transA and lda >= M)
    assert (not transB and ldb >= N) or (transB and ldb >= K)
    assert ldc >= N


def install():
    O.gemm = gemm
    O.L.load = lambda: None

    def any_ok(*a, **k):
        return None
    for name in ["colsum_acc", "embedding_fwd", "embedding_bwd", "bn_stats", "bn_apply", "bn_bwd", "highway_fwd", "highway_bwd",
                 "act_bwd", "add", "axpy", "transpose", "mask_rows", "softsign_fwd", "softsign_bwd", "add_rowvec_tb",
                 "sum_over_t", "bernoulli_mask", "softmax_fwd", "softmax_bwd", "teacher_inputs", "losses", "grad_sumsq",
                 "adam_clip", "l2_reg", "transpose_batched", "lstm_seq_fwd", "lstm_seq_bwd", "attn_rnn_fwd", "attn_rnn_bwd",
                 "attn_rnn_bwd_recurrence", "attn_energy_grad", "attn_rnn_bwd_launch"]:
        setattr(O, name, any_ok)

    def transposed_rows(x, rows, cols, ldx=None, x_off=0, front=0):
        _extent(x, x_off, (rows - 1) * (ldx or cols) + cols, "transposed_rows x")
        return torch.empty((cols, (front + rows + 3) // 4 * 4))
    O.transposed_rows = transposed_rows
    real_desc = O.attn_rnn_desc

    def desc(**kw):
        return real_desc(**{k: (v if not isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
    O.attn_rnn_desc = desc


def run(cfg, B, Tt, Tm, overrides=None):
    hp = satk.load_hparams(os.path.join(ROOT, "examples", cfg), overrides)
    eng = E.TacotronEngine.__new__(E.TacotronEngine)
    eng.hp = hp
    eng.d = satk.dims_from_hparams(hp)
    eng.device = torch.device("cpu")
    eng.ps = satk.ParamStore(eng.d, "cpu").init(1)
    eng._bufs = {}
    eng._mask_seed = 1
    eng.saved = None
    eng.global_step = 0
    eng._sumsq = torch.zeros(1)
    eng.timers = None
    eng._side = None
    eng._sides = []
    eng.refresh_transposed()
    f, l = satk.synthetic_batch(hp, B, Tt, Tm)
    for training in (False, True):
        eng.forward(f, l, training)
        eng.backward()
        eng.optimizer_step()
    print("dry run ok:", cfg, B, Tt, Tm, overrides)


if __name__ == "__main__":
    install()
    run("ljspeech_self-attention-tacotron.json", 3, 20, 24)
    run("ljspeech_tacotron.json", 2, 21, 20)
    run("vctk_self-attention-tacotron.json", 4, 18, 16)
    run("ljspeech_tacotron.json", 2, 21, 20, "attention=additive")
    run("ljspeech_self-attention-tacotron.json", 32, 148, 800)
