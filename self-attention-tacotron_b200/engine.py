"""Host-side orchestration of the hot path: encoder, decoder, losses, backward, optimiser.

Mirrors the reference's module seam ("Module protocol"):
  * ``encoder``  : SelfAttentionCBHGEncoder.call / ZoneoutEncoderV1.call   (/root/reference/modules/module.py:425-438, 333-336)
  * ``decoder``  : DualSourceTransformerDecoder.call / ExtendedDecoder.call (module.py:1493-1559, 562-623)
  * ``model_fn`` body up to loss/optimiser (/root/reference/models/models.py:351-408, 467-498)
Every arithmetic step is a libsatk.so kernel (``ops``); torch only owns device memory and streams.

Layout: activations are TIME-MAJOR, rows ordered (t, b): a conv tap / LSTM step / teacher shift is a
row offset of B, the recurrent kernels read contiguous [B, .] slabs per step, and batched attention
uses strides (row stride B*D) instead of transposes.
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import Dict, Optional

import torch

from . import ops as O
from .data import mask_keep_prob, mask_shapes
from .params import ModelDims, ParamStore, dims_from_hparams

BN_EPS = 1e-3          # tf.layers.batch_normalization default
BN_MOMENTUM = 0.99
FORGET_BIAS = 1.0      # TF LSTMCell (A.5)
_MODE = {"additive": 0, "location_sensitive": 1, "forward": 2}


def noam_lr(init_rate: float, global_step: int, step_factor: float) -> float:
    """learning_rate_decay, /root/reference/models/models.py:595-598."""
    warm = 4000.0
    step = float(global_step * step_factor + 1)
    return init_rate * warm ** 0.5 * min(step * warm ** -1.5, step ** -0.5)


def loss_step_end(labels, Td: int, r: int) -> torch.Tensor:
    """[B] int32: 1 + the last decoder step of each utterance whose loss masks (``spec_loss_mask`` over its r frames, ``binary_loss_mask``;
    models.py:467-482) are non-zero, at least 1.  No host synchronisation."""
    B = labels.binary_loss_mask.shape[0]
    lm = (labels.binary_loss_mask != 0) | (labels.spec_loss_mask[:, :Td * r].reshape(B, Td, r) != 0).any(-1)
    return (lm * torch.arange(1, Td + 1, device=lm.device, dtype=torch.int32)).amax(1).clamp_(min=1).to(torch.int32)


def train_batch_order(step_end: Optional[torch.Tensor], source_length: torch.Tensor, Tt: int, two_level: bool = True) -> torch.Tensor:
    """Permutation of a TRAIN batch (utterances are independent, so the order is free): by target length (``step_end``), then source
    length, longest first.  The attention-RNN kernels skip positions past an utterance's source length and (backward) steps past its
    last loss step, so a cluster of 4 short utterances finishes early, and with 8 clusters on 7 cluster slots the last (shortest)
    cluster then starts earlier and is short itself.

    With more clusters than slots (B > 28) the order is two-level: the backward kernels only need the 8 shortest TARGETS in the last two
    clusters (the one that frees its slot first and the one that waits for it); inside that tail and inside the head the order is
    free, so both are ordered by SOURCE length — the forward kernel walks all Td steps of every utterance and its per-step time
    depends on the source length only: the waiting cluster and the head's last cluster get short sources."""
    B = source_length.shape[0]
    key = source_length if step_end is None else step_end.to(torch.int64) * (Tt + 1) + source_length
    perm = torch.argsort(key, descending=True, stable=True)
    if step_end is not None and B > 28 and two_level:
        head, tail = perm[:B - 8], perm[B - 8:]
        head = head[torch.argsort(source_length[head], descending=True, stable=True)]
        tail = tail[torch.argsort(source_length[tail], descending=True, stable=True)]
        perm = torch.cat([head, tail])
    return perm


class TacotronEngine:
    def __init__(self, hp, device="cuda", params: Optional[ParamStore] = None, seed: int = 1234):
        self.hp = hp
        self.d: ModelDims = dims_from_hparams(hp)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise O.L.SatkError("TacotronEngine needs a CUDA device: the product path has no CPU fallback")
        O.L.load()
        self.ps = params.to(self.device) if params is not None else ParamStore(self.d, self.device).init(seed)
        self._bufs: Dict[str, torch.Tensor] = {}
        self._mask_seed = 0x5A7C + seed
        self.saved = None
        self.global_step = 0
        self._sumsq = torch.zeros(O.SUMSQ_SCRATCH, device=self.device)   # include/satk.h: SATK_SUMSQ_SCRATCH
        self.refresh_transposed()
        # weight-gradient products have no consumer before the optimiser: they run on a second stream beside the critical path
        # (dX chain + recurrent kernels, which leave most SMs idle); SATK_WGRAD_STREAM=0 keeps everything on one stream
        self._side = torch.cuda.Stream(device=self.device) if os.environ.get("SATK_WGRAD_STREAM", "1") != "0" else None
        # ... several of them, taken round-robin by the weight-gradient sections: most of those are chains of small launches
        # (transposes, a split-K product, a column sum) that fill a fraction of the SMs, so independent sections run side by side
        # (SATK_WGRAD_LANES=1: one stream, every section behind the previous one)
        self._sides = [] if self._side is None else \
            [self._side] + [torch.cuda.Stream(device=self.device) for _ in range(max(1, int(os.environ.get("SATK_WGRAD_LANES", "3"))) - 1)]
        self._side_rr = 0
        # independent branches of the graph (decoder pre-net beside the encoder, the two BiLSTM directions) fork onto a third stream
        # (high priority like the main stream below: its branches — a BiLSTM direction, dV / dK of a self-attention block, d(memory) of
        # the second source — are joined by the critical path, and must not queue behind the weight-gradient products)
        self._aux = torch.cuda.Stream(device=self.device, priority=-1 if os.environ.get("SATK_MAIN_PRIORITY", "1") != "0" else 0) \
            if self._side is not None else None
        # ... and a fourth: the LSTM-1 / pre-net tail of the decoder's backward pass, which only the weight gradients wait for
        self._aux2 = torch.cuda.Stream(device=self.device) if self._side is not None else None
        # the critical path runs on a high-priority stream of its own, so that pending CTAs of the recurrent cluster kernels are
        # placed before pending CTAs of the weight-gradient products (SATK_MAIN_PRIORITY=0 switches it off; the caller's stream is joined on both sides)
        self._main = torch.cuda.Stream(device=self.device, priority=-1) \
            if (self._side is not None and os.environ.get("SATK_MAIN_PRIORITY", "1") != "0") else None
        if self._main is not None:
            for name in ("forward", "backward", "optimizer_step"):
                setattr(self, name, self._on_main(getattr(self, name)))
        self.use_graph = os.environ.get("SATK_GRAPH", "1") != "0"             # CUDA-graphed train step (see _train_step_graphed)
        self._graphs = {}
        self._capture_seed_dev = None
        self._seed_dev = None
        self.sort_batches = os.environ.get("SATK_SORT_BATCHES", "1") != "0"   # TRAIN steps sort the batch by target / source length
        # first element of the decoder / attention suffix of the flat parameter buffer (ParamStore lays the tensors out in forward order)
        self._dec_off = min(off for n, (off, _) in self.ps.offsets.items() if n.startswith(("att1.", "att2.", "dec.")))
        self.skip_masked_steps = os.environ.get("SATK_STEP_END", "1") != "0"  # attention-RNN backward starts at the last step with a loss
        self.timers = None     # dict name -> [(start_event, end_event)] when bench.py wants per-kernel device times

    def _timed(self, name, fn, *a, **k):
        if self.timers is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        self.timers.setdefault(name, []).append((e0, e1))
        return r

    # ------------------------------------------------------------------ helpers
    @contextlib.contextmanager
    def _wg(self, after=None):
        """Weight-gradient section: launches inside run on one of the side streams (round-robin), ordered after everything issued
        so far on the main stream and after the event ``after`` (`_wg_event` of an earlier section whose results this one reads).
        Only tensors that the main stream does not write again during this backward pass may be read here, and two sections must
        not accumulate into the same gradient tensor."""
        if self._side is None:
            yield
            return
        st = self._sides[self._side_rr % len(self._sides)]
        self._side_rr += 1
        self._last_side = st
        st.wait_stream(torch.cuda.current_stream())
        if after is not None:
            st.wait_event(after)
        with torch.cuda.stream(st):
            yield

    def _wg_event(self):
        """Event behind the weight-gradient section issued last (None without side streams)."""
        if self._side is None:
            return None
        ev = torch.cuda.Event()
        ev.record(self._last_side)
        return ev

    def _wg_join(self):
        for st in self._sides:
            torch.cuda.current_stream().wait_stream(st)

    def refresh_transposed(self):
        """K-contiguous copies of all matrices (ps.flat -> ps.flat_t), one batched launch."""
        O.transpose_batched(self.ps.flat, self.ps.flat_t, self.ps.t_desc, self.ps.t_count)

    def lin(self, x, wname, out, *, K=None, k0=0, bias=None, act=None, residual=None, keep_mask=None, keep_scale=1.0):
        """out = epi(x @ W[k0:k0+K, :] + bias) with the weight read from its K-contiguous copy (tcgen05 operand layout)."""
        Wt = self.ps.pt[wname]
        N, Ktot = Wt.shape
        K = K or Ktot
        return O.linear_t(x, Wt, out, x.numel() // K, K, N, ldw=Ktot, k_off=k0, bias=bias, act=act, residual=residual,
                          keep_mask=keep_mask, keep_scale=keep_scale)

    def _on_main(self, fn):
        def run(*a, **k):
            cur = torch.cuda.current_stream()
            if cur == self._main:
                return fn(*a, **k)
            self._main.wait_stream(cur)
            with torch.cuda.stream(self._main):
                r = fn(*a, **k)
            cur.wait_stream(self._main)
            return r
        return run

    def buf(self, name: str, shape, dtype=torch.float32, zero=False) -> torch.Tensor:
        shape = tuple(int(s) for s in shape)
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            t = torch.empty(shape, device=self.device, dtype=dtype)
            self._bufs[name] = t
        if zero:
            t.zero_()
        return t

    def _span(self, store: Dict[str, torch.Tensor], first: str, n: int) -> torch.Tensor:
        """View of n floats starting at tensor ``first`` of the flat buffer (adjacent tensors)."""
        base = self.ps.flat if store is self.ps.p else self.ps.grad
        off = self.ps.offsets[first][0]
        return base[off:off + n]

    def device_masks(self, B, Tt, Td, seed_dev=None) -> Dict[str, torch.Tensor]:
        """TRAIN-mode Bernoulli keep masks generated on the device (satk_bernoulli_mask).  ``seed_dev``: the step's seed is read from
        that device word instead of being a launch argument (the launches of a captured graph), same masks for the same seed."""
        out = {}
        for i, (name, shape) in enumerate(mask_shapes(self.d, B, Tt, Td).items()):
            m = self.buf("mask." + name, shape, torch.uint8)
            if seed_dev is not None:
                O.bernoulli_mask(m, mask_keep_prob(self.d, name), i, seed_dev=seed_dev)
            else:
                O.bernoulli_mask(m, mask_keep_prob(self.d, name), self._mask_seed * 1000003 + i)
            out[name] = m
        if seed_dev is None:
            self._mask_seed += 1
        return out

    @contextlib.contextmanager
    def _fork(self, which=0):
        """Independent branch: runs on an auxiliary stream after everything issued so far; `_join` merges it back."""
        aux = getattr(self, "_aux2" if which else "_aux", None)
        if aux is None:
            yield
            return
        aux.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(aux):
            yield

    def _join(self, which=0):
        aux = getattr(self, "_aux2" if which else "_aux", None)
        if aux is not None:
            torch.cuda.current_stream().wait_stream(aux)

    # ------------------------------------------------------------------ self-attention block
    def _sa_forward(self, x, T, B, name, heads, causal, mask, keep, key):
        """SelfAttentionTransformer.call (module.py:363-371) on time-major x [T*B, D]."""
        p = self.ps.p
        D = x.shape[1]
        dh = D // heads
        R = T * B
        Kp = self.lin(x, name + ".key.W", self.buf(key + ".K", (R, D)), bias=p[name + ".key.b"])
        Vp = self.lin(x, name + ".value.W", self.buf(key + ".V", (R, D)), bias=p[name + ".value.b"])
        Qp = self.lin(x, name + ".query.W", self.buf(key + ".Q", (R, D)), bias=p[name + ".query.b"])
        S = self.buf(key + ".P", (B, heads, T, T))
        # QK^T / P.V: z-batches of the tcgen05 tile when a head is whole k-blocks (decoder: d_head = 128), else the SIMT tile
        tc = O.attn_tc_ok(T, dh)
        if tc:
            O.attn_scores_tc(Qp, Kp, S, T, B * heads, dh, alpha=1.0 / math.sqrt(dh), causal=causal)
        else:
            O.gemm(Qp, Kp, S, T, T, dh, lda=B * D, ldb=B * D, ldc=T, transB=True, alpha=1.0 / math.sqrt(dh),
                   batch1=B, batch2=heads, sA=(D, dh), sB=(D, dh), sC=(heads * T * T, T * T), causal_skip=1 if causal else 0)
        Pd = self.buf(key + ".Pd", (B, heads, T, T)) if mask is not None else None
        O.softmax_fwd(S, B * heads, T, causal, mask, 1.0 / keep, Pd)
        Pu = Pd if Pd is not None else S
        Oc = self.buf(key + ".O", (R, D))
        if tc:
            O.attn_apply_tc(Pu, Vp, Oc, T, B * heads, dh, causal=causal)
        else:
            O.gemm(Pu, Vp, Oc, T, dh, T, lda=T, ldb=B * D, ldc=B * D, batch1=B, batch2=heads,
                   sA=(heads * T * T, T * T), sB=(D, dh), sC=(D, dh), causal_skip=2 if causal else 0)
        ao = self.lin(Oc, name + ".output.W", self.buf(key + ".ao", (R, D)), bias=p[name + ".output.b"])
        tr = self.lin(ao, name + ".transform.W", self.buf(key + ".tr", (R, D)), bias=p[name + ".transform.b"], act="tanh")
        y = self.buf(key + ".y", (R, D))
        O.add(x, tr, y)
        return y, dict(x=x, K=Kp, V=Vp, Q=Qp, P=S, Pd=Pu, O=Oc, ao=ao, tr=tr, T=T, heads=heads, causal=causal,
                       mask=mask, keep=keep, name=name, key=key)

    def _sa_backward(self, sv, dy, B):
        p, g = self.ps.p, self.ps.g
        name, key, T, heads, causal = sv["name"], sv["key"], sv["T"], sv["heads"], sv["causal"]
        x = sv["x"]
        R, D = x.shape
        dh = D // heads
        scale = 1.0 / math.sqrt(dh)
        dz = self.buf(key + ".dz", (R, D))
        O.act_bwd(sv["tr"], dy, dz, "tanh")
        with self._wg():
            O.linear_dw(sv["ao"], dz, g[name + ".transform.W"], R, D, D)
            O.colsum_acc(dz, R, D, g[name + ".transform.b"])
        dao = self.buf(key + ".dao", (R, D))
        O.linear_dx(dz, p[name + ".transform.W"], dao, R)
        with self._wg():
            O.linear_dw(sv["O"], dao, g[name + ".output.W"], R, D, D)
            O.colsum_acc(dao, R, D, g[name + ".output.b"])
        dO = self.buf(key + ".dO", (R, D))
        O.linear_dx(dao, p[name + ".output.W"], dO, R)
        bs = dict(batch1=B, batch2=heads)
        sP, sX = (heads * T * T, T * T), (D, dh)
        dPd = self.buf(key + ".dPd", (B, heads, T, T))
        dV = self.buf(key + ".dV", (R, D))
        dS = self.buf(key + ".dS", (B, heads, T, T))
        dQ = self.buf(key + ".dQ", (R, D))
        dK = self.buf(key + ".dK", (R, D))
        if O.attn_tc_ok(T, dh):
            # the four gradient products on the tcgen05 tile; operands that are reduced over their row index are read as they are
            # (MN-major descriptors; the stacked [z][T][T] matrices as ONE [z*T, T] matrix, entry z = rows z*T..)
            # dV does not depend on d(scores), dK and dQ only share dS: the dV and dK branches run on the auxiliary stream
            nz, W = B * heads, B * D
            with self._fork():
                O.attn_apply_t_tc(sv["Pd"], dO, dV, T, nz, dh, causal=causal)
            O.attn_scores_tc(dO, sv["V"], dPd, T, nz, dh, causal=causal)
            O.softmax_bwd(sv["P"], dPd, nz, T, causal, dS, sv["mask"], 1.0 / sv["keep"])
            with self._fork():
                O.attn_apply_t_tc(dS, sv["Q"], dK, T, nz, dh, alpha=scale, causal=causal)
            O.attn_apply_tc(dS, sv["K"], dQ, T, nz, dh, alpha=scale, causal=causal)
            self._join()
        else:
            O.gemm(dO, sv["V"], dPd, T, T, dh, lda=B * D, ldb=B * D, ldc=T, transB=True, sA=sX, sB=sX, sC=sP,
                   causal_skip=1 if causal else 0, **bs)
            O.gemm(sv["Pd"], dO, dV, T, dh, T, lda=T, ldb=B * D, ldc=B * D, transA=True, sA=sP, sB=sX, sC=sX, **bs)
            O.softmax_bwd(sv["P"], dPd, B * heads, T, causal, dS, sv["mask"], 1.0 / sv["keep"])
            O.gemm(dS, sv["K"], dQ, T, dh, T, lda=T, ldb=B * D, ldc=B * D, alpha=scale, sA=sP, sB=sX, sC=sX,
                   causal_skip=2 if causal else 0, **bs)
            O.gemm(dS, sv["Q"], dK, T, dh, T, lda=T, ldb=B * D, ldc=B * D, transA=True, alpha=scale, sA=sP, sB=sX, sC=sX, **bs)
        dx = self.buf(key + ".dx", (R, D))
        first = True
        with self._wg():
            for nm, dt in (("query", dQ), ("key", dK), ("value", dV)):
                O.linear_dw(x, dt, g[f"{name}.{nm}.W"], R, D, D)
                O.colsum_acc(dt, R, D, g[f"{name}.{nm}.b"])
        for nm, dt in (("query", dQ), ("key", dK), ("value", dV)):
            if first:   # dx = dy (residual branch) + dQ.Wq^T
                O.gemm(dt, p[f"{name}.{nm}.W"], dx, R, D, D, lda=D, ldb=D, ldc=D, transB=True, residual=dy, ldres=D)
                first = False
            else:
                O.linear_dx(dt, p[f"{name}.{nm}.W"], dx, R, beta=1.0)
        return dx

    def _bn_forward(self, xraw, R, Cc, first, gname, act, out, training, residual=None, maxpool_T=0, B=1):
        """tf.layers.batch_normalization (+ activation, + max-pool over time) of xraw [R, Cc]: batch statistics in training (the moving
        ones are updated), moving statistics otherwise.  -> what `_bn_backward` needs."""
        gamma = self._span(self.ps.p, gname + ".gamma", Cc)
        beta = self._span(self.ps.p, gname + ".beta", Cc)
        o = self.ps.bn_off[first]
        mm, mv = self.ps.bn_mean_flat[o:o + Cc], self.ps.bn_var_flat[o:o + Cc]
        if training:
            mean, var = self.buf(first + ".bmean", (Cc,)), self.buf(first + ".bvar", (Cc,))
            O.bn_stats(xraw, R, Cc, mean, var, mov_mean=mm, mov_var=mv, momentum=BN_MOMENTUM, bessel=True)
        else:
            mean, var = mm, mv
        O.bn_apply(xraw, R, Cc, mean, var, gamma, beta, out, eps=BN_EPS, act=act, residual=residual,
                   maxpool_seq_len=maxpool_T, pos_stride=B)
        return dict(x=xraw, C=Cc, mean=mean, var=var, gamma=gamma, beta=beta, act=act, maxpool=maxpool_T > 0, gname=gname)

    def _bn_backward(self, s, R, dy, dx, scratch, training, T, B):
        first = s["gname"]
        dgamma = self._span(self.ps.g, first + ".gamma", s["C"])
        dbeta = self._span(self.ps.g, first + ".beta", s["C"])
        O.bn_bwd(s["x"], R, s["C"], s["mean"], s["var"], s["gamma"], s["beta"], dy, dx, dgamma, dbeta, scratch, eps=BN_EPS,
                 act=s["act"], maxpool_seq_len=T if s["maxpool"] else 0, pos_stride=B, use_batch_stats=training)

    def _conv_backward(self, R, B, x, Wname, k, cin, cout, draw, dx, beta, x_ld=None, draw_ld=None, draw_off=0, residual=None,
                       xT=None, drawT=None, drawT_row0=0, skip_dx=False, after=None):
        """Backward of a SAME conv1d over time-major rows.  draw: gradient wrt the raw conv output [R, cout] (ld draw_ld);
        accumulates dW (weight-gradient stream), writes / accumulates dx."""
        p, g = self.ps.p, self.ps.g
        pl = (k - 1) // 2
        with self._wg(after):
            if O.DW_MN and R >= O.DW_TC_MIN_ROWS and cin >= 64 and cout >= 48 and (x_ld or cin) % 4 == 0 and \
                    (draw_ld or cout) % 4 == 0 and draw_off % 4 == 0:
                # all taps in one tcgen05 launch straight from x and draw (the tap is a shift of the reduction coordinate)
                O.conv_dw_mn(x, draw, g[Wname], R, cin, cout, k, B, x_ld, draw_ld, draw_off)
            elif R >= O.DW_TC_MIN_ROWS and cin % 128 == 0 and cout >= 48:
                # all taps in one tcgen05 launch on row-contiguous transposes (the tap is a shift of the reduction coordinate)
                if xT is None:
                    xT = O.transposed_rows(x, R, cin, x_ld)
                if drawT is None:
                    drawT, drawT_row0 = O.transposed_rows(draw, R, cout, draw_ld, draw_off), 0
                O.conv_dw_tc(xT, drawT, g[Wname], R, cin, cout, k, B, drawT_row0)
            else:
                O.gemm(x, draw, g[Wname], cin, cout, R, lda=x_ld or cin, ldb=draw_ld or cout, ldc=cout, transA=True,
                       b_off=draw_off, batch1=k, sC=(cin * cout, 0), shift0=-pl * B, shift_per_batch1=B,
                       split_k=max(1, min(32, R // 512)), beta=1.0)
        if not skip_dx:
            O.gemm(draw, p[Wname], dx, R, cin, cout, lda=draw_ld or cout, ldb=cout, ldc=cin, transB=True, a_off=draw_off,
                   taps=k, shift0=pl * B, tap_dir=-B, sBtap=cin * cout, beta=beta, residual=residual, ldres=cin)

    def _bank_one_launch(self, R, cin, C):
        """The conv bank runs as one z-batched tcgen05 launch when its tiles are whole (128-wide convs over 128 channels)."""
        return getattr(self, "bank_one_launch", True) and C == 128 and cin % 32 == 0 and R >= 64

    # ------------------------------------------------------------------ encoder
    def encoder(self, source, source_length, training, masks):
        """-> (lstm_output [Tt,B,2H], self_attention_output [Tt,B,32] | None, [alignments])."""
        d, p = self.d, self.ps.p
        B, Tt = source.shape
        R = Tt * B
        sv = {}
        ids_tm = source.t().contiguous()
        sv["ids_tm"] = ids_tm
        x = self.buf("enc.emb", (R, d.embed))
        O.embedding_fwd(ids_tm, p["embedding"], x)
        sv["prenet_in"] = [x]
        for i, u in enumerate(d.enc_prenet):
            y = self.buf(f"enc.p{i}", (R, u))
            self.lin(x, f"enc.prenet{i}.W", y, bias=p[f"enc.prenet{i}.b"], act="relu",
                     keep_mask=masks[f"enc.prenet{i}"] if training else None, keep_scale=1.0 / (1.0 - d.enc_prenet_drop))
            x = y
            sv["prenet_in"].append(x)
        inp = x
        C, K, cin = d.conv_ch, d.bank_k, inp.shape[1]
        KC = C * K
        raw = self.buf("enc.bank_raw", (R, KC))
        if self._bank_one_launch(R, cin, C):
            # all K widths as z-batches of ONE tcgen05 launch: the K kernels are stored back to back ([K(K+1)/2 taps][C][cin])
            o = self.ps.offsets["cbhg.bank1.W"][0]
            Wt_all = self.ps.flat_t[o:o + (K * (K + 1) // 2) * C * cin]
            O.gemm(inp, Wt_all, raw, R, C, cin, lda=cin, ldb=cin, ldc=KC, transB=True, tap_dir=B, sBtap=cin * C, bank_widths=K,
                   bank_c_nstep=C, engine=2)
        else:
            for k in range(1, K + 1):
                pl = (k - 1) // 2
                O.gemm(inp, self.ps.pt[f"cbhg.bank{k}.W"], raw, R, C, cin, lda=cin, ldb=cin, ldc=KC, c_off=(k - 1) * C, transB=True,
                       taps=k, shift0=-pl * B, tap_dir=B, sBtap=cin * C)

        def bn(xraw, Cc, first, gname, act, out, residual=None, maxpool=False):
            return self._bn_forward(xraw, R, Cc, first, gname, act, out, training, residual, Tt if maxpool else 0, B)

        mp = self.buf("enc.mp", (R, KC))
        sv["bn_bank"] = bn(raw, KC, "cbhg.bank1", "cbhg.bank1", "relu", mp, maxpool=True)
        raw1 = self.buf("enc.raw1", (R, d.proj1))
        O.gemm(mp, self.ps.pt["cbhg.proj1.W"], raw1, R, d.proj1, KC, lda=KC, ldb=KC, ldc=d.proj1, transB=True, taps=3, shift0=-B,
               tap_dir=B, sBtap=KC * d.proj1)
        p1o = self.buf("enc.p1o", (R, d.proj1))
        sv["bn_p1"] = bn(raw1, d.proj1, "cbhg.proj1", "cbhg.proj1", "relu", p1o)
        raw2 = self.buf("enc.raw2", (R, d.proj2))
        O.gemm(p1o, self.ps.pt["cbhg.proj2.W"], raw2, R, d.proj2, d.proj1, lda=d.proj1, ldb=d.proj1, ldc=d.proj2, transB=True, taps=3,
               shift0=-B, tap_dir=B, sBtap=d.proj1 * d.proj2)
        hw = self.buf("enc.hw0", (R, d.proj2))
        sv["bn_p2"] = bn(raw2, d.proj2, "cbhg.proj2", "cbhg.proj2", None, hw, residual=inp)
        sv.update(inp=inp, mp=mp, p1o=p1o)
        Hn = d.enc_lstm
        sv["hwy"] = []
        for i in range(d.n_highway):
            Hb = self.lin(hw, f"cbhg.highway{i}.WH", self.buf(f"enc.hwH{i}", (R, Hn)), bias=p[f"cbhg.highway{i}.bH"], act="relu")
            Tb = self.lin(hw, f"cbhg.highway{i}.WT", self.buf(f"enc.hwT{i}", (R, Hn)), bias=p[f"cbhg.highway{i}.bT"], act="sigmoid")
            y = self.buf(f"enc.hw{i + 1}", (R, Hn))
            O.highway_fwd(Hb, Tb, hw, y)
            sv["hwy"].append((Hb, Tb, hw))
            hw = y
        sv["lstm_in"] = hw
        mem1 = self.buf("enc.mem1", (Tt, B, 2 * Hn))
        sv["lstm"] = {}
        for j, dr in ((1, "bw"), (0, "fw")):
            # the two directions are independent (disjoint column halves of mem1): the backward one runs on the auxiliary stream
            # (forked first: a fork issued after the forward direction would queue behind it)
            with (self._fork() if j == 1 else contextlib.nullcontext()):
                W = p[f"cbhg.lstm_{dr}.W"]
                xg = self.lin(hw, f"cbhg.lstm_{dr}.W", self.buf(f"enc.xg_{dr}", (R, 4 * Hn)), K=Hn, bias=p[f"cbhg.lstm_{dr}.b"])
                gates = self.buf(f"enc.gates_{dr}", (R, 4 * Hn))
                cp = self.buf(f"enc.cprev_{dr}", (R, Hn))
                hp_ = self.buf(f"enc.hprev_{dr}", (R, Hn))
                mc = masks[f"cbhg.lstm_{dr}.c"] if training else None
                mh = masks[f"cbhg.lstm_{dr}.h"] if training else None
                O.lstm_seq_fwd(xg, W[Hn:], mem1, Tt, B, Hn, reverse=(j == 1), lengths=source_length, mask_c=mc, mask_h=mh,
                               zc=d.zc, zh=d.zh, forget_bias=FORGET_BIAS, gates=gates, c_prev=cp, h_prev=hp_,
                               ld_out=2 * Hn, out_off=j * Hn)
                sv["lstm"][dr] = dict(gates=gates, c_prev=cp, h_prev=hp_, mc=mc, mh=mh)
        self._join()
        mem2, aligns = None, []
        if d.dual:
            x2 = self.lin(mem1, "enc.sa_proj.W", self.buf("enc.sa_in", (R, d.enc_sa)), bias=p["enc.sa_proj.b"])
            sv["sa"] = []
            for h in range(d.enc_sa_hops):
                mk = masks[f"enc.sa{h}"] if training else None
                x2, s = self._sa_forward(x2, Tt, B, f"enc.sa{h}", d.enc_sa_heads, False, mk, 1.0 - d.enc_sa_drop, f"enc.sa{h}")
                sv["sa"].append(s)
                aligns += [s["P"][:, i] for i in range(d.enc_sa_heads)]
            mem2 = x2
        self._enc_saved = sv
        return mem1, mem2, aligns

    def encoder_backward(self, dmem1, dmem2, B, Tt, source_length):
        d, p, g, sv = self.d, self.ps.p, self.ps.g, self._enc_saved
        R = Tt * B
        Hn = d.enc_lstm
        if d.dual:
            dx = dmem2
            for h in reversed(range(d.enc_sa_hops)):
                dx = self._sa_backward(sv["sa"][h], dx, B)
            mem1 = self._bufs["enc.mem1"]
            with self._wg():
                O.linear_dw(mem1, dx, g["enc.sa_proj.W"], R, 2 * Hn, d.enc_sa)
                O.colsum_acc(dx, R, d.enc_sa, g["enc.sa_proj.b"])
            O.linear_dx(dx, p["enc.sa_proj.W"], dmem1, R, beta=1.0)
        dhw = self.buf("enc.dhw", (R, Hn))
        dgs = {}
        for j, dr in ((1, "bw"), (0, "fw")):     # backward direction forked first, forward direction on the main stream
            with (self._fork() if j == 1 else contextlib.nullcontext()):
                s = sv["lstm"][dr]
                W = p[f"cbhg.lstm_{dr}.W"]
                dg = dgs[dr] = self.buf(f"enc.dgates_{dr}", (R, 4 * Hn))
                O.lstm_seq_bwd(W[Hn:], s["gates"], s["c_prev"], dmem1, dg, Tt, B, Hn, reverse=(j == 1), lengths=source_length,
                               mask_c=s["mc"], mask_h=s["mh"], zc=d.zc, zh=d.zh, ld_dout=2 * Hn, dout_off=j * Hn)
                gW = g[f"cbhg.lstm_{dr}.W"]
                with self._wg():
                    O.linear_dw(sv["lstm_in"], dg, gW, R, Hn, 4 * Hn)
                    O.linear_dw(s["h_prev"], dg, gW, R, Hn, 4 * Hn, w_off=Hn * 4 * Hn)
                    O.colsum_acc(dg, R, 4 * Hn, g[f"cbhg.lstm_{dr}.b"])
        O.linear_dx(dgs["fw"], p["cbhg.lstm_fw.W"][:Hn], dhw, R, beta=0.0)
        self._join()
        O.linear_dx(dgs["bw"], p["cbhg.lstm_bw.W"][:Hn], dhw, R, beta=1.0)
        for i in reversed(range(d.n_highway)):
            Hb, Tb, x = sv["hwy"][i]
            # per-layer gradient buffers: the weight-gradient stream still reads them while the next layer is computed
            dH, dT = self.buf(f"enc.dH{i}", (R, Hn)), self.buf(f"enc.dT{i}", (R, Hn))
            dxd = self.buf(f"enc.dhw_in{i % 2}", (R, Hn))
            O.highway_bwd(Hb, Tb, x, dhw, dH, dT, dxd)
            with self._wg():
                O.linear_dw(x, dH, g[f"cbhg.highway{i}.WH"], R, Hn, Hn)
                O.colsum_acc(dH, R, Hn, g[f"cbhg.highway{i}.bH"])
                O.linear_dw(x, dT, g[f"cbhg.highway{i}.WT"], R, Hn, Hn)
                O.colsum_acc(dT, R, Hn, g[f"cbhg.highway{i}.bT"])
            O.linear_dx(dH, p[f"cbhg.highway{i}.WH"], dxd, R, beta=1.0)
            O.linear_dx(dT, p[f"cbhg.highway{i}.WT"], dxd, R, beta=1.0)
            dhw = dxd
        # dhw = gradient wrt (proj2_bn + inp)
        scratch = self.buf("enc.bn_scratch", (2 * d.conv_ch * d.bank_k,))
        training = self._training

        def bn_back(s, dy, dx):
            self._bn_backward(s, R, dy, dx, scratch, training, Tt, B)

        def conv_back(x, Wname, k, cin, cout, draw, dx, beta, **kw):
            self._conv_backward(R, B, x, Wname, k, cin, cout, draw, dx, beta, **kw)

        draw2 = self.buf("enc.draw2", (R, d.proj2))
        bn_back(sv["bn_p2"], dhw, draw2)
        dp1o = self.buf("enc.dp1o", (R, d.proj1))
        conv_back(sv["p1o"], "cbhg.proj2.W", 3, d.proj1, d.proj2, draw2, dp1o, 0.0)
        draw1 = self.buf("enc.draw1", (R, d.proj1))
        bn_back(sv["bn_p1"], dp1o, draw1)
        KC = d.conv_ch * d.bank_k
        dmp = self.buf("enc.dmp", (R, KC))
        conv_back(sv["mp"], "cbhg.proj1.W", 3, KC, d.proj1, draw1, dmp, 0.0)
        draw = self.buf("enc.dbank_raw", (R, KC))
        bn_back(sv["bn_bank"], dmp, draw)
        cin = sv["inp"].shape[1]
        dinp = self.buf("enc.dinp", (R, cin))
        bank_tc = R >= O.DW_TC_MIN_ROWS and cin % 128 == 0 and d.conv_ch >= 48
        with self._wg():
            inpT = O.transposed_rows(sv["inp"], R, cin) if (bank_tc and not O.DW_MN) else None      # shared by all bank widths
            drawT = O.transposed_rows(draw, R, KC) if (bank_tc and not O.DW_MN) else None
        bank_ev = self._wg_event() if bank_tc else None     # the per-width sections below run on other side streams
        one = bank_tc and self._bank_one_launch(R, cin, d.conv_ch)
        if one:
            # dinp = dhw (residual branch, module.py:86) + sum over widths and taps of the transposed convs: ONE launch, every width a
            # z-batch that reads its 128-column block of `draw` and adds its partial sum through the TMA unit
            dinp.copy_(dhw)
            o = self.ps.offsets["cbhg.bank1.W"][0]
            W_all = self.ps.flat[o:o + (d.bank_k * (d.bank_k + 1) // 2) * cin * d.conv_ch]
            O.gemm(draw, W_all, dinp, R, cin, d.conv_ch, lda=KC, ldb=d.conv_ch, ldc=cin, transB=True, tap_dir=-B, sBtap=cin * d.conv_ch,
                   beta=1.0, bank_widths=d.bank_k, bank_a_kstep=d.conv_ch, engine=2)
        for k in range(1, d.bank_k + 1):   # weight gradients per width (+ the per-width input gradient when not fused above)
            conv_back(sv["inp"], f"cbhg.bank{k}.W", k, cin, d.conv_ch, draw, dinp, 0.0 if k == 1 else 1.0, draw_ld=KC,
                      draw_off=(k - 1) * d.conv_ch, residual=dhw if k == 1 else None, xT=inpT, drawT=drawT,
                      drawT_row0=(k - 1) * d.conv_ch, skip_dx=one, after=bank_ev)
        # encoder pre-net
        dy = dinp
        n = len(d.enc_prenet)
        for i in reversed(range(n)):
            y = sv["prenet_in"][i + 1]
            x = sv["prenet_in"][i]
            u, cin_i = y.shape[1], x.shape[1]
            dz = self.buf(f"enc.dz{i}", (R, u))
            O.act_bwd(y, dy, dz, "relu", None, 1.0 / (1.0 - d.enc_prenet_drop) if training else 1.0)
            with self._wg():
                O.linear_dw(x, dz, g[f"enc.prenet{i}.W"], R, cin_i, u)
                O.colsum_acc(dz, R, u, g[f"enc.prenet{i}.b"])
            dxp = self.buf(f"enc.dpx{i}", (R, cin_i))
            O.linear_dx(dz, p[f"enc.prenet{i}.W"], dxp, R)
            dy = dxp
        O.embedding_bwd(sv["ids_tm"], dy, g["embedding"])

    # ------------------------------------------------------------------ decoder
    def decoder_pre(self, target, speaker_embed, training, masks):
        """Teacher inputs, decoder pre-net and the input projection of LSTM-1: dense over all steps and independent of the encoder
        (helpers.py:13-55; module.py:1509-1511; multi_speaker_modules.py:27-32), so `forward` runs it beside the encoder."""
        d, p = self.d, self.ps.p
        B, Tm = target.shape[0], target.shape[1]
        Td = Tm // d.r
        Rd = Td * B
        sv = {}
        dec_in = self.buf("dec.in", (Rd, d.dec_in))
        O.teacher_inputs(target, B, Tm, d.n_mels, d.r, d.n_feed, dec_in)
        keep = 1.0 - d.dec_prenet_drop
        m0 = masks["dec.prenet0"] if training else None
        m1 = masks["dec.prenet1"] if training else None
        if d.use_speaker:
            h0 = self.lin(dec_in, "dec.prenet0.W0", self.buf("dec.ph0", (Rd, d.dec_prenet[0])), bias=p["dec.prenet0.b0"], act="relu")
            sp_pre = self.lin(speaker_embed, "dec.prenet0.Ws", self.buf("dec.sp_pre", (B, d.dec_prenet[0])), bias=p["dec.prenet0.bs"])
            sp = self.buf("dec.sp", (B, d.dec_prenet[0]))
            O.softsign_fwd(sp_pre, sp)
            O.add_rowvec_tb(h0, sp, Td, B, d.dec_prenet[0])
            dp0 = self.lin(h0, "dec.prenet0.W", self.buf("dec.p0", (Rd, d.dec_prenet[0])), bias=p["dec.prenet0.b"], act="relu",
                           keep_mask=m0, keep_scale=1.0 / keep)
            sv.update(h0=h0, sp_pre=sp_pre)
        else:
            dp0 = self.lin(dec_in, "dec.prenet0.W", self.buf("dec.p0", (Rd, d.dec_prenet[0])), bias=p["dec.prenet0.b"], act="relu",
                           keep_mask=m0, keep_scale=1.0 / keep)
        dp1 = self.lin(dp0, "dec.prenet1.W", self.buf("dec.p1", (Rd, d.dec_prenet[1])), bias=p["dec.prenet1.b"], act="relu",
                       keep_mask=m1, keep_scale=1.0 / keep)
        O.tf32_push("lstmx")
        xg1 = self.lin(dp1, "dec.lstm1.W", self.buf("dec.xg1", (Rd, 4 * d.att_rnn)), K=d.dec_prenet[1], bias=p["dec.lstm1.b"])
        O.tf32_pop()
        sv.update(dec_in=dec_in, dp0=dp0, dp1=dp1, xg1=xg1)
        return sv

    def decoder(self, mem1, mem2, source_length, target, speaker_embed, training, masks, pre=None):
        """Teacher-forced decoder.  -> (mel_tm [Td,B,r*n_mels], stop_tm [Td,B], align1, align2, dec self-attn P)."""
        d, p = self.d, self.ps.p
        Tt, B, _ = mem1.shape
        Tm = target.shape[1]
        Td = Tm // d.r
        R, Rd = Tt * B, Td * B
        sv = dict(pre if pre is not None else self.decoder_pre(target, speaker_embed, training, masks))
        dec_in, dp0, dp1, xg1 = sv["dec_in"], sv["dp0"], sv["dp1"], sv.pop("xg1")
        H1, HD, P1 = d.att_rnn, d.dec_out, d.dec_prenet[1]
        W1 = p["dec.lstm1.W"]
        # attention memories (BahdanauAttention.__init__, A.8): values masked past length, keys = values.W_mem
        values1 = self.buf("dec.values1", (R, d.mem1))
        O.mask_rows(mem1, source_length, B, Tt, d.mem1, True, values1)
        O.tf32_push("mem")
        keys1 = self.lin(values1, "att1.memory.W", self.buf("dec.keys1", (R, d.att1)))
        O.tf32_pop()
        values2 = keys2 = None
        if d.dual:
            values2 = self.buf("dec.values2", (R, d.mem2))
            O.mask_rows(mem2, source_length, B, Tt, d.mem2, True, values2)
            O.tf32_push("mem")
            keys2 = self.lin(values2, "att2.memory.W", self.buf("dec.keys2", (R, d.att2)))
            O.tf32_pop()
        X2W = H1 + d.ctx
        x2 = self.buf("dec.x2", (Rd, X2W))
        al1 = self.buf("dec.align1", (Td, B, Tt))
        al2 = self.buf("dec.align2", (Td, B, Tt)) if d.dual else None
        loc = d.attention in ("forward", "location_sensitive")
        agent = d.attention == "forward" and d.transition_agent          # forward_attention.py:111-114
        fd = O.attn_rnn_desc(
            Td=Td, B=B, Tt=Tt, H=H1, A1=d.att1, A2=d.att2, M1=d.mem1, M2=d.mem2,
            att_kernel=d.att_kernel if loc else 0, mode=_MODE[d.attention], cumulative=int(d.cumulative),
            xg=xg1, Wrec=W1[P1:], mask_c=masks["dec.lstm1.c"] if training else None,
            mask_h=masks["dec.lstm1.h"] if training else None, zc=d.zc, zh=d.zh, forget_bias=FORGET_BIAS,
            lengths=source_length, keys1=keys1, values1=values1, Wq1=p["att1.query.W"], v1=p["att1.v"],
            b1=p["att1.b"] if loc else None,
            loc_conv_w=p["att1.loc_conv.W"] if loc else None, loc_conv_b=p["att1.loc_conv.b"] if loc else None,
            loc_layer_w=p["att1.loc_layer.W"] if loc else None, att_filters=d.att_filters if loc else 0,
            keys2=keys2, values2=values2, Wq2=p["att2.query.W"] if d.dual else None, v2=p["att2.v"] if d.dual else None,
            x2=x2, align1=al1, align2=al2,
            gates=self.buf("dec.gates1", (Rd, 4 * H1)), c_prev=self.buf("dec.cprev1", (Rd, H1)),
            h_prev=self.buf("dec.hprev1", (Rd, H1)), soft1=self.buf("dec.soft1", (Td, B, Tt)),
            q_save=self.buf("dec.qsave", (Rd, d.att1 + d.att2)),
            agent_w=p["att1.agent.W"] if agent else None, agent_b=p["att1.agent.b"] if agent else None,
            u_save=self.buf("dec.usave", (Td, B)) if agent else None,
            state_final=self.buf("dec.state_final", (B, Tt)) if (loc and d.cumulative) else None)
        self._timed("attn_rnn_fwd", O.attn_rnn_fwd, fd)
        sv.update(fd=fd, dec_in=dec_in, dp0=dp0, dp1=dp1, x2=x2, values1=values1, values2=values2, keys1=keys1, keys2=keys2)
        # LSTM-2, LSTM-3 (DecoderRNNV2)
        x = x2
        sv["lstm"] = []
        for li, kin in ((2, X2W), (3, HD)):
            W = p[f"dec.lstm{li}.W"]
            O.tf32_push("lstmx")
            xg = self.lin(x, f"dec.lstm{li}.W", self.buf(f"dec.xg{li}", (Rd, 4 * HD)), K=kin, bias=p[f"dec.lstm{li}.b"])
            O.tf32_pop()
            out = self.buf(f"dec.out{li}", (Rd, HD))
            gates = self.buf(f"dec.gates{li}", (Rd, 4 * HD))
            cp, hp_ = self.buf(f"dec.cprev{li}", (Rd, HD)), self.buf(f"dec.hprev{li}", (Rd, HD))
            mc = masks[f"dec.lstm{li}.c"] if training else None
            mh = masks[f"dec.lstm{li}.h"] if training else None
            O.lstm_seq_fwd(xg, W[kin:], out, Td, B, HD, mask_c=mc, mask_h=mh, zc=d.zc, zh=d.zh, forget_bias=FORGET_BIAS,
                           gates=gates, c_prev=cp, h_prev=hp_)
            sv["lstm"].append(dict(x=x, kin=kin, gates=gates, c_prev=cp, h_prev=hp_, mc=mc, mh=mh, li=li))
            x = out
        sa_P = []
        sv["sa"] = []
        if d.dual:
            for h in range(d.dec_sa_hops):
                mk = masks[f"dec.sa{h}"] if training else None
                O.tf32_push("sa")
                x, s = self._sa_forward(x, Td, B, f"dec.sa{h}", d.dec_sa_heads, True, mk, 1.0 - d.dec_sa_drop, f"dec.sa{h}")
                O.tf32_pop()
                sv["sa"].append(s)
                sa_P += [s["P"][:, i] for i in range(d.dec_sa_heads)]
        sv["proj_in"] = x
        O.tf32_push("proj")
        mel_tm = self.lin(x, "dec.out_proj.W", self.buf("dec.mel_tm", (Rd, d.out_units)), bias=p["dec.out_proj.b"])
        O.tf32_pop()
        O.tf32_push("proj")
        stop_tm = self.lin(x, "dec.stop_proj.W", self.buf("dec.stop_tm", (Rd, 1)), bias=p["dec.stop_proj.b"])
        O.tf32_pop()
        self._dec_saved = sv
        return mel_tm, stop_tm, al1, al2, sa_P

    def decoder_backward(self, dmel_tm, dstop_tm, B, Tt, Td, source_length):
        """-> (dmem1 [Tt,B,mem1], dmem2 | None)"""
        d, p, g, sv = self.d, self.ps.p, self.ps.g, self._dec_saved
        training = self._training
        R, Rd = Tt * B, Td * B
        H1, HD = d.att_rnn, d.dec_out
        # second-generation attention backward: its location features need the forward pass only -> first thing on the auxiliary stream
        de_ws = self.buf("dec.de_ws", (O.de_ws_floats(Td, B, Tt),)) if d.dual else None
        sync_ws = self.buf("dec.eg_sync", (O.eg_sync_ints(B),), torch.int32) if d.dual else None
        feats_early = False
        if d.dual and getattr(self, "_aux", None) is not None and os.environ.get("SATK_ENERGY_FORK", "1") != "0":
            with self._fork():
                feats_early = O.attn_energy_grad(O.attn_rnn_bwd_desc(sv["fd"], de_ws=de_ws), O.EG_FEATURES, optional=True)
        # ... and the few memsets of the overlapped pair now, so that nothing sits between the producer of dx2 and the recurrence
        # launch (its clusters need whole SMs: a weight-gradient product that becomes ready in that gap takes them for 0.1 ms)
        prepared = False
        if feats_early and getattr(self, "_aux2", None) is not None and self.timers is None and os.environ.get("SATK_EG_OVERLAP", "1") != "0":
            prepared = O.attn_energy_grad_prepare(O.attn_rnn_bwd_desc(
                sv["fd"], de_ws=de_ws, sync_ws=sync_ws, dkeys1=self.buf("dec.dkeys1", (Tt * B, d.att1)),
                dkeys2=self.buf("dec.dkeys2", (Tt * B, d.att2)), dv1=g["att1.v"], dv2=g["att2.v"],
                dloc_conv_w=g.get("att1.loc_conv.W"), dloc_conv_b=g.get("att1.loc_conv.b"), dloc_layer_w=g.get("att1.loc_layer.W")))
        O.tf32_push("proj")
        x = sv["proj_in"]
        Dp = x.shape[1]
        with self._wg():
            O.linear_dw(x, dmel_tm, g["dec.out_proj.W"], Rd, Dp, d.out_units)
            O.colsum_acc(dmel_tm, Rd, d.out_units, g["dec.out_proj.b"])
            O.linear_dw(x, dstop_tm, g["dec.stop_proj.W"], Rd, Dp, 1)
            O.colsum_acc(dstop_tm, Rd, 1, g["dec.stop_proj.b"])
        dx = self.buf("dec.dproj_in", (Rd, Dp))
        O.linear_dx(dmel_tm, p["dec.out_proj.W"], dx, Rd)
        O.linear_dx(dstop_tm, p["dec.stop_proj.W"], dx, Rd, beta=1.0)
        O.tf32_pop()
        O.tf32_push("sa")
        for h in reversed(range(len(sv["sa"]))):
            dx = self._sa_backward(sv["sa"][h], dx, B)
        O.tf32_pop()
        O.tf32_push("lstmx")
        dout = dx
        for s in reversed(sv["lstm"]):
            li, kin = s["li"], s["kin"]
            W = p[f"dec.lstm{li}.W"]
            dg = self.buf(f"dec.dgates{li}", (Rd, 4 * HD))
            O.lstm_seq_bwd(W[kin:], s["gates"], s["c_prev"], dout, dg, Td, B, HD, mask_c=s["mc"], mask_h=s["mh"], zc=d.zc, zh=d.zh,
                           step_end=self.saved.get("step_end"))
            gW = g[f"dec.lstm{li}.W"]
            with self._wg():
                dgT = O.transposed_rows(dg, Rd, 4 * HD) if (Rd >= O.DW_TC_MIN_ROWS and not O.DW_MN) else None
                O.linear_dw(s["x"], dg, gW, Rd, kin, 4 * HD, yT=dgT)
                O.linear_dw(s["h_prev"], dg, gW, Rd, HD, 4 * HD, w_off=kin * 4 * HD, yT=dgT)
                del dgT
                O.colsum_acc(dg, Rd, 4 * HD, g[f"dec.lstm{li}.b"])
            dxl = self.buf(f"dec.dx_lstm{li}", (Rd, kin))
            O.linear_dx(dg, W[:kin], dxl, Rd)
            dout = dxl
        dx2 = dout                                     # [Rd, H1 + ctx]
        fd = sv["fd"]
        loc = d.attention in ("forward", "location_sensitive")
        QT = d.att1 + d.att2
        dg1 = self.buf("dec.dgates1", (Rd, 4 * H1))
        dq = self.buf("dec.dq", (Rd, QT))
        dkeys1 = self.buf("dec.dkeys1", (R, d.att1))
        dkeys2 = self.buf("dec.dkeys2", (R, d.att2)) if d.dual else None
        bd = O.attn_rnn_bwd_desc(
            fd, dx2=dx2, dgates=dg1, dq=dq, dkeys1=dkeys1, dkeys2=dkeys2,
            dv1=g["att1.v"], dv2=g["att2.v"] if d.dual else None,
            dloc_conv_w=g["att1.loc_conv.W"] if loc else None, dloc_conv_b=g["att1.loc_conv.b"] if loc else None,
            dloc_layer_w=g["att1.loc_layer.W"] if loc else None,
            dagent_w=g["att1.agent.W"] if (d.attention == "forward" and d.transition_agent) else None,
            dagent_b=g["att1.agent.b"] if (d.attention == "forward" and d.transition_agent) else None,
            step_end=self.saved.get("step_end"),
            # workspace of the second-generation kernels (d(energies) of both mechanisms, include/satk.h)
            de_ws=de_ws, sync_ws=sync_ws)
        # second generation: the recurrence, and the energy gradients (dkeys, dv, location layer / conv) as a parallel launch of
        # their own; configurations it does not cover run the first-generation kernel (everything in one launch).
        # Default: the overlapped pair — the gradient workers start beside the recurrence (programmatic dependent launch) and follow
        # its progress, so only their last chunks are left when it ends (SATK_EG_OVERLAP=0, or per-kernel timers: one after the other)
        energy_forked = False
        overlapped = (d.dual and getattr(self, "_aux2", None) is not None and self.timers is None
                      and os.environ.get("SATK_EG_OVERLAP", "1") != "0")
        if overlapped:
            if feats_early:
                self._join()                   # location features (auxiliary stream) before the pair
            overlapped = O.attn_rnn_bwd_overlapped(bd, not feats_early, prepared)
        if overlapped:
            pass
        elif d.dual and self._timed("attn_rnn_bwd", O.attn_rnn_bwd_recurrence, bd):
            # dkeys / dv / d(location layer, conv) feed nothing before the memory-layer gradients below: the launch runs on the
            # auxiliary stream beside the LSTM-1 input gradient, the dvalues products and the pre-net backward chain
            if getattr(self, "_aux", None) is not None and os.environ.get("SATK_ENERGY_FORK", "1") != "0":
                with self._fork():
                    self._timed("attn_energy_grad", O.attn_energy_grad, bd, O.EG_GRADIENTS if feats_early else O.EG_FEATURES | O.EG_GRADIENTS)
                energy_forked = True
            else:
                self._timed("attn_energy_grad", O.attn_energy_grad, bd)
        else:
            self._timed("attn_rnn_bwd", O.attn_rnn_bwd_launch, bd)
        with (self._fork(1) if overlapped else contextlib.nullcontext()):
            # nothing below feeds the encoder: with the overlapped pair it runs on its own stream beside the memory-layer gradients
            # and the encoder's backward pass (`backward` joins it)
            self._lstm1_prenet_backward(sv, dg1, dq, B, Td, training)
        O.tf32_push("mem")
        if energy_forked:
            self._join()
        dmem1, dmem2 = self._memory_backward(sv, dx2, dkeys1, dkeys2, B, Tt, Td, loc, source_length)
        O.tf32_pop()
        return dmem1, dmem2

    def _lstm1_prenet_backward(self, sv, dg1, dq, B, Td, training):
        """Weight gradients of LSTM-1 and the query layers, then the decoder pre-net backwards (module.py:1509-1511,
        multi_speaker_modules.py:27-32).  Reads d(gates) / d(queries) of the attention-RNN backward; feeds weight gradients only."""
        d, p, g = self.d, self.ps.p, self.ps.g
        Rd = Td * B
        H1, P1 = d.att_rnn, d.dec_prenet[1]
        X2W = H1 + d.ctx
        QT = d.att1 + d.att2
        # LSTM-1 weight gradients (dense over time)
        gW1 = g["dec.lstm1.W"]
        N4 = 4 * H1
        tc_dw = Rd >= O.DW_TC_MIN_ROWS
        with self._wg():
            dg1T = O.transposed_rows(dg1, Rd, N4) if (tc_dw and not O.DW_MN) else None     # shared by the three row blocks of dec.lstm1.W
            O.linear_dw(sv["dp1"], dg1, gW1, Rd, P1, N4, yT=dg1T)
            # context rows: input of step t is the context of step t-1 (zero at t=0) -> shift by one time step (B rows)
            O.linear_dw(sv["x2"], dg1, gW1, Rd, d.ctx, N4, ldx=X2W, x_off=H1, w_off=P1 * N4, shift0=-B, yT=dg1T)
            O.linear_dw(self._bufs["dec.hprev1"], dg1, gW1, Rd, H1, N4, w_off=(P1 + d.ctx) * N4, yT=dg1T)
            del dg1T
            O.colsum_acc(dg1, Rd, N4, g["dec.lstm1.b"])
            # query layers: q = out1 . Wq  (out1 = x2[:, :H1])
            O.linear_dw(sv["x2"], dq, g["att1.query.W"], Rd, H1, d.att1, ldx=X2W, ldy=QT)
            if d.dual:
                O.linear_dw(sv["x2"], dq, g["att2.query.W"], Rd, H1, d.att2, ldx=X2W, ldy=QT, y_off=d.att1)
        ddp1 = self.buf("dec.ddp1", (Rd, P1))
        O.linear_dx(dg1, p["dec.lstm1.W"][:P1], ddp1, Rd)
        O.tf32_push("prenet")
        # decoder pre-net
        keep_scale = 1.0 / (1.0 - d.dec_prenet_drop) if training else 1.0
        dz1 = self.buf("dec.dz1", (Rd, P1))
        O.act_bwd(sv["dp1"], ddp1, dz1, "relu", None, keep_scale)
        P0 = d.dec_prenet[0]
        with self._wg():
            O.linear_dw(sv["dp0"], dz1, g["dec.prenet1.W"], Rd, P0, P1)
            O.colsum_acc(dz1, Rd, P1, g["dec.prenet1.b"])
        ddp0 = self.buf("dec.ddp0", (Rd, P0))
        O.linear_dx(dz1, p["dec.prenet1.W"], ddp0, Rd)
        dz0 = self.buf("dec.dz0", (Rd, P0))
        O.act_bwd(sv["dp0"], ddp0, dz0, "relu", None, keep_scale)
        if d.use_speaker:
            h0 = sv["h0"]
            with self._wg():
                O.linear_dw(h0, dz0, g["dec.prenet0.W"], Rd, P0, P0)
                O.colsum_acc(dz0, Rd, P0, g["dec.prenet0.b"])
            dh0 = self.buf("dec.dh0", (Rd, P0))
            O.linear_dx(dz0, p["dec.prenet0.W"], dh0, Rd)
            # h0 = relu(in.W0+b0) + softsign(spk.Ws+bs): the relu output is h0 - sp, recomputed on the fly
            dsp = self.buf("dec.dsp", (B, P0))
            O.sum_over_t(dh0, Td, B, P0, dsp)
            dsp_pre = self.buf("dec.dsp_pre", (B, P0))
            O.softsign_bwd(sv["sp_pre"], dsp, dsp_pre)
            self._dspk_pre = dsp_pre
            relu0 = self.buf("dec.relu0", (Rd, P0))
            self.lin(sv["dec_in"], "dec.prenet0.W0", relu0, bias=p["dec.prenet0.b0"], act="relu")
            dz00 = self.buf("dec.dz00", (Rd, P0))
            O.act_bwd(relu0, dh0, dz00, "relu")
            with self._wg():
                O.linear_dw(sv["dec_in"], dz00, g["dec.prenet0.W0"], Rd, d.dec_in, P0)
                O.colsum_acc(dz00, Rd, P0, g["dec.prenet0.b0"])
        else:
            with self._wg():
                O.linear_dw(sv["dec_in"], dz0, g["dec.prenet0.W"], Rd, d.dec_in, P0)
                O.colsum_acc(dz0, Rd, P0, g["dec.prenet0.b"])
        O.tf32_pop()

    def _memory_backward(self, sv, dx2, dkeys1, dkeys2, B, Tt, Td, loc, source_length):
        """d(memory) of both sources: through the contexts (dx2[:, H1:] holds d(context) of every step) and through the keys
        (keys = values . W_mem, attention bias folded in), masked past the source lengths; weight gradients of the memory layers.
        The two sources are independent: the second one runs on the auxiliary stream.  -> (dmem1, dmem2 | None)"""
        d, p, g = self.d, self.ps.p, self.ps.g
        R = Tt * B
        H1 = d.att_rnn
        X2W = H1 + d.ctx
        with self._wg():
            if loc:
                O.colsum_acc(dkeys1, R, d.att1, g["att1.b"])
            O.linear_dw(sv["values1"], dkeys1, g["att1.memory.W"], R, d.mem1, d.att1)
            if d.dual:
                O.linear_dw(sv["values2"], dkeys2, g["att2.memory.W"], R, d.mem2, d.att2)
        dmem2 = None
        if d.dual:
            with self._fork():
                dval2 = self.buf("dec.dvalues2", (R, d.mem2))
                O.gemm(self._bufs["dec.align2"], dx2, dval2, Tt, d.mem2, Td, lda=B * Tt, ldb=B * X2W, ldc=B * d.mem2, transA=True,
                       b_off=H1 + d.mem1, batch1=B, sA=(Tt, 0), sB=(X2W, 0), sC=(d.mem2, 0))
                O.linear_dx(dkeys2, p["att2.memory.W"], dval2, R, beta=1.0)
                dmem2 = self.buf("dec.dmem2", (Tt, B, d.mem2))
                O.mask_rows(dval2, source_length, B, Tt, d.mem2, True, dmem2)
        # values: dvalues[j,b,:] = sum_t align[t,b,j] * dctx_total[t,b,:]
        dval1 = self.buf("dec.dvalues1", (R, d.mem1))
        O.gemm(self._bufs["dec.align1"], dx2, dval1, Tt, d.mem1, Td, lda=B * Tt, ldb=B * X2W, ldc=B * d.mem1, transA=True,
               b_off=H1, batch1=B, sA=(Tt, 0), sB=(X2W, 0), sC=(d.mem1, 0))
        O.linear_dx(dkeys1, p["att1.memory.W"], dval1, R, beta=1.0)
        dmem1 = self.buf("dec.dmem1", (Tt, B, d.mem1))
        O.mask_rows(dval1, source_length, B, Tt, d.mem1, True, dmem1)
        if d.dual:
            self._join()
        return dmem1, dmem2

    # ------------------------------------------------------------------ model_fn body
    # ------------------------------------------------------------------ PostNetV2 (models/models.py:92-100,440-462)
    def _frames_of(self, x_tm, Td, B, name):
        """decoder rows [Td*B, r*n_mels] (r frames per row) -> frame-major rows [T_mel*B, n_mels] (frame f = t*r + i)."""
        d = self.d
        out = self.buf(name, (Td, d.r, B, d.n_mels))
        out.copy_(x_tm.view(Td, B, d.r, d.n_mels).permute(0, 2, 1, 3))
        return out.view(Td * d.r * B, d.n_mels)

    def _rows_of(self, x_fm, Td, B, name):
        """frame-major rows [T_mel*B, n_mels] -> decoder rows [Td*B, r*n_mels]."""
        d = self.d
        out = self.buf(name, (Td, B, d.r, d.n_mels))
        out.copy_(x_fm.view(Td, d.r, B, d.n_mels).permute(0, 2, 1, 3))
        return out.view(Td * B, d.r * d.n_mels)

    def postnet(self, mel_tm, Td, B, training, masks, key="post"):
        """PostNetV2 on the decoder's mel output: conv1d (SAME, no bias) -> BN -> tanh (none on the last layer) -> dropout, x
        num_postnet_v2_layers, Dense(num_mels), residual (tacotron2 PostNetV2, RECALLED; oracle/model.py: postnet_v2).
        The convolutions run over FRAMES, so the rows are re-ordered frame-major first.  -> postnet output as decoder rows."""
        d, p = self.d, self.ps.p
        R, C, k = Td * d.r * B, d.postnet_ch, d.postnet_kernel
        pl = (k - 1) // 2
        x0 = self._frames_of(mel_tm, Td, B, key + ".x0")
        sv = dict(x=[x0], bn=[], Td=Td, B=B, training=training, masks=masks)
        x, cin = x0, d.n_mels
        O.tf32_push("postnet")
        for i in range(d.postnet_layers):
            raw = self.buf(f"{key}.raw{i}", (R, C))
            O.gemm(x, self.ps.pt[f"postnet.conv{i}.W"], raw, R, C, cin, lda=cin, ldb=cin, ldc=C, transB=True, taps=k, shift0=-pl * B,
                   tap_dir=B, sBtap=cin * C)
            a = self.buf(f"{key}.a{i}", (R, C))
            act = "tanh" if i < d.postnet_layers - 1 else None
            sv["bn"].append(self._bn_forward(raw, R, C, f"postnet.conv{i}", f"postnet.conv{i}", act, a, training))
            if training:
                O.mask_scale(a, masks[f"postnet.conv{i}"], 1.0 / (1.0 - d.postnet_drop), a)
            sv["x"].append(a)
            x, cin = a, C
        out = self.buf(key + ".out", (R, d.n_mels))
        self.lin(x, "postnet.proj.W", out, bias=p["postnet.proj.b"], residual=x0)
        O.tf32_pop()
        self._post_saved = sv
        return self._rows_of(out, Td, B, key + ".out_tm")

    def postnet_backward(self, dpost_tm, dmel_tm):
        """Backward of `postnet`: weight gradients, and d(postnet output) through the residual and the convolutions added onto
        d(mel output) (decoder rows)."""
        d, p, g, sv = self.d, self.ps.p, self.ps.g, self._post_saved
        Td, B, training, masks = sv["Td"], sv["B"], sv["training"], sv["masks"]
        R, C, k = Td * d.r * B, d.postnet_ch, d.postnet_kernel
        dout = self._frames_of(dpost_tm, Td, B, "post.dout")
        O.tf32_push("postnet")
        xL = sv["x"][-1]
        with self._wg():
            O.linear_dw(xL, dout, g["postnet.proj.W"], R, C, d.n_mels)
            O.colsum_acc(dout, R, d.n_mels, g["postnet.proj.b"])
        dx = self.buf("post.dx_last", (R, C))
        O.linear_dx(dout, p["postnet.proj.W"], dx, R)
        scratch = self.buf("post.bn_scratch", (2 * C,))
        for i in reversed(range(d.postnet_layers)):
            if training:
                O.mask_scale(dx, masks[f"postnet.conv{i}"], 1.0 / (1.0 - d.postnet_drop), dx)
            draw = self.buf(f"post.draw{i}", (R, C))
            self._bn_backward(sv["bn"][i], R, dx, draw, scratch, training, 0, B)
            cin = d.n_mels if i == 0 else C
            dxi = self.buf(f"post.dx{i % 2}", (R, cin)) if i > 0 else self.buf("post.dx0", (R, cin))
            self._conv_backward(R, B, sv["x"][i], f"postnet.conv{i}.W", k, cin, C, draw, dxi, 0.0)
            dx = dxi
        O.tf32_pop()
        # residual branch + convolution branch, back in decoder rows
        O.axpy(1.0, dpost_tm, dmel_tm)
        O.axpy(1.0, self._rows_of(dx, Td, B, "post.dx0_tm"), dmel_tm)

    def forward(self, features, labels, training: bool, masks: Optional[Dict[str, torch.Tensor]] = None):
        """Forward pass + losses (+ loss gradients wrt the predictions).  Inputs must be CUDA tensors."""
        d = self.d
        source, source_length = features.source, features.source_length
        B, Tt = source.shape
        Tm = labels.mel.shape[1]
        Td = Tm // d.r
        perm = None
        # 1 + the last decoder step of each utterance whose loss masks are non-zero (models.py:467-482): later steps carry exactly
        # zero gradient (masked losses, causal decoder), so the attention-RNN backward kernel starts its walk there
        step_end = loss_step_end(labels, Td, d.r) if training else None
        if training and getattr(self, "sort_batches", True) and B > 1:
            perm = train_batch_order(step_end, source_length, Tt, os.environ.get("SATK_SORT_TWO_LEVEL", "1") != "0")
            step_end = step_end.index_select(0, perm) if step_end is not None else None
            sel = lambda x: x.index_select(0, perm) if torch.is_tensor(x) else x      # noqa: E731
            features = features._replace(source=sel(source), source_length=sel(source_length), speaker_id=sel(features.speaker_id))
            labels = labels._replace(mel=sel(labels.mel), target_length=sel(labels.target_length), done=sel(labels.done),
                                     spec_loss_mask=sel(labels.spec_loss_mask), binary_loss_mask=sel(labels.binary_loss_mask))
            source, source_length = features.source, features.source_length
            if masks is not None:      # caller-provided keep masks follow their utterances (batch is dim 0 of the [B,heads,T,T] masks)
                masks = {k: v.index_select(0 if ".sa" in k else 1, perm) for k, v in masks.items()}
        if training and masks is None:
            # (under graph capture the mask launches are part of the graph and read the step's seed from a device word)
            masks = self.device_masks(B, Tt, Td, seed_dev=getattr(self, "_capture_seed_dev", None))
        # descriptors saved for the backward pass hold raw device pointers: the (possibly re-ordered) inputs stay referenced until then
        self._keepalive = (features, labels, masks)
        self._training = training
        spk = None
        if d.use_speaker:
            spk = self.buf("spk_embed", (B, d.speaker_dim))
            O.embedding_fwd(features.speaker_id, self.ps.p["speaker_embedding"], spk, offset=d.speaker_offset)
        with self._fork():      # pre-net + LSTM-1 input projection do not depend on the encoder
            O.tf32_push("prenet")
            pre = self.decoder_pre(labels.mel, spk, training, masks)
            O.tf32_pop()
        O.tf32_push("enc")
        mem1, mem2, enc_al = self._timed("sec.encoder_fwd", self.encoder, source, source_length, training, masks)
        O.tf32_pop()
        # (the encoder forks / joins the auxiliary stream itself for the BiLSTM directions, which also orders `pre` before here)
        self._join()
        mel_tm, stop_tm, al1, al2, dec_sa = self._timed("sec.decoder_fwd", self.decoder, mem1, mem2, source_length, labels.mel, spk,
                                                        training, masks, pre)
        out3 = self.buf("loss3", (3,))
        dmel = self.buf("dec.dmel_tm", mel_tm.shape)
        dstop = self.buf("dec.dstop_tm", stop_tm.shape)
        O.losses(mel_tm, stop_tm, labels.mel, labels.done, labels.spec_loss_mask, labels.binary_loss_mask, B, Tm, d.n_mels, d.r,
                 out3, dmel, dstop, self.buf("loss_scratch", (4,)))
        if d.l2_weight > 0:     # loss = mel_loss + done_loss + regularization_loss (models/models.py:482)
            l2 = self.buf("l2_loss", (1,), zero=True)
            O.l2_reg(self.ps.flat, self.ps.l2_mask(), d.l2_weight, loss_acc=l2)
            O.add(out3[2:], l2, out3[2:])
        post_tm, dpost, post3 = None, None, None
        if d.postnet_v2:        # + postnet_v2_mel_loss (models/models.py:116-118, 479-482)
            post_tm = self.postnet(mel_tm, Td, B, training, masks)
            post3 = self.buf("post.loss3", (3,))
            dpost = self.buf("post.dpost_tm", post_tm.shape)
            O.losses(post_tm, stop_tm, labels.mel, labels.done, labels.spec_loss_mask, labels.binary_loss_mask, B, Tm, d.n_mels, d.r,
                     post3, dpost, self.buf("post.dstop_unused", stop_tm.shape), self.buf("post.loss_scratch", (4,)))
            O.add(out3[2:], post3[:1], out3[2:])
        self.saved = dict(B=B, Tt=Tt, Td=Td, Tm=Tm, source_length=source_length, dmel=dmel, dstop=dstop, features=features,
                          dpost=dpost,
                          # (PostNetV2 spreads gradient onto the masked frames — conv taps, batch statistics —: no step is skipped)
                          step_end=step_end if (getattr(self, "skip_masked_steps", True) and not d.postnet_v2) else None)
        # with a sorted batch every per-utterance output is in SORTED order: row i belongs to utterance perm[i] of the caller's batch
        return dict(mel_tm=mel_tm, stop_tm=stop_tm, align1_tm=al1, align2_tm=al2, enc_self_P=enc_al, dec_self_P=dec_sa,
                    memory1_tm=mem1, memory2_tm=mem2, losses=out3, perm=perm, mel_postnet_tm=post_tm,
                    postnet_v2_mel_loss=None if post3 is None else post3[:1])

    # ------------------------------------------------------------------ free-running decode (PREDICT)
    def _build_decode_step(self, B, Tt, Tmax, use_stop_token, min_iters, forced=False):
        """Descriptors + launch sequence of ONE free-running decoder step (module.py:762-778, rnn_wrappers.py:47-124,188-214).
        All tensors are persistent buffers and every kernel reads the step index from ``t_dev``, so the returned closure can
        be captured once in a CUDA graph and replayed."""
        d, p = self.d, self.ps.p
        H1, HD, P1, P0 = d.att_rnn, d.dec_out, d.dec_prenet[1], d.dec_prenet[0]
        CTX, OU = d.ctx, d.out_units
        X2W = H1 + CTX
        W1C, W2C, W3C = P1 + CTX + H1, X2W + HD, 2 * HD
        b_ = self.buf
        t_dev = b_("pred.t", (1,), torch.int32)
        done = b_("pred.done", (1,), torch.int32)
        mel_hist = b_("pred.mel_hist", (Tmax + 1, B, OU))
        stop_hist = b_("pred.stop_hist", (Tmax, B))
        # concatenated recurrent input rows, double-buffered on the parity of t (see satk_rowgemm_desc.a_pstride)
        cell_in, x2c, x3c = b_("pred.cell_in", (2, B, W1C)), b_("pred.x2cat", (2, B, W2C)), b_("pred.x3cat", (2, B, W3C))
        st = {n: b_("pred." + n, (B, H1 if n.endswith("1") else HD)) for n in ("c1", "h1", "c2", "h2", "c3", "h3")}
        q = b_("pred.q", (B, d.att1 + d.att2))
        o3 = b_("pred.o3", (B, HD))
        pp0 = b_("pred.pp0", (B, P0))
        aprev, alpha, u = b_("pred.aprev", (B, Tt)), b_("pred.alpha", (B, Tt)), b_("pred.u", (B,))
        al1 = b_("pred.align1", (Tmax, B, Tt))
        al2 = b_("pred.align2", (Tmax, B, Tt)) if d.dual else None
        lengths = b_("pred.lengths", (B,), torch.int64)
        loc = d.attention in ("forward", "location_sensitive")
        zo = dict(zc=d.zc, zh=d.zh, forget_bias=FORGET_BIAS)
        steps = []
        # pre-net (module.py:1509-1511; speaker variant multi_speaker_modules.py:27-32); no dropout outside training
        a_in = dict(lda=OU, a_off=OU - d.dec_in, a_tstride=B * OU, t_ptr=t_dev)
        if getattr(self, "fused_decode_tail", True) and max(d.dec_in, P0, P1) <= 256:
            # the whole pre-net in one cluster-per-utterance launch (satk_mlp_chain)
            if d.use_speaker:
                layers = [dict(W=p["dec.prenet0.W0"], bias=p["dec.prenet0.b0"], act="relu", residual=b_("dec.sp", (B, P0))),
                          dict(W=p["dec.prenet0.W"], bias=p["dec.prenet0.b"], act="relu")]
            else:
                layers = [dict(W=p["dec.prenet0.W"], bias=p["dec.prenet0.b"], act="relu")]
            layers.append(dict(W=p["dec.prenet1.W"], bias=p["dec.prenet1.b"], act="relu"))
            steps.append(O.mlp_chain_desc(mel_hist, B, d.dec_in, layers, cell_in, x_ld=OU, x_off=OU - d.dec_in, x_tstride=B * OU,
                                          out_ld=W1C, out_pstride=B * W1C, t_ptr=t_dev))
        else:
            if d.use_speaker:
                h0 = b_("pred.h0", (B, P0))
                sp = b_("dec.sp", (B, P0))
                steps.append(O.rowgemm_desc(mel_hist, B, d.dec_in, [dict(W=p["dec.prenet0.W0"], bias=p["dec.prenet0.b0"], act="relu", C=h0,
                                                                          residual=sp)], **a_in))
                steps.append(O.rowgemm_desc(h0, B, P0, [dict(W=p["dec.prenet0.W"], bias=p["dec.prenet0.b"], act="relu", C=pp0)]))
            else:
                steps.append(O.rowgemm_desc(mel_hist, B, d.dec_in, [dict(W=p["dec.prenet0.W"], bias=p["dec.prenet0.b"], act="relu", C=pp0)],
                                            **a_in))
            steps.append(O.rowgemm_desc(pp0, B, P0, [dict(W=p["dec.prenet1.W"], bias=p["dec.prenet1.b"], act="relu", C=cell_in, ldc=W1C,
                                                          c_pstride=B * W1C)], t_ptr=t_dev))
        # LSTM-1 on [prenet | attention | h] (AttentionWrapper concat, A.7) with the cell update in the epilogue
        steps.append(O.rowgemm_desc(cell_in, B, W1C, [dict(W=p["dec.lstm1.W"], bias=p["dec.lstm1.b"])], a_pstride=B * W1C, t_ptr=t_dev,
                                    lstm=dict(H=H1, c=st["c1"], h=st["h1"], out=x2c, ld_out=W2C, out_pstride=B * W2C,
                                              hdst=cell_in, ld_hdst=W1C, hdst_off=P1 + CTX, hdst_pstride=B * W1C, **zo)))
        QP = (d.att1 + d.att2 + 7) // 8
        fused_q = getattr(self, "fused_decode_tail", True) and QP <= 32 and d.att1 % QP == 0     # query layers inside satk_attn_step
        if not fused_q:
            qm = [dict(W=p["att1.query.W"], C=q, ldc=d.att1 + d.att2)]
            if d.dual:
                qm.append(dict(W=p["att2.query.W"], C=q, ldc=d.att1 + d.att2, c_off=d.att1))
            steps.append(O.rowgemm_desc(x2c, B, H1, qm, lda=W2C, a_pstride=B * W2C, t_ptr=t_dev))
        qkw = dict(Wq1=p["att1.query.W"], Wq2=p["att2.query.W"] if d.dual else None, q_x=x2c, q_x_ld=W2C, q_x_pstride=B * W2C,
                   q_in=H1) if fused_q else {}
        bufs = self._bufs
        agent = d.attention == "forward" and d.transition_agent
        steps.append(O.attn_step_desc(
            B=B, Tt=Tt, A1=d.att1, A2=d.att2, M1=d.mem1, M2=d.mem2, att_kernel=d.att_kernel if loc else 0,
            att_filters=d.att_filters if loc else 0, mode=_MODE[d.attention], cumulative=int(d.cumulative), use_agent=int(agent),
            t_ptr=t_dev, lengths=lengths, q=q, ldq=d.att1 + d.att2, keys1=bufs["dec.keys1"], values1=bufs["dec.values1"],
            v1=p["att1.v"], b1=p["att1.b"] if loc else None, loc_conv_w=p["att1.loc_conv.W"] if loc else None,
            loc_conv_b=p["att1.loc_conv.b"] if loc else None, loc_layer_w=p["att1.loc_layer.W"] if loc else None,
            keys2=bufs["dec.keys2"] if d.dual else None, values2=bufs["dec.values2"] if d.dual else None,
            v2=p["att2.v"] if d.dual else None, agent_w=p["att1.agent.W"] if agent else None,
            agent_b=p["att1.agent.b"] if agent else None, aprev=aprev, alpha=alpha, u=u,
            ctx_dst0=cell_in.data_ptr() + 4 * P1, ld0=W1C, pstride0=B * W1C,
            ctx_dst1=x2c.data_ptr() + 4 * H1, ld1=W2C, pstride1=B * W2C, align1=al1, align2=al2,
            # forced-alignment mode: the step replays given alignments (teacher_forcing_attention.py:13-78)
            forced1=self.buf("pred.forced1", (Tmax, B, Tt)) if forced else None,
            forced2=self.buf("pred.forced2", (Tmax, B, Tt)) if (forced and d.dual) else None, **qkw))
        # LSTM-2 / LSTM-3 (DecoderRNNV2 on ConcatOutputAndAttentionWrapper, module.py:1024,1525-1534)
        steps.append(O.rowgemm_desc(x2c, B, W2C, [dict(W=p["dec.lstm2.W"], bias=p["dec.lstm2.b"])], a_pstride=B * W2C, t_ptr=t_dev,
                                    lstm=dict(H=HD, c=st["c2"], h=st["h2"], out=x3c, ld_out=W3C, out_pstride=B * W3C,
                                              hdst=x2c, ld_hdst=W2C, hdst_off=X2W, hdst_pstride=B * W2C, **zo)))
        steps.append(O.rowgemm_desc(x3c, B, W3C, [dict(W=p["dec.lstm3.W"], bias=p["dec.lstm3.b"])], a_pstride=B * W3C, t_ptr=t_dev,
                                    lstm=dict(H=HD, c=st["c3"], h=st["h3"], out=o3, ld_out=HD, out_pstride=0,
                                              hdst=x3c, ld_hdst=W3C, hdst_off=HD, hdst_pstride=B * W3C, **zo)))
        x = o3
        probs = []
        tick_fused = False
        if getattr(self, "fused_decode_tail", True) and HD == 256 and OU + 1 <= 256:
            # fused tail: self-attention hops over the KV cache + mel / stop projections, one cluster per utterance (satk_sa_tail)
            hops = []
            for h in range(d.dec_sa_hops if d.dual else 0):
                n = f"dec.sa{h}"
                pr = b_(f"pred.P{h}", (B, d.dec_sa_heads, Tmax, Tmax), zero=True)
                probs.append(pr)
                hops.append(dict(Wk=p[n + ".key.W"], bk=p[n + ".key.b"], Wv=p[n + ".value.W"], bv=p[n + ".value.b"],
                                 Wq=p[n + ".query.W"], bq=p[n + ".query.b"], Wo=p[n + ".output.W"], bo=p[n + ".output.b"],
                                 Wt=p[n + ".transform.W"], bt=p[n + ".transform.b"], Kc=b_(f"pred.K{h}", (Tmax, B, HD)),
                                 Vc=b_(f"pred.V{h}", (Tmax, B, HD)), probs=pr))
            steps.append(O.sa_tail_desc(B=B, D=HD, heads=d.dec_sa_heads if d.dual else 1, Tmax=Tmax, t_ptr=t_dev, x=o3, ldx=HD, hops=hops,
                                        W_out=p["dec.out_proj.W"], b_out=p["dec.out_proj.b"], W_stop=p["dec.stop_proj.W"],
                                        b_stop=p["dec.stop_proj.b"], mel_dst=mel_hist, mel_tstride=B * OU, stop_dst=stop_hist,
                                        tick=dict(counter=b_("pred.tick_counter", (1,), torch.int32, zero=True), done_step=done,
                                                  min_iters=min_iters, use_stop=use_stop_token)))
            tick_fused = True
        else:
            if d.dual:
                # TransformerWrapper (rnn_wrappers.py:111-124) with cached keys / values: row t of the causal attention
                for h in range(d.dec_sa_hops):
                    n = f"dec.sa{h}"
                    D = HD
                    Kc, Vc = b_(f"pred.K{h}", (Tmax, B, D)), b_(f"pred.V{h}", (Tmax, B, D))
                    Qb, Ob, ao, y = b_(f"pred.Q{h}", (B, D)), b_(f"pred.O{h}", (B, D)), b_(f"pred.ao{h}", (B, D)), b_(f"pred.y{h}", (B, D))
                    pr = b_(f"pred.P{h}", (B, d.dec_sa_heads, Tmax, Tmax), zero=True)
                    probs.append(pr)
                    steps.append(O.rowgemm_desc(x, B, D, [
                        dict(W=p[n + ".key.W"], bias=p[n + ".key.b"], C=Kc, c_tstride=B * D),
                        dict(W=p[n + ".value.W"], bias=p[n + ".value.b"], C=Vc, c_tstride=B * D),
                        dict(W=p[n + ".query.W"], bias=p[n + ".query.b"], C=Qb)], t_ptr=t_dev))
                    steps.append(O.sa_step_desc(B=B, D=D, heads=d.dec_sa_heads, Tmax=Tmax, t_ptr=t_dev, q=Qb, ldq=D, Kc=Kc, Vc=Vc,
                                                out=Ob, ldo=D, probs=pr))
                    steps.append(O.rowgemm_desc(Ob, B, D, [dict(W=p[n + ".output.W"], bias=p[n + ".output.b"], C=ao)]))
                    steps.append(O.rowgemm_desc(ao, B, D, [dict(W=p[n + ".transform.W"], bias=p[n + ".transform.b"], act="tanh", C=y,
                                                                residual=x)]))
                    x = y
            # OutputAndStopTokenTransparentWrapper (rnn_wrappers.py:188-214): mel frames of step t -> row t+1 (row 0 = go frame)
            steps.append(O.rowgemm_desc(x, B, HD, [
                dict(W=p["dec.out_proj.W"], bias=p["dec.out_proj.b"], C=mel_hist, c_off=B * OU, c_tstride=B * OU),
                dict(W=p["dec.stop_proj.W"], bias=p["dec.stop_proj.b"], C=stop_hist, ldc=1, c_tstride=B)], t_ptr=t_dev))

        def run_step():
            for s_ in steps:
                if isinstance(s_, O.RowGemmDesc):
                    O.rowgemm(s_)
                elif isinstance(s_, O.AttnStepDesc):
                    O.attn_step(s_)
                elif isinstance(s_, O.SaTailDesc):
                    O.sa_tail(s_)
                elif isinstance(s_, O.MlpChainDesc):
                    O.mlp_chain(s_)
                else:
                    O.sa_step(s_)
            if not tick_fused:
                O.decode_tick(t_dev, stop_hist if use_stop_token else None, B, min_iters, done)

        state = dict(t_dev=t_dev, done=done, mel_hist=mel_hist, stop_hist=stop_hist, cell_in=cell_in, x2c=x2c, x3c=x3c, st=st,
                     aprev=aprev, alpha=alpha, u=u, al1=al1, al2=al2, lengths=lengths, probs=probs, steps=steps)
        return run_step, state

    def predict(self, features, max_iters: Optional[int] = None, use_stop_token: bool = True, min_iters: int = 10,
                use_graph: bool = True, check_every: int = 64, forced_alignments=None):
        """Free-running inference (model_fn in PREDICT mode, models/models.py:351-408 with is_training=False; decoder
        branch module.py:762-778).  Returns mel [B, T*r, n_mels], stop logits [B, T], alignments (B, Tt, T) and the decoder
        self-attention alignments; T = number of executed steps (stop token or max_iters).
        ``forced_alignments`` = (align1 [T,B,Tt], align2 | None), time-major: forced-alignment mode (teacher_forcing_attention.py)."""
        d = self.d
        source, source_length = features.source, features.source_length
        B, Tt = source.shape
        Tmax = int(max_iters or d.max_iters)
        self._training = False
        spk = None
        if d.use_speaker:
            spk = self.buf("spk_embed", (B, d.speaker_dim))
            O.embedding_fwd(features.speaker_id, self.ps.p["speaker_embedding"], spk, offset=d.speaker_offset)
        mem1, mem2, enc_al = self.encoder(source, source_length, False, None)
        p = self.ps.p
        R = Tt * B
        # attention memories (BahdanauAttention.__init__, A.8) — same buffers as the teacher-forced decoder
        values1 = self.buf("dec.values1", (R, d.mem1))
        O.mask_rows(mem1, source_length, B, Tt, d.mem1, True, values1)
        self.lin(values1, "att1.memory.W", self.buf("dec.keys1", (R, d.att1)))
        if d.dual:
            values2 = self.buf("dec.values2", (R, d.mem2))
            O.mask_rows(mem2, source_length, B, Tt, d.mem2, True, values2)
            self.lin(values2, "att2.memory.W", self.buf("dec.keys2", (R, d.att2)))
        if d.use_speaker:
            sp_pre = self.lin(spk, "dec.prenet0.Ws", self.buf("dec.sp_pre", (B, d.dec_prenet[0])), bias=p["dec.prenet0.bs"])
            O.softsign_fwd(sp_pre, self.buf("dec.sp", (B, d.dec_prenet[0])))
        # the step descriptors (and the captured graph) hold raw pointers of the shared memory / key buffers: `buf` re-allocates a
        # buffer when another (B, Tt) passed through forward / train_step in between, so the pointers are part of the key
        shared = ("dec.values1", "dec.keys1", "dec.values2", "dec.keys2", "dec.sp", "spk_embed", "enc.mem1", "enc.mem2")
        forced = forced_alignments is not None
        key = (B, Tt, Tmax, bool(use_stop_token), int(min_iters), bool(getattr(self, "fused_decode_tail", True)), forced,
               tuple(self._bufs[k].data_ptr() for k in shared if k in self._bufs), mem1.data_ptr(), 0 if mem2 is None else mem2.data_ptr())
        cache = getattr(self, "_decode_cache", None)
        if cache is None or cache["key"] != key:
            run_step, stt = self._build_decode_step(B, Tt, Tmax, use_stop_token, min_iters, forced)
            cache = self._decode_cache = dict(key=key, run_step=run_step, state=stt, graph=None)
        run_step, stt = cache["run_step"], cache["state"]
        # initial state: zero LSTM states / attention, alpha_0 = one-hot(0), u_0 = 0.5 (forward_attention.py:128-136)
        for k in ("cell_in", "x2c", "x3c", "aprev", "alpha"):
            stt[k].zero_()
        for v in stt["st"].values():
            v.zero_()
        stt["mel_hist"][0].zero_()
        stt["alpha"][:, 0] = 1.0
        stt["u"].fill_(0.5)
        stt["t_dev"].zero_()
        stt["done"].fill_(-1)
        stt["lengths"].copy_(source_length)
        if forced:
            # forced-alignment mode (models/models.py:411-427): alignments [T,B,Tt] of a teacher-forced pass replace the attention
            # mechanisms; the decoder still feeds back its own output
            f1, f2 = forced_alignments
            if f1.shape[0] < Tmax:
                raise ValueError(f"forced alignments cover {f1.shape[0]} decoder steps, {Tmax} requested")
            self._bufs["pred.forced1"].copy_(f1[:Tmax])
            if d.dual:
                self._bufs["pred.forced2"].copy_(f2[:Tmax])
        for pr in stt["probs"]:
            pr.zero_()
        n_run = 0
        if use_graph and Tmax > 1:
            run_step()                                   # step 0 eagerly (also sets kernel attributes before the capture)
            n_run = 1
            if cache["graph"] is None:
                torch.cuda.synchronize()
                saved_state = {k: v.clone() for k, v in self._bufs.items() if k.startswith("pred.")}
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    run_step()
                cache["graph"] = g
                for k, v in saved_state.items():         # the capture does not execute, but keep the state provably intact
                    self._bufs[k].copy_(v)
            g = cache["graph"]
        T_done = None
        while n_run < Tmax:
            chunk = min(check_every, Tmax - n_run)
            for _ in range(chunk):
                if use_graph and Tmax > 1:
                    g.replay()
                else:
                    run_step()
            n_run += chunk
            if use_stop_token:
                ds = int(stt["done"].item())
                if ds >= 0:
                    T_done = ds + 1
                    break
        T = T_done if T_done is not None else Tmax
        if use_stop_token and T_done is None:
            ds = int(stt["done"].item())
            if ds >= 0:
                T = ds + 1
        mel = stt["mel_hist"][1:T + 1].view(T, B, d.r, d.n_mels).permute(1, 0, 2, 3).reshape(B, T * d.r, d.n_mels)
        out = dict(mel=mel, stop=stt["stop_hist"][:T].t(), mel_tm=stt["mel_hist"][1:T + 1], stop_tm=stt["stop_hist"][:T],
                   alignment=stt["al1"][:T].permute(1, 2, 0),
                   alignment2=stt["al2"][:T].permute(1, 2, 0) if d.dual else None,
                   dec_self_P=[pr[:, i, :T, :T] for pr in stt["probs"] for i in range(d.dec_sa_heads)],
                   enc_self_P=enc_al, steps=T, steps_executed=n_run)
        if d.postnet_v2:       # "mel_postnet" of the prediction dict (models/models.py:210)
            post_tm = self.postnet(stt["mel_hist"][1:T + 1].reshape(T * B, d.r * d.n_mels), T, B, False, None, key="pred.post")
            out["mel_postnet_tm"] = post_tm.view(T, B, d.r * d.n_mels)
            out["mel_postnet"] = post_tm.view(T, B, d.r, d.n_mels).permute(1, 0, 2, 3).reshape(B, T * d.r, d.n_mels)
        return out

    def validate(self, features, labels, forced_alignments=None):
        """EVAL-mode decode WITHOUT teacher forcing (tacotron2 ValidationHelper(teacher_forcing=False), models/models.py:384-395,
        467-482): the decoder feeds back its own output for exactly Tm / r steps (no stop token) and the losses are taken against
        the labels.  -> (losses [3] = mel, done, total; the free-running outputs of `predict`)."""
        d = self.d
        B, Tm = labels.mel.shape[0], labels.mel.shape[1]
        Td = Tm // d.r
        out = self.predict(features, max_iters=Td, use_stop_token=False, forced_alignments=forced_alignments)
        out3 = self.buf("val.loss3", (3,))
        mel_tm = out["mel_tm"].reshape(Td * B, d.r * d.n_mels)
        stop_tm = out["stop_tm"].reshape(Td * B, 1)
        O.losses(mel_tm, stop_tm, labels.mel, labels.done, labels.spec_loss_mask, labels.binary_loss_mask, B, Tm, d.n_mels, d.r,
                 out3, self.buf("val.dmel", mel_tm.shape), self.buf("val.dstop", stop_tm.shape), self.buf("val.scratch", (4,)))
        if d.postnet_v2:
            post_tm = self.postnet(mel_tm, Td, B, False, None, key="val.post")
            post3 = self.buf("val.post.loss3", (3,))
            O.losses(post_tm, stop_tm, labels.mel, labels.done, labels.spec_loss_mask, labels.binary_loss_mask, B, Tm, d.n_mels, d.r,
                     post3, self.buf("val.post.dmel", mel_tm.shape), self.buf("val.post.dstop", stop_tm.shape),
                     self.buf("val.post.scratch", (4,)))
            O.add(out3[2:], post3[:1], out3[2:])
            out = dict(out, mel_postnet_tm=post_tm.view(Td, B, d.r * d.n_mels), postnet_v2_mel_loss=post3[:1])
        return out3, out

    def backward(self, allreduce=None):
        """BPTT through decoder and encoder; gradients land in ``self.ps.grad`` (zeroed first).

        ``allreduce(flat, async_op=False)``: the data-parallel gradient sum (train.py:68 MirroredStrategy).  When the
        callable advertises ``supports_async`` the flat buffer goes out as TWO buckets: the decoder / attention suffix is reduced
        (asynchronously, on the collective's own stream) while the encoder's backward pass still runs, the encoder prefix
        right after it."""
        s = self.saved
        self.ps.grad.zero_()
        if s.get("dpost") is not None:         # PostNetV2 first: it adds onto d(mel output)
            self.postnet_backward(s["dpost"], s["dmel"])
        dmem1, dmem2 = self._timed("sec.decoder_bwd", self.decoder_backward, s["dmel"], s["dstop"], s["B"], s["Tt"], s["Td"],
                                   s["source_length"])
        if self.d.use_speaker:
            self._join(1)                      # the pre-net chain (its own stream beside the overlapped attention backward)
            g, p = self.ps.g, self.ps.p
            dsp_pre = self._dspk_pre
            spk = self._bufs["spk_embed"]
            O.linear_dw(spk, dsp_pre, g["dec.prenet0.Ws"], s["B"], self.d.speaker_dim, self.d.dec_prenet[0], split_k=1)
            O.colsum_acc(dsp_pre, s["B"], self.d.dec_prenet[0], g["dec.prenet0.bs"])
            dspk = self.buf("dspk_embed", (s["B"], self.d.speaker_dim))
            O.linear_dx(dsp_pre, p["dec.prenet0.Ws"], dspk, s["B"])
            O.embedding_bwd(s["features"].speaker_id, dspk, g["speaker_embedding"], offset=self.d.speaker_offset)
        handle = None
        overlapped = (allreduce is not None and getattr(allreduce, "supports_async", False) and self.d.l2_weight == 0
                      and os.environ.get("SATK_AR_OVERLAP", "1") != "0")
        if overlapped:
            self._join(1)
            self._wg_join()                                          # the decoder's weight gradients are in
            handle = allreduce(self.ps.grad[self._dec_off:], async_op=True)
        O.tf32_push("enc")
        self._timed("sec.encoder_bwd", self.encoder_backward, dmem1, dmem2, s["B"], s["Tt"], s["source_length"])
        O.tf32_pop()
        self._join(1)
        self._wg_join()                                              # every weight gradient is in before all-reduce / Adam
        if self.d.l2_weight > 0:
            # gradient of l2_regularization_loss (models/models.py:470-478): scale * w on the regularised tensors.  Every replica adds it
            # (TF: each replica's loss carries the term, gradients are averaged), the 1/world_size after the all-reduce restores it.
            O.l2_reg(self.ps.flat, self.ps.l2_mask(), self.d.l2_weight, g=self.ps.grad)
        if overlapped:
            allreduce(self.ps.grad[:self._dec_off])
            if handle is not None:
                handle.wait()                                        # orders the current stream behind the asynchronous bucket
        elif allreduce is not None:
            allreduce(self.ps.grad)

    def optimizer_step(self, world_size: int = 1):
        """clip_by_global_norm(1.0) + Adam + noam LR (models.py:485-498).  With world_size > 1 the caller has
        all-reduced (summed) ``ps.grad``; the 1/world_size average is folded into the kernel."""
        hp = self.hp
        lr = noam_lr(hp.initial_learning_rate, self.global_step, hp.learning_rate_step_factor) if hp.decay_learning_rate \
            else hp.initial_learning_rate
        O.grad_sumsq(self.ps.grad, self._sumsq)
        O.adam_clip(self.ps.flat, self.ps.grad, self.ps.adam_m, self.ps.adam_v, self._sumsq, 1.0 / world_size, 1.0, lr,
                    hp.adam_beta1, hp.adam_beta2, hp.adam_eps, self.global_step + 1)
        self.refresh_transposed()
        self.global_step += 1
        return lr

    def train_step(self, features, labels, masks=None, allreduce=None, world_size: int = 1):
        # (data-parallel steps: the NCCL collectives of the bucketed all-reduce are captured too — thread-local capture mode, and the
        # graphs must be released before the process group is destroyed: `release_graphs`; SATK_GRAPH_NCCL=0 keeps those steps eager)
        if (masks is None and (allreduce is None or os.environ.get("SATK_GRAPH_NCCL", "1") != "0") and getattr(self, "use_graph", False)
                and self.timers is None and self._side is not None
                and not torch.cuda.is_current_stream_capturing()):
            return self._train_step_graphed(features, labels, allreduce, world_size)
        out = self.forward(features, labels, True, masks)
        self.backward(allreduce)
        out["lr"] = self._timed("sec.optimizer", self.optimizer_step, world_size)
        return out

    # ------------------------------------------------------------------ CUDA-graphed train step
    def _train_step_graphed(self, features, labels, allreduce, world_size):
        """forward + backward (+ bucketed all-reduce) of one shape bucket (B, T_text, T_mel) replayed from a CUDA graph: the ~350
        launches of a step cost the host ~3 ms of enqueue time, and the bursts of short encoder / dense kernels are launch-bound.
        The pieces that depend on host scalars stay eager around the replay: the copy of the step's inputs into the graph's static
        buffers, the step's mask seed (one device word: the keep-mask launches themselves are part of the graph and read it) and
        clip + Adam (learning rate, step count).  The first two calls of a bucket run eagerly (they allocate the engine's buffers);
        the third one captures."""
        B, Tt = features.source.shape
        Tm = labels.mel.shape[1]
        key = (B, Tt, Tm, world_size, features.speaker_id is not None)
        st = self._graphs.get(key)
        if st is None:
            st = self._graphs[key] = dict(calls=0, graph=None)
        if st["graph"] is None:
            st["calls"] += 1
            if st["calls"] <= 2 or st.get("failed"):
                out = self.forward(features, labels, True, None)
                self.backward(allreduce)
                out["lr"] = self.optimizer_step(world_size)
                return out
            try:
                self._capture_train_graph(st, features, labels, allreduce)
            except Exception as ex:      # noqa: BLE001 - anything the capture refuses: stay eager, say so once
                st["failed"] = True
                torch.cuda.synchronize()
                import warnings
                warnings.warn(f"CUDA-graph capture of the train step failed ({type(ex).__name__}: {ex}); staying eager")
                return self.train_step(features, labels, None, allreduce, world_size)
        # eager prologue: this step's inputs and keep masks into the static buffers the graph reads
        for dst, src in zip(st["inputs"], self._graph_inputs(features, labels)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        self._seed_dev.fill_(self._mask_seed)      # the mask launches inside the graph draw from this step's seed
        self._mask_seed += 1
        st["graph"].replay()
        O.add_launches(st["launches"])
        out = dict(st["out"])
        out["lr"] = self.optimizer_step(world_size)
        return out

    def release_graphs(self) -> None:
        """Drop the captured train-step graphs (they hold NCCL resources: call this before torch.distributed.destroy_process_group)."""
        self._graphs.clear()
        torch.cuda.synchronize()

    @staticmethod
    def _graph_inputs(features, labels):
        return [features.source, features.source_length, features.speaker_id, labels.mel, labels.target_length, labels.done,
                labels.spec_loss_mask, labels.binary_loss_mask]

    def _capture_train_graph(self, st, features, labels, allreduce):
        dev = self.device
        ins = [None if x is None else x.to(dev).clone() for x in self._graph_inputs(features, labels)]
        sf = features._replace(source=ins[0], source_length=ins[1], speaker_id=ins[2])
        sl = labels._replace(mel=ins[3], target_length=ins[4], done=ins[5], spec_loss_mask=ins[6], binary_loss_mask=ins[7])
        B, Tt = ins[0].shape
        seed0 = self._mask_seed
        self.device_masks(B, Tt, ins[3].shape[1] // self.d.r)      # allocates the mask buffers outside the capture
        self._mask_seed = seed0                                    # (that draw does not count: the replay below is this step)
        if getattr(self, "_seed_dev", None) is None:
            self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._capture_seed_dev = self._seed_dev
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = O.launches()
        try:
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                out = self.forward(sf, sl, True, None)
                self.backward(allreduce)
        finally:
            self._capture_seed_dev = None
        st.update(graph=g, inputs=ins, out=out, launches=O.launches() - l0, static=(sf, sl))
