"""ORACLE — test infrastructure, not product code.

CPU restatement (torch, fp32 or fp64) of the reference's teacher-forced training hot path:
encoder (CBHG + BiLSTM + self-attention) and attention decoder of Self-Attention Tacotron.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import this package; the product (``self-attention-tacotron_b200``) never does.

PARITY UNPINNED: the reference ships no golden vectors, checkpoints or fixtures for this path
 and cannot be executed here (TensorFlow 1.x and the un-vendored ``tacotron2``
dependency are absent).  Each function below cites the reference file:line it
restates; arithmetic that lives out of tree (TF1 layers, ``tacotron2@6af04c7f``) follows the
published semantics collected in SURVEY.md Appendix A and is marked "A.n".  RECALLED items are
exposed as switches on ``Switches`` so a wrong recollection is a flag flip.

Conventions: weights are TF layout ([in, out], conv [k, C_in, C_out]); ``P`` maps the names of
``params.param_specs`` to tensors; ``masks`` maps the names of ``data.mask_shapes`` to 0/1 keep
masks (TRAIN mode).  ``training=False`` is the deterministic eval arithmetic (moving BN stats,
zoneout interpolation, no dropout).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class Switches:
    """RECALLED semantics that could not be checked against source."""
    zoneout_on_output: bool = False     # A.6: cell output is the un-zoned h_new
    bn_eps: float = 1e-3                # A.3: tf.layers.batch_normalization default epsilon
    bn_momentum: float = 0.99           # A.3
    bn_bessel_moving_var: bool = True   # A.3: fused BN feeds the unbiased variance to the moving average
    forget_bias: float = 1.0            # A.5
    decoder_v2_num_layers: int = 2      # A.7: DecoderRNNV2 = attention cell + 2 zoneout LSTMs


SW = Switches()


# ----------------------------------------------------------------------------------------------
# primitives (Appendix A)
# ----------------------------------------------------------------------------------------------
def dense(x, W, b=None):
    y = x @ W
    return y if b is None else y + b


def dropout_mask(x, mask, keep):
    """tf.layers.dropout with an explicit keep mask: kept units scaled by 1/keep (A.2)."""
    if mask is None:
        return x
    return x * (mask.to(x.dtype) / keep)


def prenet(x, W, b, mask, keep):
    """tacotron2 PreNet = Dense(relu) -> dropout (A.2); call sites module.py:394,426,1509."""
    return dropout_mask(torch.relu(dense(x, W, b)), mask, keep)


def conv1d_same(x, W):
    """tf.layers.Conv1D(padding="SAME", use_bias=False), stride 1 (A.3).
    x [B,T,Cin], W [k,Cin,Cout]; cross-correlation, pad_left=(k-1)//2, extra pad on the right."""
    k = W.shape[0]
    pl = (k - 1) // 2
    pr = k - 1 - pl
    xp = F.pad(x.transpose(1, 2), (pl, pr))
    return F.conv1d(xp, W.permute(2, 1, 0)).transpose(1, 2)


def batch_norm(x, gamma, beta, mov_mean, mov_var, training, stats_out=None, key=None):
    """tf.layers.batch_normalization over all B*T positions, padding included (A.3)."""
    if training:
        mean = x.mean(dim=(0, 1))
        var = x.var(dim=(0, 1), unbiased=False)
        if stats_out is not None:
            n = x.shape[0] * x.shape[1]
            stats_out[key] = (mean.detach(), (var * (n / max(n - 1, 1)) if SW.bn_bessel_moving_var else var).detach())
    else:
        mean, var = mov_mean, mov_var
    inv = torch.rsqrt(var + SW.bn_eps) * gamma
    return x * inv + (beta - mean * inv)


def conv1d_bn(x, P, name, act, training, stats_out):
    """tacotron2 Conv1d: conv(SAME, no bias) -> BN -> activation (A.3); module.py:46-68."""
    y = conv1d_same(x, P[name + ".W"])
    y = batch_norm(y, P[name + ".gamma"], P[name + ".beta"], P.get(name + ".mean"), P.get(name + ".var"),
                   training, stats_out, name)
    return act(y) if act is not None else y


def maxpool2_same(x):
    """MaxPooling1D(pool=2, stride=1, SAME): y[t]=max(x[t],x[t+1]), y[T-1]=x[T-1] (A.3); module.py:54,80."""
    nxt = torch.cat([x[:, 1:], x[:, -1:]], dim=1)
    return torch.maximum(x, nxt)


def highway(x, WH, bH, WT, bT):
    """tacotron2 HighwayNet (A.4); module.py:72,91."""
    h = torch.relu(dense(x, WH, bH))
    t = torch.sigmoid(dense(x, WT, bT))
    return h * t + x * (1.0 - t)


def lstm_cell(x, c, h, W, b):
    """TF LSTMCell: z=[x,h]W+b split i,j,f,o; forget_bias 1 (A.5)."""
    z = torch.cat([x, h], dim=-1) @ W + b
    i, j, f, o = z.chunk(4, dim=-1)
    c_new = torch.sigmoid(f + SW.forget_bias) * c + torch.sigmoid(i) * torch.tanh(j)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return c_new, h_new


def zoneout(new, prev, mask, z, training):
    """tacotron2 ZoneoutLSTMCell state update (A.6).
    training: (1-z)*dropout(new-prev, keep=1-z)+prev == keep_mask*(new-prev)+prev; eval: (1-z)*new+z*prev."""
    if training:
        return prev + mask.to(new.dtype) * (new - prev)
    return (1.0 - z) * new + z * prev


def zoneout_lstm_step(x, c, h, W, b, mc, mh, zc, zh, training):
    c_new, h_new = lstm_cell(x, c, h, W, b)
    c_st = zoneout(c_new, c, mc, zc, training)
    h_st = zoneout(h_new, h, mh, zh, training)
    out = h_st if SW.zoneout_on_output else h_new
    return out, c_st, h_st


def zoneout_lstm_sequence(x, lengths, W, b, mc, mh, zc, zh, training, reverse=False):
    """One direction of tf.nn.bidirectional_dynamic_rnn(sequence_length=lengths) (A.6; module.py:93-108):
    past a sequence's length the output is zero and the state is carried through unchanged; the backward
    direction runs over the length-reversed sequence (tf.reverse_sequence) and is reversed back."""
    B, T, _ = x.shape
    H = W.shape[1] // 4
    c = x.new_zeros(B, H)
    h = x.new_zeros(B, H)
    outs = [None] * T
    tpos = torch.arange(T, device=x.device)
    for s in range(T):
        if reverse:
            # step s of the reversed sequence reads original position len-1-s
            idx = (lengths - 1 - s).clamp(min=0)
            xt = x[torch.arange(B), idx]
        else:
            xt = x[:, s]
        valid = (s < lengths).to(x.dtype)[:, None]
        m_c = None if mc is None else mc[s]
        m_h = None if mh is None else mh[s]
        out, c_n, h_n = zoneout_lstm_step(xt, c, h, W, b, m_c, m_h, zc, zh, training)
        c = valid * c_n + (1 - valid) * c
        h = valid * h_n + (1 - valid) * h
        outs[s] = out * valid
    y = torch.stack(outs, dim=1)               # [B,T,H] in processing order
    if reverse:
        # position p holds processing step len-1-p (zero past length)
        idx = (lengths[:, None] - 1 - tpos[None, :])
        ok = (idx >= 0)
        y = torch.gather(y, 1, idx.clamp(min=0)[:, :, None].expand(-1, -1, H)) * ok[:, :, None].to(x.dtype)
    return y


def multihead_attention(x, P, name, heads, causal, mask, keep):
    """MultiHeadAttention.call + ScaledDotProductAttentionMechanism.__call__
    (self_attention.py:108-128, 45-65).  No padding mask (use_padding_mask defaults to False and is never
    set: self_attention.py:133-135, module.py:353-356); causal mask when ``use_subsequent_mask``
    (self_attention.py:80-86).  Returns (output, [alignment per head])."""
    B, T, D = x.shape
    dh = P[name + ".key.W"].shape[1] // heads

    def split(t):
        return t.view(B, T, heads, dh).permute(0, 2, 1, 3)

    k = split(dense(x, P[name + ".key.W"], P[name + ".key.b"]))
    v = split(dense(x, P[name + ".value.W"], P[name + ".value.b"]))
    q = split(dense(x, P[name + ".query.W"], P[name + ".query.b"]))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if causal:
        tri = torch.ones(T, T, dtype=torch.bool, device=x.device).tril()
        s = torch.where(tri, s, torch.full_like(s, -float("inf")))
    a = torch.softmax(s, dim=-1)
    ad = dropout_mask(a, mask, keep)
    o = (ad @ v).permute(0, 2, 1, 3).reshape(B, T, heads * dh)
    o = dense(o, P[name + ".output.W"], P[name + ".output.b"])
    return o, [a[:, i] for i in range(heads)]


def self_attention_transformer(x, P, name, heads, causal, mask, keep):
    """SelfAttentionTransformer.call (module.py:363-371): MHA -> Dense(tanh) -> residual."""
    o, al = multihead_attention(x, P, name, heads, causal, mask, keep)
    t = torch.tanh(dense(o, P[name + ".transform.W"], P[name + ".transform.b"]))
    return x + t, al


# ----------------------------------------------------------------------------------------------
# encoder (module.py:425-438 / 77-110)
# ----------------------------------------------------------------------------------------------
def encoder_forward(P, d, source, source_length, training, masks=None, stats_out=None):
    masks = masks or {}
    x = P["embedding"][source]                                               # models.py:351 (A.1)
    for i in range(len(d.enc_prenet)):                                       # module.py:426
        mk = masks.get(f"enc.prenet{i}") if training else None                # masks are time-major [Tt,B,u]
        x = prenet(x, P[f"enc.prenet{i}.W"], P[f"enc.prenet{i}.b"],
                   None if mk is None else mk.transpose(0, 1), 1.0 - d.enc_prenet_drop)
    inp = x
    bank = torch.cat([conv1d_bn(inp, P, f"cbhg.bank{k}", torch.relu, training, stats_out)
                      for k in range(1, d.bank_k + 1)], dim=-1)              # module.py:78
    mp = maxpool2_same(bank)                                                 # module.py:80
    p1 = conv1d_bn(mp, P, "cbhg.proj1", torch.relu, training, stats_out)     # module.py:82
    p2 = conv1d_bn(p1, P, "cbhg.proj2", None, training, stats_out)           # module.py:83
    hwy = p2 + inp                                                           # module.py:86
    for i in range(d.n_highway):                                             # module.py:91
        hwy = highway(hwy, P[f"cbhg.highway{i}.WH"], P[f"cbhg.highway{i}.bH"],
                      P[f"cbhg.highway{i}.WT"], P[f"cbhg.highway{i}.bT"])
    outs = []
    for dr in ("fw", "bw"):                                                  # module.py:93-110
        outs.append(zoneout_lstm_sequence(
            hwy, source_length, P[f"cbhg.lstm_{dr}.W"], P[f"cbhg.lstm_{dr}.b"],
            masks.get(f"cbhg.lstm_{dr}.c"), masks.get(f"cbhg.lstm_{dr}.h"), d.zc, d.zh, training,
            reverse=(dr == "bw")))
    lstm_out = torch.cat(outs, dim=-1)
    if not d.dual:
        return lstm_out, None, []
    sa = dense(lstm_out, P["enc.sa_proj.W"], P["enc.sa_proj.b"])             # module.py:429
    aligns: List[torch.Tensor] = []
    for h in range(d.enc_sa_hops):                                           # module.py:435
        sa, al = self_attention_transformer(sa, P, f"enc.sa{h}", d.enc_sa_heads, False,
                                            masks.get(f"enc.sa{h}") if training else None, 1.0 - d.enc_sa_drop)
        aligns += al
    return lstm_out, sa, aligns


# ----------------------------------------------------------------------------------------------
# attention mechanisms (forward_attention.py, TF BahdanauAttention A.8)
# ----------------------------------------------------------------------------------------------
def seq_mask(lengths, T, dtype):
    return (torch.arange(T, device=lengths.device)[None, :] < lengths[:, None]).to(dtype)


def attention_memory(memory, lengths, W_mem):
    """BahdanauAttention.__init__ (A.8): values = memory zeroed past length; keys = values . W_mem."""
    values = memory * seq_mask(lengths, memory.shape[1], memory.dtype)[:, :, None]
    return values @ W_mem, values


def masked_softmax(e, lengths):
    """_maybe_mask_score(-inf past length) + softmax (A.8)."""
    m = seq_mask(lengths, e.shape[1], torch.bool)
    return torch.softmax(torch.where(m, e, torch.full_like(e, -float("inf"))), dim=-1)


def location_energy(P, d, q, prev_align, keys):
    """forward_attention.py:92-103 + _location_sensitive_score :13-26."""
    f = conv1d_same(prev_align[:, :, None], P["att1.loc_conv.W"]) + P["att1.loc_conv.b"]   # :98-100, SAME k: left (k-1)//2
    pf = f @ P["att1.loc_layer.W"]                                                         # :101
    return (P["att1.v"] * torch.tanh(keys + q[:, None, :] + pf + P["att1.b"])).sum(-1)     # :26


def attention1_step(P, d, query, state, keys, values, lengths):
    """One call of the attention-1 mechanism; returns (alignments fed to context, next_state)."""
    q = query @ P["att1.query.W"]
    if d.attention == "additive":                                            # TF _bahdanau_score (A.8)
        e = (P["att1.v"] * torch.tanh(keys + q[:, None, :])).sum(-1)
        a = masked_softmax(e, lengths)
        return a, (a,)
    prev_a = state[0]
    e = location_energy(P, d, q, prev_a, keys)
    a = masked_softmax(e, lengths)                                           # forward_attention.py:105
    if d.attention == "location_sensitive":                                  # tacotron2 LocationSensitiveAttention (A.8)
        return a, ((a + prev_a) if d.cumulative else a,)
    prev_a, prev_alpha, prev_u = state
    shifted = F.pad(prev_alpha[:, :-1], (1, 0))                              # forward_attention.py:108
    alpha = ((1 - prev_u) * prev_alpha + prev_u * shifted + 1e-7) * a        # :109
    alpha = alpha / alpha.sum(dim=1, keepdim=True)                           # :110
    if d.transition_agent:                                                   # :111-114
        ctx = (alpha[:, None, :] @ values).squeeze(1)
        u = torch.sigmoid(torch.cat([ctx, q], dim=-1) @ P["att1.agent.W"] + P["att1.agent.b"])
    else:
        u = prev_u                                                           # :116
    nxt = ((a + prev_a) if d.cumulative else a, alpha, u)                    # :118-121
    return alpha, nxt


def attention1_initial_state(d, B, Tt, dtype, device):
    """forward_attention.py:128-136 (alpha0 = one-hot(0), u0 = 0.5); Bahdanau: zeros."""
    a0 = torch.zeros(B, Tt, dtype=dtype, device=device)
    if d.attention == "forward":
        alpha0 = a0.clone()
        alpha0[:, 0] = 1.0
        return (a0, alpha0, torch.full((B, 1), 0.5, dtype=dtype, device=device))
    return (a0,)


def attention2_step(P, query, keys, lengths):
    q = query @ P["att2.query.W"]
    e = (P["att2.v"] * torch.tanh(keys + q[:, None, :])).sum(-1)
    return masked_softmax(e, lengths)


# ----------------------------------------------------------------------------------------------
# decoder (module.py:1493-1559, 726-760, 562-623; rnn_wrappers.py; helpers.py)
# ----------------------------------------------------------------------------------------------
def decoder_prenet(P, d, x, spk_embed, masks, training, step=None):
    """DecoderPreNetWrapper over (PreNet|MultiSpeakerPreNet, PreNet) (module.py:1506-1511,
    multi_speaker_modules.py:27-32).  ``x`` [B,Td,dec_in] (all steps) or [B,dec_in] with ``step``."""
    def m(i):
        if not training or masks is None or masks.get(f"dec.prenet{i}") is None:
            return None
        mk = masks[f"dec.prenet{i}"]                                          # time-major [Td,B,u]
        return mk.transpose(0, 1) if step is None else mk[step]
    keep = 1.0 - d.dec_prenet_drop
    if d.use_speaker:
        h0 = torch.relu(dense(x, P["dec.prenet0.W0"], P["dec.prenet0.b0"]))
        sp = F.softsign(dense(spk_embed, P["dec.prenet0.Ws"], P["dec.prenet0.bs"]))
        h0 = h0 + (sp if x.dim() == 2 else sp[:, None, :])
        h = dropout_mask(torch.relu(dense(h0, P["dec.prenet0.W"], P["dec.prenet0.b"])), m(0), keep)
    else:
        h = prenet(x, P["dec.prenet0.W"], P["dec.prenet0.b"], m(0), keep)
    return prenet(h, P["dec.prenet1.W"], P["dec.prenet1.b"], m(1), keep)


def teacher_inputs(d, target):
    """TransformerTrainingHelper / TrainingHelper (helpers.py:13-55, 224-225): input of step t is the
    last n_feed_frame frames of target group t-1; zeros (go frame) at t=0."""
    B, Tm, nm = target.shape
    Td = Tm // d.r
    grouped = target.reshape(B, Td, nm * d.r)[:, :, -nm * d.n_feed:]
    go = target.new_zeros(B, 1, nm * d.n_feed)
    return torch.cat([go, grouped[:, :-1]], dim=1)


def decoder_rnn_forward(P, d, memory1, memory2, source_length, dec_inputs, spk_embed, training, masks=None):
    """dynamic_decode(BasicDecoder(DecoderRNNV2(attention cell))) under teacher forcing
    (module.py:744-747; A.7, A.11).  Returns decoder_outputs [B,Td,dec_out], alignment histories."""
    masks = masks or {}
    B, Td, _ = dec_inputs.shape
    Tt = memory1.shape[1]
    dt, dev = memory1.dtype, memory1.device
    keys1, values1 = attention_memory(memory1, source_length, P["att1.memory.W"])
    if d.dual:
        keys2, values2 = attention_memory(memory2, source_length, P["att2.memory.W"])
    pre = decoder_prenet(P, d, dec_inputs, spk_embed, masks, training)        # identical per step -> dense over t
    H1, HD = d.att_rnn, d.dec_out
    c1 = h1 = memory1.new_zeros(B, H1)
    c2 = h2 = memory1.new_zeros(B, HD)
    c3 = h3 = memory1.new_zeros(B, HD)
    attn = memory1.new_zeros(B, d.ctx)                                        # AttentionWrapper.zero_state: attention = 0
    st1 = attention1_initial_state(d, B, Tt, dt, dev)
    outs, al1, al2 = [], [], []

    def mk(name, t):
        v = masks.get(name)
        return None if v is None else v[t]

    for t in range(Td):
        cell_in = torch.cat([pre[:, t], attn], dim=-1)                        # AttentionWrapper: concat(inputs, attention)
        out1, c1, h1 = zoneout_lstm_step(cell_in, c1, h1, P["dec.lstm1.W"], P["dec.lstm1.b"],
                                         mk("dec.lstm1.c", t), mk("dec.lstm1.h", t), d.zc, d.zh, training)
        a1, st1 = attention1_step(P, d, out1, st1, keys1, values1, source_length)
        ctx1 = (a1[:, None, :] @ values1).squeeze(1)                          # == _calculate_context, forward_attention.py:29-41
        al1.append(a1)
        if d.dual:
            a2 = attention2_step(P, out1, keys2, source_length)
            ctx2 = (a2[:, None, :] @ values2).squeeze(1)
            al2.append(a2)
            attn = torch.cat([ctx1, ctx2], dim=-1)
        else:
            attn = ctx1
        x2 = torch.cat([out1, attn], dim=-1)                                  # ConcatOutputAndAttentionWrapper
        out2, c2, h2 = zoneout_lstm_step(x2, c2, h2, P["dec.lstm2.W"], P["dec.lstm2.b"],
                                         mk("dec.lstm2.c", t), mk("dec.lstm2.h", t), d.zc, d.zh, training)
        out3, c3, h3 = zoneout_lstm_step(out2, c3, h3, P["dec.lstm3.W"], P["dec.lstm3.b"],
                                         mk("dec.lstm3.c", t), mk("dec.lstm3.h", t), d.zc, d.zh, training)
        outs.append(out3)
    dec_out = torch.stack(outs, dim=1)
    al1 = torch.stack(al1, dim=2)                                             # (B, Tt, Td)  models.py:401/407
    al2 = torch.stack(al2, dim=2) if d.dual else None
    return dec_out, al1, al2


def decoder_forward(P, d, memory1, memory2, source_length, target, spk_embed, training, masks=None,
                    inference_branch=False):
    """DualSourceTransformerDecoder.call / ExtendedDecoder.call (teacher forced)."""
    masks = masks or {}
    dec_inputs = teacher_inputs(d, target)
    dec_out, al1, al2 = decoder_rnn_forward(P, d, memory1, memory2, source_length, dec_inputs, spk_embed,
                                            training, masks)
    sa_aligns: List[torch.Tensor] = []
    x = dec_out
    if d.dual:
        if not inference_branch:                                              # module.py:749-757
            for h in range(d.dec_sa_hops):
                x, al = self_attention_transformer(x, P, f"dec.sa{h}", d.dec_sa_heads, True,
                                                   masks.get(f"dec.sa{h}") if training else None, 1.0 - d.dec_sa_drop)
                sa_aligns += al
        else:                                                                 # module.py:762-778, rnn_wrappers.py:111-124
            rows = []
            for t in range(dec_out.shape[1]):
                hist = dec_out[:, :t + 1]
                for h in range(d.dec_sa_hops):
                    hist, al = self_attention_transformer(hist, P, f"dec.sa{h}", d.dec_sa_heads, True, None, 1.0)
                rows.append(hist[:, -1])
            x = torch.stack(rows, dim=1)
    mel = dense(x, P["dec.out_proj.W"], P["dec.out_proj.b"])                  # module.py:758 / 599
    stop = dense(x, P["dec.stop_proj.W"], P["dec.stop_proj.b"])               # module.py:759
    B = mel.shape[0]
    return mel.reshape(B, -1, d.n_mels), stop, al1, al2, sa_aligns            # module.py:1558


def decoder_free_running(P, d, memory1, memory2, source_length, spk_embed, max_iters, min_iters=10, use_stop_token=True, forced=None):
    """PREDICT-mode decoder (module.py:762-778): dynamic_decode of
    OutputAndStopTokenTransparentWrapper(TransformerWrapper(RNNStateHistoryWrapper(decoder_cell))) driven by
    StopTokenBasedInferenceHelper.  Each step: the decoder cell (pre-net, LSTM-1 + attention(s), LSTM-2/3) on the
    previous step's last n_feed_frame predicted frames (zeros at t=0, helpers.py:224-225); its output is appended to the
    history (rnn_wrappers.py:69-75); the self-attention stack is re-run over the WHOLE history with the causal mask and the
    last row is kept (rnn_wrappers.py:111-124); mel and stop projections (rnn_wrappers.py:188-214).  The loop ends after the
    first step t with sigmoid(stop) > 0.5 for every utterance and t > min_iters (helpers.py:103-107 semantics), or after
    max_iters steps.  Returns mel [B, T*r, n_mels], stop [B, T], alignments (B, Tt, T)."""
    B = memory1.shape[0]
    Tt = memory1.shape[1]
    dt, dev = memory1.dtype, memory1.device
    keys1, values1 = attention_memory(memory1, source_length, P["att1.memory.W"])
    if d.dual:
        keys2, values2 = attention_memory(memory2, source_length, P["att2.memory.W"])
    H1, HD = d.att_rnn, d.dec_out
    c1 = h1 = memory1.new_zeros(B, H1)
    c2 = h2 = memory1.new_zeros(B, HD)
    c3 = h3 = memory1.new_zeros(B, HD)
    attn = memory1.new_zeros(B, d.ctx)
    st1 = attention1_initial_state(d, B, Tt, dt, dev)
    inp = memory1.new_zeros(B, d.n_mels * d.n_feed)
    hist, mels, stops, al1, al2 = [], [], [], [], []
    for t in range(max_iters):
        pre = decoder_prenet(P, d, inp, spk_embed, None, False, step=t)
        cell_in = torch.cat([pre, attn], dim=-1)
        out1, c1, h1 = zoneout_lstm_step(cell_in, c1, h1, P["dec.lstm1.W"], P["dec.lstm1.b"], None, None, d.zc, d.zh, False)
        if forced is not None:
            # forced-alignment mode: TeacherForcingForwardAttention / TeacherForcingAdditiveAttention.__call__
            # (modules/teacher_forcing_attention.py:28-35,63-70): alignments = teacher_alignments[:, index]
            a1 = forced[0][:, :, t]
        else:
            a1, st1 = attention1_step(P, d, out1, st1, keys1, values1, source_length)
        ctx1 = (a1[:, None, :] @ values1).squeeze(1)
        al1.append(a1)
        if d.dual:
            a2 = forced[1][:, :, t] if forced is not None else attention2_step(P, out1, keys2, source_length)
            ctx2 = (a2[:, None, :] @ values2).squeeze(1)
            al2.append(a2)
            attn = torch.cat([ctx1, ctx2], dim=-1)
        else:
            attn = ctx1
        x2 = torch.cat([out1, attn], dim=-1)
        out2, c2, h2 = zoneout_lstm_step(x2, c2, h2, P["dec.lstm2.W"], P["dec.lstm2.b"], None, None, d.zc, d.zh, False)
        out3, c3, h3 = zoneout_lstm_step(out2, c3, h3, P["dec.lstm3.W"], P["dec.lstm3.b"], None, None, d.zc, d.zh, False)
        hist.append(out3)
        x = out3
        if d.dual:
            full = torch.stack(hist, dim=1)
            for h in range(d.dec_sa_hops):
                full, _ = self_attention_transformer(full, P, f"dec.sa{h}", d.dec_sa_heads, True, None, 1.0)
            x = full[:, -1]
        mel_t = dense(x, P["dec.out_proj.W"], P["dec.out_proj.b"])
        stop_t = dense(x, P["dec.stop_proj.W"], P["dec.stop_proj.b"]).squeeze(-1)
        mels.append(mel_t)
        stops.append(stop_t)
        inp = mel_t[:, -d.n_mels * d.n_feed:]
        if use_stop_token and t > min_iters and bool((torch.sigmoid(stop_t) > 0.5).all()):
            break
    mel = torch.stack(mels, dim=1).reshape(B, -1, d.n_mels)
    return (mel, torch.stack(stops, dim=1), torch.stack(al1, dim=2), torch.stack(al2, dim=2) if d.dual else None)


def model_forced_alignment(P, d, features, labels):
    """use_forced_alignment_mode outside training (models/models.py:384-427): a teacher-forced decode gives alignment1 / alignment2
    (B, Tt, Td); a second decode (is_validation=True, teacher_forcing=False: own outputs are fed back for Tm / r steps) runs with the
    teacher-forcing attention mechanisms that replay them (attention_factories.py:40-66)."""
    first = model_forward(P, d, features, labels, False)
    spk = P["speaker_embedding"][features.speaker_id - d.speaker_offset] if d.use_speaker else None
    mem1, mem2, _ = encoder_forward(P, d, features.source, features.source_length, False, None, None)
    Td = labels.mel.shape[1] // d.r
    mel, stop, al1, al2 = decoder_free_running(P, d, mem1, mem2, features.source_length, spk, Td, use_stop_token=False,
                                               forced=(first["alignment"], first.get("alignment2")))
    return dict(mel=mel, stop=stop, alignment=al1, alignment2=al2, mel_with_teacher=first["mel"])


def postnet_v2(P, d, mel, training, masks=None, stats_out=None):
    """PostNetV2 (models/models.py:92-100,440-462; class tacotron2.tacotron.tacotron_v2.PostNetV2 of the un-vendored dependency,
    RECALLED): num_postnet_v2_layers tacotron2 Conv1d layers (conv SAME without bias -> BN -> tanh, no activation on the last one,
    dropout postnet_v2_drop_rate behind each in training), a Dense projection back to num_mels and the residual connection.
    mel [B, T_mel, num_mels]; masks["postnet.conv<i>"] are frame-major [T_mel, B, channels] keep masks."""
    x = mel
    keep = 1.0 - d.postnet_drop
    for i in range(d.postnet_layers):
        act = torch.tanh if i < d.postnet_layers - 1 else None
        x = conv1d_bn(x, P, f"postnet.conv{i}", act, training, stats_out)
        if training:
            x = dropout_mask(x, masks[f"postnet.conv{i}"].transpose(0, 1), keep)
    return mel + dense(x, P["postnet.proj.W"], P["postnet.proj.b"])


def model_predict(P, d, features, max_iters=None, min_iters=10, use_stop_token=True):
    """model_fn in PREDICT mode (models/models.py:351-408 with is_training=False, no labels)."""
    spk = None
    if d.use_speaker:
        spk = P["speaker_embedding"][features.speaker_id - d.speaker_offset]
    mem1, mem2, enc_aligns = encoder_forward(P, d, features.source, features.source_length, False, None, None)
    mel, stop, al1, al2 = decoder_free_running(P, d, mem1, mem2, features.source_length, spk, max_iters or d.max_iters,
                                               min_iters, use_stop_token)
    out = dict(mel=mel, stop=stop, alignment=al1, alignment2=al2)
    if getattr(d, "postnet_v2", False):
        out["mel_postnet"] = postnet_v2(P, d, mel, False)              # models/models.py:210 "mel_postnet"
    return out


# ----------------------------------------------------------------------------------------------
# model_fn body: losses and optimiser (models/models.py:351-408, 467-498, 595-598)
# ----------------------------------------------------------------------------------------------
def spec_loss_l1(pred, target, mask):
    """tacotron2.losses.spec_loss 'l1' = compute_weighted_loss(SUM_BY_NONZERO_WEIGHTS) (A.9)."""
    w = mask[:, :, None]
    return (torch.abs(pred - target) * w).sum() / (torch.count_nonzero(mask).to(pred.dtype) * pred.shape[-1])


def binary_loss(logit, done, mask):
    """tf.losses.sigmoid_cross_entropy(done, squeeze(logit), weights=mask) (A.9)."""
    x = logit.squeeze(-1)
    ce = torch.clamp(x, min=0) - x * done + torch.log1p(torch.exp(-torch.abs(x)))
    return (ce * mask).sum() / torch.count_nonzero(mask).to(x.dtype)


def model_forward(P, d, features, labels, training, masks=None, stats_out=None):
    """model_fn up to the loss (models/models.py:351-482).  Returns a dict of outputs."""
    source, source_length = features.source, features.source_length
    spk = None
    if d.use_speaker:
        spk = P["speaker_embedding"][features.speaker_id - d.speaker_offset]  # A.1 index_offset
    mem1, mem2, enc_aligns = encoder_forward(P, d, source, source_length, training, masks, stats_out)
    mel, stop, al1, al2, dec_sa = decoder_forward(P, d, mem1, mem2, source_length, labels.mel.to(mem1.dtype), spk,
                                                  training, masks)
    mel_loss = spec_loss_l1(mel, labels.mel.to(mel.dtype), labels.spec_loss_mask.to(mel.dtype))
    done_loss = binary_loss(stop, labels.done.to(mel.dtype), labels.binary_loss_mask.to(mel.dtype))
    reg_loss = l2_regularization_loss(P, d) if getattr(d, "l2_weight", 0.0) > 0 else torch.zeros((), dtype=mel.dtype)
    post, post_loss = None, torch.zeros((), dtype=mel.dtype)
    if getattr(d, "postnet_v2", False):                                # models/models.py:92-100,116-118 / 440-462,479-482
        post = postnet_v2(P, d, mel, training, masks, stats_out)
        post_loss = spec_loss_l1(post, labels.mel.to(mel.dtype), labels.spec_loss_mask.to(mel.dtype))
    return dict(mel=mel, stop=stop, alignment=al1, alignment2=al2, regularization_loss=reg_loss,
                mel_postnet=post, postnet_v2_mel_loss=post_loss,
                enc_self_alignments=[a.transpose(1, 2) for a in enc_aligns],  # models.py:398 (B, T_mem, T_query)
                dec_self_alignments=[a.transpose(1, 2) for a in dec_sa],
                memory1=mem1, memory2=mem2,
                mel_loss=mel_loss, done_loss=done_loss, loss=mel_loss + done_loss + reg_loss + post_loss)   # models.py:482


# models/models.py:470-473
L2_BLACKLIST = ["embedding", "bias", "batch_normalization", "output_projection_wrapper/kernel", "lstm_cell",
                "output_and_stop_token_wrapper/dense/", "output_and_stop_token_wrapper/dense_1/", "stop_token_projection/kernel"]


def tf_variable_name(n: str, d) -> str:
    """The part of the reference's TF variable name that its black-list can match, for trainable tensor `n` of the parameter store.
    In-tree names: forward_attention.py:17,21,73,78,86 (attention_variable / attention_bias / location_features_convolution /
    location_features_layer / transition_factor_projection), module.py:718-723 (out_projection / stop_token_projection).
    RECALLED (tacotron2 / TF layers): dense and conv layers own "kernel" and "bias", tf.layers.batch_normalization
    owns "batch_normalization[_k]/gamma|beta", LSTMCell (inside ZoneoutLSTMCell) owns "lstm_cell/kernel|bias", Embedding owns
    "embedding", the ExtendedDecoder projects through OutputAndStopTokenWrapper ("output_and_stop_token_wrapper/dense[_1]/")."""
    leaf = n.rsplit(".", 1)[-1]
    if n in ("embedding", "speaker_embedding"):
        return n
    if ".lstm" in n:
        return "lstm_cell/" + ("kernel" if leaf == "W" else "bias")
    if leaf in ("gamma", "beta"):
        return "batch_normalization/" + leaf
    if n == "att1.v" or n == "att2.v":
        return "attention_variable" if n == "att1.v" and d.attention != "additive" else "attention_v"
    if n == "att1.b":
        return "attention_bias"
    if n.startswith("dec.out_proj") or n.startswith("dec.stop_proj"):
        kind = "kernel" if leaf == "W" else "bias"
        if d.dual:
            return ("out_projection/" if "out_proj" in n else "stop_token_projection/") + kind
        return ("output_and_stop_token_wrapper/dense/" if "out_proj" in n else "output_and_stop_token_wrapper/dense_1/") + kind
    return n + ("/kernel" if leaf.startswith("W") else "/bias")


def l2_regularization_loss(P, d):
    """modules/regularizers.py:11-18 with the black-list of models/models.py:470-478: scale * sum of tf.nn.l2_loss (= sum(w^2) / 2) over
    the trainable variables whose name contains no black-listed substring."""
    total = 0.0
    for n, w in P.items():
        if n.endswith(".mean") or n.endswith(".var"):      # BN moving statistics are not trainable
            continue
        name = tf_variable_name(n, d)
        if any(black in name for black in L2_BLACKLIST):
            continue
        total = total + 0.5 * (w.double() ** 2).sum()
    return (total * d.l2_weight).to(next(iter(P.values())).dtype)


def noam_lr(init_rate, global_step, step_factor):
    """learning_rate_decay (models/models.py:595-598)."""
    warm = 4000.0
    step = float(global_step * step_factor + 1)
    return init_rate * warm ** 0.5 * min(step * warm ** -1.5, step ** -0.5)


def clip_by_global_norm(grads: List[torch.Tensor], clip: float = 1.0):
    """tf.clip_by_global_norm (models/models.py:493)."""
    norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    scale = clip / torch.maximum(norm, torch.tensor(clip, dtype=norm.dtype))
    return [g * scale for g in grads], norm


def adam_update(p, g, m, v, lr, t, b1, b2, eps):
    """tf.train.AdamOptimizer dense update: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); eps outside the sqrt."""
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.sub_(lr_t * m / (torch.sqrt(v) + eps))


class OracleTrainer:
    """Full train step on the CPU restatement: fwd, autograd bwd, clip, Adam, BN moving stats."""

    def __init__(self, d, hp, params: Dict[str, torch.Tensor], dtype=torch.float32):
        self.d, self.hp = d, hp
        self.names = [n for n in params if not (n.endswith(".mean") or n.endswith(".var"))]
        self.P = {k: v.detach().clone().to(dtype) for k, v in params.items()}
        for n in self.names:
            self.P[n].requires_grad_(True)
        self.m = {n: torch.zeros_like(self.P[n]) for n in self.names}
        self.v = {n: torch.zeros_like(self.P[n]) for n in self.names}
        self.global_step = 0

    def loss_and_grads(self, features, labels, masks, training=True):
        for n in self.names:
            self.P[n].grad = None
        stats = {}
        out = model_forward(self.P, self.d, features, labels, training, masks, stats)
        out["loss"].backward()
        grads = {n: (self.P[n].grad if self.P[n].grad is not None else torch.zeros_like(self.P[n])) for n in self.names}
        return out, grads, stats

    def train_step(self, features, labels, masks):
        hp = self.hp
        out, grads, stats = self.loss_and_grads(features, labels, masks, True)
        clipped, norm = clip_by_global_norm([grads[n] for n in self.names], 1.0)
        lr = noam_lr(hp.initial_learning_rate, self.global_step, hp.learning_rate_step_factor) \
            if hp.decay_learning_rate else hp.initial_learning_rate
        t = self.global_step + 1
        with torch.no_grad():
            for n, g in zip(self.names, clipped):
                adam_update(self.P[n], g, self.m[n], self.v[n], lr, t, hp.adam_beta1, hp.adam_beta2, hp.adam_eps)
            for key, (mean, var) in stats.items():                           # UPDATE_OPS, models.py:497
                mom = SW.bn_momentum
                self.P[key + ".mean"].mul_(mom).add_(mean, alpha=1 - mom)
                self.P[key + ".var"].mul_(mom).add_(var, alpha=1 - mom)
        self.global_step += 1
        out["grad_norm"] = norm
        out["lr"] = lr
        return out
