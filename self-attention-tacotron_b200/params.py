"""Model dimensions and the parameter set of the hot path.

Every tensor here corresponds to a TF variable the reference creates on the path
``model_fn -> encoder -> decoder``.  Kernels keep the
TF layout ``[in, out]`` (``tf.layers.Dense``) / ``[k, C_in, C_out]`` (``tf.layers.Conv1D``)
so a TF checkpoint could later be mapped tensor-for-tensor.

All trainable tensors live in ONE flat fp32 buffer (each tensor padded to a multiple of 4
floats so every view is 16-byte aligned).  The flat buffer is what the gradient all-reduce
(one ncclAllReduce) and the fused clip+Adam kernel operate on.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class ModelDims:
    # model family: True = DualSourceSelfAttentionTacotronModel (models/models.py:275),
    #               False = ExtendedTacotronV1Model (models/models.py:20)
    dual: bool
    num_symbols: int
    embed: int
    enc_prenet: Tuple[int, int]
    conv_ch: int
    bank_k: int
    proj1: int
    proj2: int
    n_highway: int
    enc_lstm: int           # units per direction = cbhg_out_units // 2
    enc_sa: int             # encoder self-attention units (dual only)
    enc_sa_heads: int
    enc_sa_hops: int
    enc_sa_drop: float
    enc_prenet_drop: float
    dec_prenet: Tuple[int, int]
    dec_prenet_drop: float
    att_rnn: int            # attention LSTM units (LSTM-1)
    att1: int               # units of attention-1 score space
    att2: int               # units of attention-2 score space (dual only)
    att_kernel: int
    att_filters: int
    attention: str          # "forward" | "location_sensitive" | "additive"
    cumulative: bool
    transition_agent: bool
    dec_out: int            # LSTM-2/3 units
    dec_sa: int
    dec_sa_heads: int
    dec_sa_hops: int
    dec_sa_drop: float
    n_mels: int
    r: int
    n_feed: int
    zc: float
    zh: float
    use_speaker: bool
    num_speakers: int
    speaker_dim: int
    speaker_offset: int
    max_iters: int
    l2_weight: float = 0.0  # l2_regularization_weight when use_l2_regularization, else 0 (models/models.py:470-478)
    forced_alignment: bool = False   # use_forced_alignment_mode (models/models.py:411-427): EVAL / PREDICT decode with replayed alignments
    # PostNetV2 (models/models.py:92-100,440-462; tacotron2.tacotron.tacotron_v2.PostNetV2, RECALLED): postnet_layers x
    # [conv1d(kernel, channels, no bias) -> BN -> tanh (none on the last) -> dropout], Dense(num_mels), residual onto mel_output
    postnet_v2: bool = False
    postnet_layers: int = 5
    postnet_kernel: int = 5
    postnet_ch: int = 512
    postnet_drop: float = 0.5

    @property
    def mem1(self) -> int:      # depth of attention-1 memory (BiLSTM output)
        return 2 * self.enc_lstm

    @property
    def mem2(self) -> int:      # depth of attention-2 memory (encoder self-attention output)
        return self.enc_sa if self.dual else 0

    @property
    def ctx(self) -> int:       # concatenated context fed back into LSTM-1
        return self.mem1 + self.mem2

    @property
    def dec_in(self) -> int:    # teacher-forced decoder input width
        return self.n_mels * self.n_feed

    @property
    def out_units(self) -> int:
        return self.n_mels * self.r


def dims_from_hparams(hp) -> ModelDims:
    dual = hp.tacotron_model == "DualSourceSelfAttentionTacotronModel"
    if not dual and hp.tacotron_model != "ExtendedTacotronV1Model":
        raise ValueError(f"Unknown Tacotron model: {hp.tacotron_model}")
    if dual and hp.encoder != "SelfAttentionCBHGEncoder":
        raise ValueError(f"Unknown encoder: {hp.encoder}")
    if not dual and hp.encoder != "ZoneoutEncoderV1":
        raise ValueError(f"Unknown encoder: {hp.encoder}")
    if dual and hp.decoder != "DualSourceTransformerDecoder":
        raise ValueError(f"Unknown decoder: {hp.decoder}")
    if not dual and hp.decoder != "ExtendedDecoder":
        raise ValueError(f"Unknown decoder: {hp.decoder}")
    if hp.decoder_version != "v2":
        raise ValueError("only decoder_version=v2 (DecoderRNNV2) is on the hot path")
    if hp.attention not in ("forward", "location_sensitive", "additive"):
        raise ValueError(f"Unknown attention mechanism: {hp.attention}")
    if bool(getattr(hp, "use_forced_alignment_mode", False)) and (
            hp.forced_alignment_attention not in ("teacher_forcing_forward", "teacher_forcing_additive")
            or (dual and hp.forced_alignment_attention2 not in ("teacher_forcing_forward", "teacher_forcing_additive"))):
        raise ValueError("forced_alignment_attention[2] must be teacher_forcing_forward or teacher_forcing_additive (attentions.py:40-52)")
    if dual and hp.attention2 != "additive":
        raise ValueError("attention2 must be 'additive' on the hot path")
    if not hp.use_zoneout_at_encoder:
        raise ValueError("use_zoneout_at_encoder=False (plain CBHG with GRU) is out of scope")
    # switches of the reference's model_fn that this implementation does not build (SURVEY §8 f3 / out of scope): refuse loudly
    for flag, what in (("use_external_speaker_embedding", "external speaker embeddings (multi_speaker_tacotron)"),
                       ("use_language_embedding", "language embeddings (multi_speaker_tacotron)"),
                       ("speaker_embedd_to_postnet", "speaker embedding into the post-net"),
                       ("channel_id_to_postnet", "channel labels into the post-net"),
                       ("use_accent_type", "accent-type inputs")):
        if bool(getattr(hp, flag, False)):
            raise NotImplementedError(f"{flag}=True: {what} is not built in this implementation")
    # switches with a built default and an unbuilt alternative: the non-default value would train / predict a different model
    if bool(hp.use_speaker_embedding) and not bool(getattr(hp, "speaker_embedd_to_prenet", True)):
        raise NotImplementedError("speaker_embedd_to_prenet=False: the speaker embedding always feeds MultiSpeakerPreNet here "
                                  "(models/models.py:387, multi_speaker_modules.py:27-32)")
    for flag, what in (("speaker_embedd_to_decoder", "speaker embedding concatenated onto the encoder memories (models/models.py:366-372)"),
                       ("apply_dropout_on_inference", "pre-net dropout at inference time"),
                       ("language_embedd_to_input", "language embedding into the encoder input"),
                       ("language_embedd_to_decoder", "language embedding into the decoder")):
        if bool(getattr(hp, flag, False)):
            raise NotImplementedError(f"{flag}=True: {what} is not built in this implementation")
    if int(getattr(hp, "speaker_for_synthesis", -1)) > -1:
        raise NotImplementedError("speaker_for_synthesis > -1: overriding the speaker id at synthesis time is not built")
    # limits of the attention-RNN / LSTM sequence kernels (csrc/attn_rnn*.cu, lstm_seq.cu): fail at construction, not at the first step
    if hp.attention in ("forward", "location_sensitive") and (hp.attention_filters > 8 or hp.attention_kernel > 32):
        raise NotImplementedError(f"attention_filters={hp.attention_filters} / attention_kernel={hp.attention_kernel}: the location "
                                  "convolution of the attention kernels holds at most 8 filters of 32 taps (the shipped configurations use 5 x 10)")
    if hp.attention_out_units != 256 or hp.cbhg_out_units != 256 or hp.decoder_out_units != 256:
        raise NotImplementedError("attention_out_units / cbhg_out_units / decoder_out_units must be 256 (recurrent kernels are built for "
                                  "LSTM widths 128 (encoder) and 256)")
    if dual and (hp.attention1_out_units, hp.attention2_out_units, hp.self_attention_out_units) != (224, 32, 32):
        raise NotImplementedError("dual-source decoder: attention1_out_units / attention2_out_units / self_attention_out_units must be "
                                  "224 / 32 / 32")
    return ModelDims(
        dual=dual, num_symbols=hp.num_symbols, embed=hp.embedding_dim,
        enc_prenet=tuple(hp.encoder_prenet_out_units), conv_ch=hp.conv_channels, bank_k=hp.max_filter_width,
        proj1=hp.projection1_out_channels, proj2=hp.projection2_out_channels, n_highway=hp.num_highway,
        enc_lstm=hp.cbhg_out_units // 2,
        enc_sa=hp.self_attention_out_units, enc_sa_heads=hp.self_attention_num_heads,
        enc_sa_hops=hp.self_attention_num_hop, enc_sa_drop=hp.self_attention_drop_rate,
        enc_prenet_drop=hp.encoder_prenet_drop_rate,
        dec_prenet=tuple(hp.decoder_prenet_out_units), dec_prenet_drop=hp.decoder_prenet_drop_rate,
        att_rnn=hp.attention_out_units,
        att1=hp.attention1_out_units if dual else hp.attention_out_units,
        att2=hp.attention2_out_units if dual else 0,
        att_kernel=hp.attention_kernel, att_filters=hp.attention_filters, attention=hp.attention,
        cumulative=bool(hp.cumulative_weights), transition_agent=bool(hp.use_forward_attention_transition_agent),
        dec_out=hp.decoder_out_units,
        dec_sa=hp.decoder_self_attention_out_units, dec_sa_heads=hp.decoder_self_attention_num_heads,
        dec_sa_hops=hp.decoder_self_attention_num_hop, dec_sa_drop=hp.decoder_self_attention_drop_rate,
        n_mels=hp.num_mels, r=hp.outputs_per_step, n_feed=hp.n_feed_frame,
        zc=hp.zoneout_factor_cell, zh=hp.zoneout_factor_output,
        use_speaker=bool(hp.use_speaker_embedding), num_speakers=hp.num_speakers,
        speaker_dim=hp.speaker_embedding_dim, speaker_offset=hp.speaker_embedding_offset,
        max_iters=hp.max_iters, forced_alignment=bool(getattr(hp, "use_forced_alignment_mode", False)),
        postnet_v2=bool(getattr(hp, "use_postnet_v2", False)), postnet_layers=int(getattr(hp, "num_postnet_v2_layers", 5)),
        postnet_kernel=int(getattr(hp, "postnet_v2_kernel_size", 5)), postnet_ch=int(getattr(hp, "postnet_v2_out_channels", 512)),
        postnet_drop=float(getattr(hp, "postnet_v2_drop_rate", 0.5)),
        l2_weight=float(hp.l2_regularization_weight) if bool(getattr(hp, "use_l2_regularization", False)) else 0.0)


def l2_regularized(d: ModelDims, name: str) -> bool:
    """Is trainable tensor `name` part of l2_regularization_loss (models/models.py:470-478, regularizers.py:11-18)?  The reference
    black-lists TF variable names containing "embedding", "bias", "batch_normalization", "lstm_cell", the ExtendedDecoder's
    output / stop-token dense layers ("output_and_stop_token_wrapper/dense[_1]/", "output_projection_wrapper/kernel") and
    "stop_token_projection/kernel".  In this store's names: embeddings, every bias (`attention_bias` of forward_attention.py:21
    included), BN gamma / beta, every LSTM kernel, the stop projection, and — single-attention model only — the mel projection
    (the transformer decoder's "out_projection/kernel", module.py:718-720, matches no black-list entry and IS regularised)."""
    leaf = name.rsplit(".", 1)[-1]
    if "embedding" in name or leaf in ("gamma", "beta") or leaf.startswith("b") or ".lstm" in name:
        return False
    if name == "dec.stop_proj.W" or (name == "dec.out_proj.W" and not d.dual):
        return False
    return True


# init kinds: "glorot" (fan_in, fan_out from the last two dims; conv receptive field included),
# "zeros", "ones", "const:<v>", "embed" (truncated normal, sd 0.5)
ParamSpec = Tuple[str, Tuple[int, ...], str]


def param_specs(d: ModelDims) -> List[ParamSpec]:
    """Trainable tensors in forward order."""
    S: List[ParamSpec] = []
    S.append(("embedding", (d.num_symbols, d.embed), "embed"))                       # models.py:283
    if d.use_speaker:
        S.append(("speaker_embedding", (d.num_speakers, d.speaker_dim), "embed"))    # models.py:299
    cin = d.embed
    for i, u in enumerate(d.enc_prenet):                                             # module.py:394
        S += [(f"enc.prenet{i}.W", (cin, u), "glorot"), (f"enc.prenet{i}.b", (u,), "zeros")]
        cin = u
    c0 = cin
    # conv bank (module.py:46-53).  The 16 gamma / beta vectors are laid out back to back so the bank's
    # batch-norm runs as ONE [rows, K*conv_ch] pass over the concatenated conv outputs.
    for k in range(1, d.bank_k + 1):
        S.append((f"cbhg.bank{k}.W", (k, c0, d.conv_ch), "glorot"))
    for k in range(1, d.bank_k + 1):
        S.append((f"cbhg.bank{k}.gamma", (d.conv_ch,), "ones"))
    for k in range(1, d.bank_k + 1):
        S.append((f"cbhg.bank{k}.beta", (d.conv_ch,), "zeros"))
    S += [("cbhg.proj1.W", (3, d.conv_ch * d.bank_k, d.proj1), "glorot"),            # module.py:56-61
          ("cbhg.proj1.gamma", (d.proj1,), "ones"), ("cbhg.proj1.beta", (d.proj1,), "zeros"),
          ("cbhg.proj2.W", (3, d.proj1, d.proj2), "glorot"),                         # module.py:63-68
          ("cbhg.proj2.gamma", (d.proj2,), "ones"), ("cbhg.proj2.beta", (d.proj2,), "zeros")]
    hw = d.enc_lstm
    assert d.proj2 == c0 and d.proj2 == hw, "adjustment_layer path (module.py:88-89) not on the hot path"
    for i in range(d.n_highway):                                                     # module.py:72
        S += [(f"cbhg.highway{i}.WH", (hw, hw), "glorot"), (f"cbhg.highway{i}.bH", (hw,), "zeros"),
              (f"cbhg.highway{i}.WT", (hw, hw), "glorot"), (f"cbhg.highway{i}.bT", (hw,), "const:-1.0")]
    for dr in ("fw", "bw"):                                                          # module.py:93-108
        S += [(f"cbhg.lstm_{dr}.W", (hw + d.enc_lstm, 4 * d.enc_lstm), "glorot"),
              (f"cbhg.lstm_{dr}.b", (4 * d.enc_lstm,), "zeros")]
    if d.dual:
        S += [("enc.sa_proj.W", (d.mem1, d.enc_sa), "glorot"), ("enc.sa_proj.b", (d.enc_sa,), "zeros")]  # module.py:413
        for h in range(d.enc_sa_hops):                                               # module.py:415-423
            for nm in ("key", "value", "query", "output"):                           # self_attention.py:103-106
                S += [(f"enc.sa{h}.{nm}.W", (d.enc_sa, d.enc_sa), "glorot"), (f"enc.sa{h}.{nm}.b", (d.enc_sa,), "zeros")]
            S += [(f"enc.sa{h}.transform.W", (d.enc_sa, d.enc_sa), "glorot"),        # module.py:358
                  (f"enc.sa{h}.transform.b", (d.enc_sa,), "zeros")]
    # decoder pre-net (module.py:1506-1511 / multi_speaker_modules.py:19-23)
    if d.use_speaker:
        S += [("dec.prenet0.W0", (d.dec_in, d.dec_prenet[0]), "glorot"), ("dec.prenet0.b0", (d.dec_prenet[0],), "zeros"),
              ("dec.prenet0.Ws", (d.speaker_dim, d.dec_prenet[0]), "glorot"), ("dec.prenet0.bs", (d.dec_prenet[0],), "zeros"),
              ("dec.prenet0.W", (d.dec_prenet[0], d.dec_prenet[0]), "glorot"), ("dec.prenet0.b", (d.dec_prenet[0],), "zeros")]
    else:
        S += [("dec.prenet0.W", (d.dec_in, d.dec_prenet[0]), "glorot"), ("dec.prenet0.b", (d.dec_prenet[0],), "zeros")]
    S += [("dec.prenet1.W", (d.dec_prenet[0], d.dec_prenet[1]), "glorot"), ("dec.prenet1.b", (d.dec_prenet[1],), "zeros")]
    # attention 1 (forward_attention.py:50-86, 13-26 / TF BahdanauAttention)
    S += [("att1.memory.W", (d.mem1, d.att1), "glorot"), ("att1.query.W", (d.att_rnn, d.att1), "glorot")]
    if d.attention in ("forward", "location_sensitive"):
        S += [("att1.loc_conv.W", (d.att_kernel, 1, d.att_filters), "glorot"), ("att1.loc_conv.b", (d.att_filters,), "zeros"),
              ("att1.loc_layer.W", (d.att_filters, d.att1), "glorot"),
              ("att1.v", (d.att1,), "glorot"), ("att1.b", (d.att1,), "zeros")]
        if d.attention == "forward" and d.transition_agent:
            S += [("att1.agent.W", (d.mem1 + d.att1, 1), "glorot"), ("att1.agent.b", (1,), "zeros")]
    else:
        S += [("att1.v", (d.att1,), "glorot")]
    if d.dual:
        S += [("att2.memory.W", (d.mem2, d.att2), "glorot"), ("att2.query.W", (d.att_rnn, d.att2), "glorot"),
              ("att2.v", (d.att2,), "glorot")]
    # LSTM-1 (attention RNN), LSTM-2, LSTM-3 (DecoderRNNV2)
    S += [("dec.lstm1.W", (d.dec_prenet[1] + d.ctx + d.att_rnn, 4 * d.att_rnn), "glorot"),
          ("dec.lstm1.b", (4 * d.att_rnn,), "zeros"),
          ("dec.lstm2.W", (d.att_rnn + d.ctx + d.dec_out, 4 * d.dec_out), "glorot"),
          ("dec.lstm2.b", (4 * d.dec_out,), "zeros"),
          ("dec.lstm3.W", (d.dec_out + d.dec_out, 4 * d.dec_out), "glorot"),
          ("dec.lstm3.b", (4 * d.dec_out,), "zeros")]
    if d.dual:
        for h in range(d.dec_sa_hops):                                               # module.py:707-715
            for nm in ("key", "value", "query", "output"):
                S += [(f"dec.sa{h}.{nm}.W", (d.dec_sa, d.dec_sa), "glorot"), (f"dec.sa{h}.{nm}.b", (d.dec_sa,), "zeros")]
            S += [(f"dec.sa{h}.transform.W", (d.dec_sa, d.dec_sa), "glorot"), (f"dec.sa{h}.transform.b", (d.dec_sa,), "zeros")]
        proj_in = d.dec_sa
    else:
        proj_in = d.dec_out
    S += [("dec.out_proj.W", (proj_in, d.out_units), "glorot"), ("dec.out_proj.b", (d.out_units,), "zeros"),   # module.py:718-724
          ("dec.stop_proj.W", (proj_in, 1), "glorot"), ("dec.stop_proj.b", (1,), "zeros")]
    if d.postnet_v2:                                                                 # models/models.py:92-100,440-462
        cin = d.n_mels
        for i in range(d.postnet_layers):
            S += [(f"postnet.conv{i}.W", (d.postnet_kernel, cin, d.postnet_ch), "glorot"),
                  (f"postnet.conv{i}.gamma", (d.postnet_ch,), "ones"), (f"postnet.conv{i}.beta", (d.postnet_ch,), "zeros")]
            cin = d.postnet_ch
        S += [("postnet.proj.W", (cin, d.n_mels), "glorot"), ("postnet.proj.b", (d.n_mels,), "zeros")]
    return S


def bn_names(d: ModelDims) -> List[Tuple[str, int]]:
    """Batch-norm layers carrying (non-trainable) moving mean / variance."""
    out = [(f"cbhg.bank{k}", d.conv_ch) for k in range(1, d.bank_k + 1)]
    out += [("cbhg.proj1", d.proj1), ("cbhg.proj2", d.proj2)]
    if d.postnet_v2:
        out += [(f"postnet.conv{i}", d.postnet_ch) for i in range(d.postnet_layers)]
    return out


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= s
    return n


def num_trainable(d: ModelDims) -> int:
    return sum(_numel(s) for _, s, _ in param_specs(d))


class ParamStore:
    """Flat fp32 buffers (params / grads / Adam m, v) + named views + BN moving statistics."""

    def __init__(self, d: ModelDims, device="cpu", dtype=torch.float32):
        self.dims = d
        self.specs = param_specs(d)
        self.offsets: "OrderedDict[str, Tuple[int, Tuple[int, ...]]]" = OrderedDict()
        off = 0
        for name, shape, _ in self.specs:
            self.offsets[name] = (off, shape)
            off += (_numel(shape) + 3) // 4 * 4
        self.total = off
        self.flat = torch.zeros(off, device=device, dtype=dtype)
        self.grad = torch.zeros(off, device=device, dtype=dtype)
        self.adam_m = torch.zeros(off, device=device, dtype=dtype)
        self.adam_v = torch.zeros(off, device=device, dtype=dtype)
        self.p = self._views(self.flat)
        self.g = self._views(self.grad)
        # K-contiguous ("transposed") copy of every matrix: [..., in, out] -> [..., out, in]; the layout the tcgen05
        # GEMM tile wants for x.W products.  Refreshed by the engine after each optimiser step (one batched launch).
        self.flat_t = torch.zeros(off, device=device, dtype=dtype)
        self.pt = {n: self.flat_t[o:o + _numel(s)].view(s[:-2] + (s[-1], s[-2])) for n, (o, s) in self.offsets.items() if len(s) >= 2}
        desc = []
        for n, (o, s) in self.offsets.items():
            if len(s) >= 2:
                per = s[-2] * s[-1]
                for i in range(_numel(s[:-2])):
                    desc += [o + i * per, s[-2], s[-1]]
        self.t_desc = torch.tensor(desc, dtype=torch.int32, device=device)
        self.t_count = len(desc) // 3
        # BN moving statistics: one flat buffer per kind, same ordering as bn_names() (bank layers adjacent)
        self.bn: Dict[str, torch.Tensor] = {}
        tot = sum(c for _, c in bn_names(d))
        self.bn_mean_flat = torch.zeros(tot, device=device, dtype=dtype)
        self.bn_var_flat = torch.ones(tot, device=device, dtype=dtype)
        self.bn_off: Dict[str, int] = {}
        o = 0
        for name, c in bn_names(d):
            self.bn[name + ".mean"] = self.bn_mean_flat[o:o + c]
            self.bn[name + ".var"] = self.bn_var_flat[o:o + c]
            self.bn_off[name] = o
            o += c
        self.step = 0

    def l2_mask(self) -> torch.Tensor:
        """Flat {0, 1} mask of the l2-regularised parameters (``l2_regularized``), built on first use."""
        if getattr(self, "_l2_mask", None) is None:
            m = torch.zeros_like(self.flat)
            for n, (o, s) in self.offsets.items():
                if l2_regularized(self.dims, n):
                    m[o:o + _numel(s)] = 1.0
            self._l2_mask = m
        return self._l2_mask

    def _views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {n: flat[o:o + _numel(s)].view(s) for n, (o, s) in self.offsets.items()}

    def init(self, seed: int = 1234, mode: str = "glorot") -> "ParamStore":
        """``glorot``: the TF initialisers.  ``random``: every tensor, biases and BN
        statistics included, gets non-trivial values — used by parity tests so that no term is
        multiplied by an accidental zero."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for name, shape, kind in self.specs:
            t = torch.empty(shape, dtype=torch.float32)
            if kind == "glorot":
                if len(shape) == 1:
                    fan_in = fan_out = shape[0]
                else:
                    rf = _numel(shape[:-2])
                    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                t = (torch.rand(shape, generator=g) * 2 - 1) * lim
            elif kind == "embed":
                t = torch.randn(shape, generator=g).clamp_(-2, 2) * 0.5
            elif kind == "zeros":
                t = torch.zeros(shape) if mode == "glorot" else (torch.rand(shape, generator=g) - 0.5) * 0.2
            elif kind == "ones":
                t = torch.ones(shape) if mode == "glorot" else 1.0 + (torch.rand(shape, generator=g) - 0.5) * 0.4
            elif kind.startswith("const:"):
                c = float(kind.split(":")[1])
                t = torch.full(shape, c) if mode == "glorot" else c + (torch.rand(shape, generator=g) - 0.5) * 0.2
            else:
                raise ValueError(kind)
            self.p[name].copy_(t.to(self.flat.device, self.flat.dtype))
        for k in self.bn:
            if mode == "glorot":
                self.bn[k].fill_(0.0 if k.endswith(".mean") else 1.0)
            else:
                r = torch.rand(self.bn[k].shape, generator=g)
                self.bn[k].copy_(((r - 0.5) * 0.2 if k.endswith(".mean") else 0.5 + r).to(self.flat.device, self.flat.dtype))
        return self

    def to(self, device=None, dtype=None) -> "ParamStore":
        other = ParamStore(self.dims, device=device or self.flat.device, dtype=dtype or self.flat.dtype)
        other.flat.copy_(self.flat)
        other.grad.copy_(self.grad)
        other.adam_m.copy_(self.adam_m)
        other.adam_v.copy_(self.adam_v)
        for k in self.bn:
            other.bn[k].copy_(self.bn[k])
        other.step = self.step
        return other

    def as_dict(self) -> Dict[str, torch.Tensor]:
        """name -> tensor (trainable views + BN moving stats); the oracle's input format."""
        out = dict(self.p)
        out.update(self.bn)
        return out
