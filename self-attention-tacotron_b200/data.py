"""Tensor contract of a training batch + synthetic LJSpeech/VCTK-shaped batches.

``SourceData`` / ``MelData`` carry the same fields as the reference's tf.data pipeline emits
(/root/reference/datasets/ljspeech/dataset.py:25-37; speaker fields from
/root/reference/datasets/vctk/dataset.py:36-38).  Padding values follow ``group_by_batch``
(dataset.py:247-281): source id 0, mel = ``silence_mel_level_db``, done = 1, masks = 0; the target
layout follows ``_prepare_target`` (dataset.py:127-167): r leading and r trailing silence frames,
length rounded up to a multiple of r, ``done`` = one-hot of the last decoder step.

There is no network and no TFRecord corpus here, so batches are synthetic with the shapes and
value ranges of the real thing.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Dict, Optional

import torch

SourceData = namedtuple("SourceData", ["id", "key", "source", "source_length", "text", "speaker_id"])
MelData = namedtuple("MelData", ["id", "key", "mel", "mel_width", "target_length", "done",
                                 "spec_loss_mask", "binary_loss_mask"])
SourceDataForPrediction = namedtuple("SourceDataForPrediction",
                                     ["id", "key", "source", "source_length", "text", "speaker_id",
                                      "mel", "mel_width", "target_length"])


def synthetic_batch(hp, batch_size: int, t_text: int, t_mel: int, seed: int = 1234, device="cpu",
                    full_length: bool = False):
    """One padded batch: (SourceData, MelData).  ``t_mel`` must be a multiple of outputs_per_step."""
    r = hp.outputs_per_step
    assert t_mel % r == 0
    g = torch.Generator(device="cpu").manual_seed(seed)
    B = batch_size
    # text: ids 1..67 (/root/reference/preprocess/text.py:21-38), 0 = pad / silence at both ends
    src_len = torch.randint(max(2, int(0.6 * t_text + 0.999)), t_text + 1, (B,), generator=g)
    src_len[0] = t_text
    if full_length:
        src_len[:] = t_text
    source = torch.randint(1, 68, (B, t_text), generator=g)
    pos = torch.arange(t_text)[None, :]
    source[:, 0] = 0
    source[torch.arange(B), src_len - 1] = 0
    source = torch.where(pos < src_len[:, None], source, torch.zeros_like(source))
    # target: N(0,1) clipped to +-4 (mean/std normalised mel), silence frames = silence_mel_level_db
    lo = max(2 * r + r, int(0.6 * t_mel) // r * r)
    tgt_len = (torch.randint(lo // r, t_mel // r + 1, (B,), generator=g) * r)
    tgt_len[B - 1] = t_mel
    if full_length:
        tgt_len[:] = t_mel
    mel = torch.randn(B, t_mel, hp.num_mels, generator=g).clamp_(-4, 4)
    tpos = torch.arange(t_mel)[None, :]
    sil = (tpos < r) | (tpos >= (tgt_len[:, None] - r))
    mel = torch.where(sil[:, :, None], torch.full_like(mel, float(hp.silence_mel_level_db)), mel)
    spec_mask = (tpos < tgt_len[:, None]).float()
    dpos = torch.arange(t_mel // r)[None, :]
    dlen = (tgt_len // r)[:, None]
    done = (dpos >= dlen - 1).float()          # last valid step = 1, padding = 1 (dataset.py:156-157,276)
    bin_mask = (dpos < dlen).float()
    spk = None
    if hp.use_speaker_embedding:
        spk = torch.randint(hp.speaker_embedding_offset, hp.speaker_embedding_offset + hp.num_speakers, (B,), generator=g)
    ids = torch.arange(B)
    keys = [f"SYN{seed:04d}-{i:04d}" for i in range(B)]
    dev = torch.device(device)
    feats = SourceData(ids.to(dev), keys, source.to(dev), src_len.to(dev), [""] * B,
                       None if spk is None else spk.to(dev))
    labels = MelData(ids.to(dev), keys, mel.to(dev), torch.full((B,), hp.num_mels).to(dev), tgt_len.to(dev),
                     done.to(dev), spec_mask.to(dev), bin_mask.to(dev))
    return feats, labels


def mask_shapes(d, B: int, Tt: int, Td: int) -> Dict[str, tuple]:
    """Shapes of every Bernoulli keep-mask a TRAIN-mode step consumes (1 = keep).

    prenet dropout (tacotron2 PreNet), zoneout on c and h of every ZoneoutLSTMCell
    (A.6), attention-probability dropout (self_attention.py:61)."""
    m = {}
    # every mask is TIME-MAJOR (the layout the kernels use); attention masks are [B, heads, Tq, Tk]
    for i, u in enumerate(d.enc_prenet):
        m[f"enc.prenet{i}"] = (Tt, B, u)
    for dr in ("fw", "bw"):
        m[f"cbhg.lstm_{dr}.c"] = (Tt, B, d.enc_lstm)
        m[f"cbhg.lstm_{dr}.h"] = (Tt, B, d.enc_lstm)
    if d.dual:
        for h in range(d.enc_sa_hops):
            m[f"enc.sa{h}"] = (B, d.enc_sa_heads, Tt, Tt)
        for h in range(d.dec_sa_hops):
            m[f"dec.sa{h}"] = (B, d.dec_sa_heads, Td, Td)
    for i, u in enumerate(d.dec_prenet):
        m[f"dec.prenet{i}"] = (Td, B, u)
    for i, u in ((1, d.att_rnn), (2, d.dec_out), (3, d.dec_out)):
        m[f"dec.lstm{i}.c"] = (Td, B, u)
        m[f"dec.lstm{i}.h"] = (Td, B, u)
    return m


def mask_keep_prob(d, name: str) -> float:
    if name.startswith("enc.prenet"):
        return 1.0 - d.enc_prenet_drop
    if name.startswith("dec.prenet"):
        return 1.0 - d.dec_prenet_drop
    if name.startswith("enc.sa"):
        return 1.0 - d.enc_sa_drop
    if name.startswith("dec.sa"):
        return 1.0 - d.dec_sa_drop
    if name.endswith(".c"):
        return 1.0 - d.zc
    if name.endswith(".h"):
        return 1.0 - d.zh
    raise KeyError(name)


def make_masks(d, B: int, Tt: int, Td: int, seed: int = 99, device="cpu") -> Dict[str, torch.Tensor]:
    """Host-generated uint8 keep-masks (parity tests feed the SAME masks to oracle and CUDA path)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for name, shape in mask_shapes(d, B, Tt, Td).items():
        out[name] = (torch.rand(shape, generator=g) < mask_keep_prob(d, name)).to(torch.uint8).to(device)
    return out
