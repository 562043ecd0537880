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
    if getattr(d, "postnet_v2", False):              # dropout behind every PostNetV2 convolution, frame-major [T_mel, B, channels]
        for i in range(d.postnet_layers):
            m[f"postnet.conv{i}"] = (Td * d.r, B, d.postnet_ch)
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
    if name.startswith("postnet."):
        return 1.0 - d.postnet_drop
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


# ------------------------------------------------------------------------------------------------------------------
# The reference's input pipeline on TFRecord files (datasets/ljspeech/dataset.py:74-281, datasets/vctk/dataset.py) restated on numpy /
# torch.  Reading and parsing is in tfrecord.py; this part is `_prepare_target`, `group_by_batch` and `merge_target_to_source`.
# ------------------------------------------------------------------------------------------------------------------
def prepare_target(rec, hp):
    """`DatasetSource._prepare_target.convert` (datasets/ljspeech/dataset.py:127-167) for one utterance: normalise the mel by the
    corpus mean / stddev, add r silence frames at both ends, pad the tail with silence to a multiple of r, build done / masks.
    -> dict(mel [L,num_mels] float32, target_length L, done [L/r], spec_loss_mask [L], binary_loss_mask [L/r])."""
    import numpy as np
    r = hp.outputs_per_step
    mel = (np.asarray(rec.mel, np.float32) - np.asarray(hp.average_mel_level_db, np.float32)) / np.asarray(hp.stddev_mel_level_db,
                                                                                                        np.float32)
    sil = np.float32(hp.silence_mel_level_db)
    mel = np.pad(mel, ((r, r), (0, 0)), constant_values=sil)
    L = int(rec.target_length) + 2 * r                       # +2r for head and tail silence
    if L % r != 0:
        padded = (L // r + 1) * r
        mel = np.pad(mel, ((0, padded - L), (0, 0)), constant_values=sil)
        L = padded
    done = np.concatenate([np.zeros(L // r - 1, np.float32), np.ones(1, np.float32)])
    return dict(mel=mel.astype(np.float32), target_length=L, done=done, spec_loss_mask=np.ones(L, np.float32),
                binary_loss_mask=np.ones(L // r, np.float32))


def padded_batch(sources, targets, hp, device="cpu"):
    """`window.padded_batch` of datasets/ljspeech/dataset.py:247-281: pad every field to the longest item of the batch with the
    reference's padding values (source 0, mel = silence level, done 1, masks 0).  -> (SourceData, MelData) of torch tensors."""
    B = len(sources)
    Tt = max(len(s.source) for s in sources)
    Tm = max(t["target_length"] for t in targets)
    r = hp.outputs_per_step
    source = torch.zeros(B, Tt, dtype=torch.int64)
    mel = torch.full((B, Tm, hp.num_mels), float(hp.silence_mel_level_db))
    done = torch.ones(B, Tm // r)
    smask, bmask = torch.zeros(B, Tm), torch.zeros(B, Tm // r)
    for i, (s, t) in enumerate(zip(sources, targets)):
        source[i, :len(s.source)] = torch.from_numpy(s.source.astype("int64"))
        L = t["target_length"]
        mel[i, :L] = torch.from_numpy(t["mel"])
        done[i, :L // r] = torch.from_numpy(t["done"])
        smask[i, :L] = 1.0
        bmask[i, :L // r] = 1.0
    ids = torch.tensor([s.id for s in sources], dtype=torch.int64)
    keys = [s.key for s in sources]
    spk = None
    if getattr(hp, "use_speaker_embedding", False) and sources[0].speaker_id is not None:
        spk = torch.tensor([s.speaker_id for s in sources], dtype=torch.int64)
    dev = torch.device(device)
    mv = lambda x: x.to(dev) if x is not None else None      # noqa: E731
    feats = SourceData(mv(ids), keys, mv(source), mv(torch.tensor([s.source_length for s in sources], dtype=torch.int64)),
                       [s.text for s in sources], mv(spk))
    labels = MelData(mv(ids), keys, mv(mel), mv(torch.full((B,), hp.num_mels, dtype=torch.int64)),
                     mv(torch.tensor([t["target_length"] for t in targets], dtype=torch.int64)), mv(done), mv(smask), mv(bmask))
    return feats, labels


def group_by_batch(pairs, hp, batch_size=None):
    """`ZippedDataset.group_by_batch` (datasets/ljspeech/dataset.py:225-283): tf `group_by_window` with the reference's key function
    — bucket = min(num_buckets, min(target_length - approx_min_target_length, 0) // bucket_width), restated literally — and windows of
    5 batches; a full window is emitted as padded batches, the remaining windows are flushed at the end of the input."""
    bs = batch_size if batch_size is not None else hp.batch_size
    windows: Dict[int, list] = {}
    for src, tgt in pairs:
        key = min(hp.batch_num_buckets, min(tgt["target_length"] - hp.approx_min_target_length, 0) // hp.batch_bucket_width)
        w = windows.setdefault(key, [])
        w.append((src, tgt))
        if len(w) == bs * 5:
            for i in range(0, len(w), bs):
                yield padded_batch([x[0] for x in w[i:i + bs]], [x[1] for x in w[i:i + bs]], hp)
            windows[key] = []
    for key in list(windows):
        w = windows[key]
        for i in range(0, len(w), bs):
            yield padded_batch([x[0] for x in w[i:i + bs]], [x[1] for x in w[i:i + bs]], hp)


def tfrecord_input_fn(source_files, target_files, hp, batch_size=None, for_prediction=False, filter_max_output_length=None,
                      repeat=False, shuffle_buffer_size=0, seed=0, max_source_length=None):
    """input_fn over the reference's pre-processed TFRecord files (train.py:40-66 / predict_mel.py:39-45): source and target files
    are read in the given order and zipped record by record.  `for_prediction` applies `merge_target_to_source`
    (datasets/ljspeech/dataset.py:309-322): features become SourceDataForPrediction carrying the ground-truth mel.

    The stages of the reference's training / evaluation pipeline (train.py:53-54,65), in its order:
      * ``filter_max_output_length`` — `filter_by_max_output_length` (dataset.py:197-202): utterances whose padded target is longer
        than ``max_iters * outputs_per_step`` frames are dropped (None = the reference's behaviour: on for training / evaluation as
        in train.py, off with ``for_prediction`` — predict_mel.py does not filter);
      * ``repeat`` — `.repeat(count=None)`: start over when the files are exhausted (the caller bounds the run with ``steps``);
      * ``shuffle_buffer_size`` — `.shuffle(hparams.suffle_buffer_size)`: a sliding buffer of that many utterances, one drawn at
        random per step (tf.data semantics; ``seed`` makes the order reproducible); 0 = file order;
      * `group_by_batch` (bucketed, padded batches).
    ``hparams.source == 'phone'`` selects the phone sequence of VCTK records as the source (datasets/vctk/dataset.py:144-146).
    ``max_source_length`` (not in the reference, which has no T_text limit): drop utterances with more symbols than that — the
    attention-RNN kernels take T_text <= 192 and refuse longer batches loudly."""
    from . import tfrecord as TF
    import random

    max_out = int(hp.max_iters) * int(hp.outputs_per_step)
    use_phone = getattr(hp, "source", None) == "phone"
    do_filter = (not for_prediction) if filter_max_output_length is None else bool(filter_max_output_length)

    def gen():
        def pairs():
            while True:
                for sf, tf_ in zip(source_files, target_files):
                    for s, t in zip(TF.read_source_file(sf), TF.read_mel_file(tf_)):
                        if s.key != t.key:
                            raise ValueError(f"source / target records out of step: {s.key} vs {t.key}")
                        if use_phone:       # datasets/vctk/dataset.py:144-146: hparams.source == 'phone' selects the phone sequence
                            if s.phone is None:
                                raise ValueError(f"hparams.source='phone' but record {s.key} carries no phone / phone_length / phone_txt")
                            s = s._replace(source=s.phone, source_length=s.phone_length, text=s.phone_txt or "")
                        tgt = prepare_target(t, hp)
                        if do_filter and int(tgt["target_length"]) > max_out:
                            continue
                        if max_source_length is not None and int(s.source_length) > max_source_length:
                            continue
                        yield s, tgt
                if not repeat:
                    return

        def shuffled(it):
            if shuffle_buffer_size <= 1:
                yield from it
                return
            rng, buf = random.Random(seed), []
            for x in it:
                buf.append(x)
                if len(buf) >= shuffle_buffer_size:
                    yield buf.pop(rng.randrange(len(buf)))
            while buf:
                yield buf.pop(rng.randrange(len(buf)))

        for feats, labels in group_by_batch(shuffled(pairs()), hp, batch_size):
            if for_prediction:
                feats = SourceDataForPrediction(feats.id, feats.key, feats.source, feats.source_length, feats.text, feats.speaker_id,
                                                labels.mel, labels.mel_width, labels.target_length)
            yield feats, labels
    return gen
