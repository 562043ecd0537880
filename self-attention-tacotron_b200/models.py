"""Estimator surface of the hot path ("Estimator surface").

Mirrors what ``train.py`` / ``predict_mel.py`` touch in the reference:
  * ``tacotron_model_factory(hparams, model_dir, run_config, warm_start_from=None)``
    (/root/reference/models/models.py:1363-1381) selecting by ``hparams.tacotron_model``;
  * objects with ``model_fn(features, labels, mode, params)`` returning a spec that carries
    ``loss`` / ``train_op`` / ``predictions`` (models.py:278,514,562,588), and ``train`` /
    ``evaluate`` / ``predict`` drivers (tf.estimator.Estimator's public methods used at
    train.py:88 and predict_mel.py:54).
TensorFlow's session/graph machinery is replaced by eager calls into the CUDA engine; the tensors
flowing through ``features`` / ``labels`` keep the reference's field names (data.py).
"""
from __future__ import annotations

import os
from collections import namedtuple
from typing import Callable, Dict, Iterable, Optional

import torch

from .data import MelData, SourceData
from .engine import TacotronEngine
from .tf_names import WarmStartSettings, select_warm_start, tf_variable_name  # noqa: F401


class ModeKeys:
    TRAIN = "train"
    EVAL = "eval"
    PREDICT = "infer"


EstimatorSpec = namedtuple("EstimatorSpec", ["mode", "loss", "train_op", "predictions", "eval_metric_ops", "scalars"])
EstimatorSpec.__new__.__defaults__ = (None, None, None, None, None)


def _to_device(nt, device):
    """Host -> device copy of every tensor field (non-blocking: pinned host buffers overlap with the host's enqueue work).
    (A copy stream of its own was tried: no measurable gain at 8.4 MB per step, and `record_stream` makes the caching allocator defer
    the reuse of the input blocks, which showed up as occasional 20 ms steps.)"""
    return type(nt)(*[x.to(device, non_blocking=True) if torch.is_tensor(x) else x for x in nt])


class _TacotronEstimator:
    """Common driver: owns the engine (weights, optimiser state) and the checkpoint directory."""

    _model_name = ""

    def __init__(self, params, model_dir: Optional[str] = None, config=None, warm_start_from: Optional[str] = None,
                 device: Optional[str] = None, allreduce: Optional[Callable[[torch.Tensor], None]] = None, world_size: int = 1):
        if params.tacotron_model != self._model_name:
            raise ValueError(f"{type(self).__name__} built with tacotron_model={params.tacotron_model}")
        self.params = params
        self.model_dir = model_dir
        self.config = config
        self.device = torch.device(device or "cuda")
        self.engine = TacotronEngine(params, self.device)
        self._allreduce = allreduce
        self._world_size = world_size
        if model_dir and os.path.exists(self._ckpt_path()):
            self.restore(self._ckpt_path())          # resume = re-run with the same --checkpoint-dir (train.py:70-80); a checkpoint in
                                                      # model_dir wins over the warm start, as in tf.estimator
        elif isinstance(warm_start_from, WarmStartSettings):
            # train.py:76-78: only the trainable tensors whose name matches vars_to_warm_start are initialised from the checkpoint;
            # optimiser state, BN moving statistics and the step counter start fresh
            self.warm_start(warm_start_from.ckpt_to_initialize_from, warm_start_from.vars_to_warm_start)
        elif warm_start_from:
            self.restore(warm_start_from)

    # ---- checkpointing (flat buffers; a TF-variable name map is a later row of SURVEY §8f)
    def _ckpt_path(self) -> str:
        return os.path.join(self.model_dir, "model.satk.pt")

    def save(self, path: Optional[str] = None) -> str:
        path = path or self._ckpt_path()
        os.makedirs(os.path.dirname(path), exist_ok=True)
        ps = self.engine.ps
        torch.save(dict(flat=ps.flat.cpu(), adam_m=ps.adam_m.cpu(), adam_v=ps.adam_v.cpu(), bn_mean=ps.bn_mean_flat.cpu(),
                        bn_var=ps.bn_var_flat.cpu(), global_step=self.engine.global_step,
                        names=list(ps.offsets.keys()), offsets=[o for o, _ in ps.offsets.values()],
                        shapes=[tuple(sh) for _, sh in ps.offsets.values()]), path)
        return path

    def restore(self, path: str) -> None:
        st = torch.load(path, map_location="cpu")
        ps = self.engine.ps
        if st["names"] != list(ps.offsets.keys()) or st["flat"].numel() != ps.flat.numel():
            raise ValueError(f"checkpoint {path} does not match this model's parameter set")
        ps.flat.copy_(st["flat"]); ps.adam_m.copy_(st["adam_m"]); ps.adam_v.copy_(st["adam_v"])
        ps.bn_mean_flat.copy_(st["bn_mean"]); ps.bn_var_flat.copy_(st["bn_var"])
        self.engine.global_step = int(st["global_step"])
        self.engine.refresh_transposed()

    def _labels_of_prediction_record(self, features) -> MelData:
        """MelData view of a SourceDataForPrediction record (datasets/ljspeech/dataset.py:40-49,309-322: mel, mel_width, target_length)."""
        mel = getattr(features, "mel", None)
        if mel is None:
            raise ValueError("forced-alignment mode needs the ground-truth mel of the prediction records (SourceDataForPrediction.mel)")
        r = self.engine.d.r
        B, Tm = mel.shape[0], mel.shape[1]
        tl = features.target_length.to(mel.device)
        pos = torch.arange(Tm, device=mel.device)[None, :]
        spec_mask = (pos < tl[:, None]).float()
        dpos = torch.arange(Tm // r, device=mel.device)[None, :]
        dl = (tl // r)[:, None]
        return MelData(features.id, features.key, mel, features.mel_width, tl, (dpos >= dl - 1).float(), spec_mask, (dpos < dl).float())

    def _forced_alignment_decode(self, features, labels):
        eng, d = self.engine, self.engine.d
        sfeat = SourceData(features.id, features.key, features.source, features.source_length, features.text, features.speaker_id)
        tf_out = eng.forward(sfeat, labels, False)
        forced = (tf_out["align1_tm"].clone(), tf_out["align2_tm"].clone() if d.dual else None)
        return eng.predict(sfeat, max_iters=labels.mel.shape[1] // d.r, use_stop_token=False, forced_alignments=forced)

    def warm_start(self, path: str, vars_to_warm_start=".*") -> list:
        """tf.estimator.WarmStartSettings(ckpt_to_initialize_from=path, vars_to_warm_start=...) (train.py:76-78, hparams.py:200-202): copy
        the selected trainable tensors from a checkpoint of this implementation (`model.satk.pt`; TF checkpoints cannot be read without
        TensorFlow).  The checkpoint may belong to a model with a different parameter set: tensors are matched by name and shape.
        Returns the names that were initialised."""
        if os.path.isdir(path):
            path = os.path.join(path, "model.satk.pt")
        st = torch.load(path, map_location="cpu")
        ps = self.engine.ps
        if "offsets" in st:
            src_off = {n: (o, tuple(sh)) for n, o, sh in zip(st["names"], st["offsets"], st["shapes"])}
        else:                # checkpoint written before the shape table existed: it must be this model's parameter set
            if st["names"] != list(ps.offsets.keys()) or st["flat"].numel() != ps.flat.numel():
                raise ValueError(f"warm start: {path} has no shape table and does not match this model's parameter set")
            src_off = {n: (o, tuple(sh)) for n, (o, sh) in ps.offsets.items()}
        done = []
        for n in select_warm_start(ps.offsets.keys(), self.engine.d, vars_to_warm_start):
            o, sh = ps.offsets[n]
            if n in src_off and tuple(src_off[n][1]) == tuple(sh):
                so = src_off[n][0]
                ps.p[n].copy_(st["flat"][so:so + ps.p[n].numel()].view(sh))
                done.append(n)
        self.engine.refresh_transposed()
        return done

    # ---- model_fn (models.py:278 / :23)
    def model_fn(self, features, labels, mode, params=None, masks=None) -> EstimatorSpec:
        eng, d = self.engine, self.engine.d
        features = _to_device(features, self.device)
        if mode == ModeKeys.PREDICT:
            if d.forced_alignment:
                # forced-alignment mode (models/models.py:411-427, predict_mel.py with use_forced_alignment_mode): a teacher-forced pass
                # over the ground-truth mel of the prediction record gives the alignments, a second decode that feeds back its own
                # output replays them (TeacherForcingForwardAttention / TeacherForcingAdditiveAttention) for the target length
                out = self._forced_alignment_decode(features, self._labels_of_prediction_record(features))
            else:
                # free-running decode (predict_mel.py:36-74; decoder branch module.py:762-778), stop-token terminated
                out = eng.predict(features, max_iters=getattr(params or self.params, "max_iters", None))
            preds = {"id": features.id, "key": features.key, "mel": out["mel"], "stop_token": out["stop"],
                     "alignment": out["alignment"], "source": features.source, "text": features.text}
            if d.postnet_v2:
                preds["mel_postnet"] = out["mel_postnet"]                        # models/models.py:210
            gt = getattr(features, "mel", None)
            if gt is None and labels is not None:
                gt = labels.mel
            if gt is not None:
                preds["ground_truth_mel"] = gt
            if d.dual:
                preds["alignment2"] = out["alignment2"]
                for i, a in enumerate(out["dec_self_P"]):
                    preds[f"alignment{3 + i}"] = a.transpose(1, 2)
                for i, a in enumerate(out["enc_self_P"]):
                    preds[f"alignment{5 + i}"] = a.transpose(1, 2)
            return EstimatorSpec(mode=mode, predictions=preds)
        labels = _to_device(labels, self.device)
        training = mode == ModeKeys.TRAIN
        B, Tm = labels.mel.shape[0], labels.mel.shape[1]
        Td = Tm // d.r
        if training:
            if d.forced_alignment:
                raise NotImplementedError("use_forced_alignment_mode in TRAIN mode (a second, alignment-replaying decode that is "
                                          "differentiated together with the first) is not built; EVAL and PREDICT are")
            out = eng.train_step(features, labels, masks, allreduce=self._allreduce, world_size=self._world_size)
            losses = out["losses"]
            scalars = {"mel_loss": losses[0], "done_loss": losses[1], "learning_rate": out["lr"]}
            return EstimatorSpec(mode=mode, loss=losses[2], train_op=eng.global_step, predictions=None, eval_metric_ops=None, scalars=scalars)
        # EVAL (models/models.py:384-395,500-545): `loss` / `mel_loss` / `done_loss` and the mel + alignments handed to MetricsSaver come
        # from a decode WITHOUT teacher forcing over the target length (ValidationHelper(teacher_forcing=False)); a second, teacher-forced
        # decode gives the `*_with_teacher` metrics
        tf_out = eng.forward(features, labels, False)
        with_teacher = tf_out["losses"].clone()
        post_with_teacher = tf_out["postnet_v2_mel_loss"].clone() if d.postnet_v2 else None
        if d.forced_alignment:   # models/models.py:411-427: the plain metrics / predictions come from the alignment-replaying decode
            forced = (tf_out["align1_tm"].clone(), tf_out["align2_tm"].clone() if d.dual else None)
            losses, out = eng.validate(features, labels, forced_alignments=forced)
        else:
            losses, out = eng.validate(features, labels)
        preds = {"id": features.id, "key": features.key, "mel": out["mel"], "ground_truth_mel": labels.mel, "stop_token": out["stop"],
                 "alignment": out["alignment"],                                  # (B, Tt, Td), models.py:406
                 "source": features.source, "text": features.text}
        if d.dual:
            preds["alignment2"] = out["alignment2"]                              # models.py:407
            for i, a in enumerate(out["dec_self_P"]):                            # alignment3/4, models.py:403-404
                preds[f"alignment{3 + i}"] = a.transpose(1, 2)
            for i, a in enumerate(out["enc_self_P"]):                            # alignment5.., models.py:398
                preds[f"alignment{5 + i}"] = a.transpose(1, 2)
        scalars = {"loss": losses[2], "mel_loss": losses[0], "done_loss": losses[1],
                   "loss_with_teacher": with_teacher[2], "mel_loss_with_teacher": with_teacher[0], "done_loss_with_teacher": with_teacher[1]}
        if d.postnet_v2:                                                         # models/models.py:174-187,260-269
            B_, Td_ = labels.mel.shape[0], labels.mel.shape[1] // d.r
            preds["mel_postnet"] = out["mel_postnet_tm"].view(Td_, B_, d.r, d.n_mels).permute(1, 0, 2, 3).reshape(B_, Td_ * d.r, d.n_mels)
            scalars["postnet_v2_mel_loss"] = out["postnet_v2_mel_loss"][0]
            scalars["postnet_v2_mel_loss_with_teacher"] = post_with_teacher[0]
        return EstimatorSpec(mode=mode, loss=losses[2], train_op=None, predictions=preds, eval_metric_ops=dict(scalars), scalars=scalars)

    # ---- drivers (tf.estimator.Estimator.train / evaluate / predict)
    def train(self, input_fn: Callable[[], Iterable], steps: Optional[int] = None, max_steps: Optional[int] = None, hooks=None):
        n = 0
        last = None
        for features, labels in input_fn():
            if steps is not None and n >= steps:
                break
            if max_steps is not None and self.engine.global_step >= max_steps:
                break
            last = self.model_fn(features, labels, ModeKeys.TRAIN, self.params)
            n += 1
            every = getattr(self.params, "save_checkpoints_steps", 0)
            if self.model_dir and every and self.engine.global_step % every == 0:
                self.save()
        if self.model_dir and n:
            self.save()
        return last

    def evaluate(self, input_fn: Callable[[], Iterable], steps: Optional[int] = None) -> Dict[str, float]:
        sums: Dict[str, float] = {}
        n = 0
        for features, labels in input_fn():
            if steps is not None and n >= steps:
                break
            spec = self.model_fn(features, labels, ModeKeys.EVAL, self.params)
            for k, v in spec.eval_metric_ops.items():
                sums[k] = sums.get(k, 0.0) + float(v)
            n += 1
        return {k: v / max(n, 1) for k, v in sums.items()} | {"global_step": self.engine.global_step}

    def predict(self, input_fn: Callable[[], Iterable], checkpoint_path: Optional[str] = None):
        if checkpoint_path:
            self.restore(checkpoint_path)
        for item in input_fn():
            # tf.estimator.Estimator.predict runs model_fn in PREDICT mode on the features alone (predict_mel.py:54); a
            # (features, labels) pair is accepted so the training input_fn can be reused — labels only add ground_truth_mel
            features, labels = item if (isinstance(item, tuple) and len(item) == 2 and not hasattr(item, "_fields")) else (item, None)
            yield self.model_fn(features, labels, ModeKeys.PREDICT, self.params).predictions


class DualSourceSelfAttentionTacotronModel(_TacotronEstimator):
    """models/models.py:275"""
    _model_name = "DualSourceSelfAttentionTacotronModel"


class ExtendedTacotronV1Model(_TacotronEstimator):
    """models/models.py:20"""
    _model_name = "ExtendedTacotronV1Model"


def tacotron_model_factory(hparams, model_dir, run_config, warm_start_from=None, **kw):
    """models/models.py:1363-1381 (the MGC/LF0 vocoder-parameter models are out of scope)."""
    if hparams.tacotron_model == "DualSourceSelfAttentionTacotronModel":
        return DualSourceSelfAttentionTacotronModel(hparams, model_dir, config=run_config, warm_start_from=warm_start_from, **kw)
    if hparams.tacotron_model == "ExtendedTacotronV1Model":
        return ExtendedTacotronV1Model(hparams, model_dir, config=run_config, warm_start_from=warm_start_from, **kw)
    if hparams.tacotron_model in ("MgcLf0TacotronModel", "DualSourceSelfAttentionMgcLf0TacotronModel"):
        raise NotImplementedError(f"{hparams.tacotron_model}: MGC/LF0 models are outside the hot-path scope")
    raise ValueError(f"Unknown Tacotron model: {hparams.tacotron_model}")
