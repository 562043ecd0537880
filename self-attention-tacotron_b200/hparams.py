"""Config contract of the hot path ("Config contract").

The reference keeps one flat ``tf.contrib.training.HParams`` registry
(/root/reference/hparams.py:10-226) that is overridden first by a JSON file
(``parse_json``, /root/reference/train.py:111-114) and then by a ``k=v,k=v`` string
(``parse``, /root/reference/train.py:116).  TensorFlow is not available here, so this
module provides a small self-contained ``HParams`` with the same three entry points
(``parse_json`` / ``parse`` / ``values``) and a registry holding the same key names and
default values, so the reference's ``hparams.json`` files load unchanged.

Only the keys matter to the hot path; keys that configure out-of-scope subsystems
(pySpark preprocessing, tf.data tuning, vocoder features) are carried so that a JSON
file which sets them is still accepted.
"""
from __future__ import annotations

import copy
import json
import re
from typing import Any, Dict


class HParams:
    """Attribute bag with typed overrides (subset of tf.contrib.training.HParams behaviour)."""

    def __init__(self, **kwargs: Any):
        object.__setattr__(self, "_v", {})
        for k, v in kwargs.items():
            self.add_hparam(k, v)

    # -- registry -----------------------------------------------------------------
    def add_hparam(self, name: str, value: Any) -> None:
        if name in self._v:
            raise ValueError(f"Hyperparameter name is already registered: {name}")
        if isinstance(value, tuple):
            value = list(value)
        self._v[name] = value

    def set_hparam(self, name: str, value: Any) -> None:
        if name not in self._v:
            raise ValueError(f"Unknown hyperparameter: {name}")
        self._v[name] = self._coerce(name, value)

    def _coerce(self, name: str, value: Any) -> Any:
        cur = self._v[name]
        if isinstance(cur, list):
            if not isinstance(value, (list, tuple)):
                raise ValueError(f"Must pass a list for multi-valued hyperparameter: {name}")
            proto = cur[0] if cur else None
            return [self._cast(proto, x, name) for x in value]
        if isinstance(value, (list, tuple)):
            raise ValueError(f"Must not pass a list for single-valued hyperparameter: {name}")
        return self._cast(cur, value, name)

    @staticmethod
    def _cast(proto: Any, value: Any, name: str) -> Any:
        if proto is None:
            return value
        if isinstance(proto, bool):
            if isinstance(value, str):
                if value.lower() in ("true", "1"):
                    return True
                if value.lower() in ("false", "0"):
                    return False
                raise ValueError(f"Could not parse bool for {name}: {value}")
            return bool(value)
        if isinstance(proto, int):
            if isinstance(value, float) and value != int(value):
                raise ValueError(f"Could not cast {value} to int for {name}")
            return int(value)
        if isinstance(proto, float):
            return float(value)
        if isinstance(proto, str):
            return str(value)
        return value

    # -- overrides ----------------------------------------------------------------
    def override_from_dict(self, values: Dict[str, Any]) -> "HParams":
        for k, v in values.items():
            self.set_hparam(k, v)
        return self

    def parse_json(self, values_json: str) -> "HParams":
        return self.override_from_dict(json.loads(values_json))

    _ITEM = re.compile(r"\s*(?P<name>[A-Za-z_][A-Za-z0-9_]*)\s*=\s*(?P<val>\[[^\]]*\]|[^,\[]*)\s*(,|$)")

    def parse(self, values: str) -> "HParams":
        """``"a=1,b=[1,2],c=foo"`` overrides, as accepted by the reference CLI."""
        pos = 0
        values = values or ""
        while pos < len(values):
            m = self._ITEM.match(values, pos)
            if not m:
                raise ValueError(f"Malformed hyperparameter value: {values[pos:]}")
            pos = m.end()
            name, raw = m.group("name"), m.group("val").strip()
            if raw.startswith("["):
                body = raw[1:-1].strip()
                self.set_hparam(name, [x.strip() for x in body.split(",")] if body else [])
            else:
                self.set_hparam(name, raw)
        return self

    # -- access -------------------------------------------------------------------
    def values(self) -> Dict[str, Any]:
        return copy.deepcopy(self._v)

    def to_json(self, indent=None) -> str:
        return json.dumps(self._v, indent=indent, sort_keys=True)

    def __contains__(self, name: str) -> bool:
        return name in self._v

    def __getattr__(self, name: str) -> Any:
        try:
            return object.__getattribute__(self, "_v")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name: str, value: Any) -> None:
        if name in self._v:
            self.set_hparam(name, value)
        else:
            raise AttributeError(f"Unknown hyperparameter: {name} (use add_hparam)")

    def copy(self) -> "HParams":
        return HParams(**copy.deepcopy(self._v))


# Default registry.  Grouped by subsystem; names and defaults follow
# /root/reference/hparams.py:10-226 one for one (that file is the contract).
_DEFAULTS: Dict[str, Dict[str, Any]] = {
    "audio": dict(
        num_mels=80, num_mgcs=60, num_freq=2049, sample_rate=48000, frame_length_ms=50.0,
        frame_shift_ms=12.5, ref_level_db=20, average_mel_level_db=[0.0], stddev_mel_level_db=[0.0],
        min_mel_level_db=[0.0], silence_mel_level_db=-3.0),
    "vocoder_features(out of scope)": dict(
        mgc_dim=60, mgc_alpha=0.77, mgc_gamma=0.0, mgc_fft_len=4096,
        num_lf0s=256, f0_max=529.0, f0_min=66.0, lf0_loss_factor=0.5),
    "dataset": dict(
        dataset="vctk.dataset.DatasetSource", num_symbols=256, source="phoneme",
        source_file_extension="source.tfrecord", target_file_extension="target.tfrecord"),
    "model": dict(
        tacotron_model="ExtendedTacotronV1Model", outputs_per_step=2, n_feed_frame=2, embedding_dim=256),
    "accent(out of scope)": dict(
        use_accent_type=False, accent_type_embedding_dim=32, num_accent_type=129,
        accent_type_offset=0x3100, accent_type_unknown=0x3180, accent_type_prenet_out_units=[32, 16],
        encoder_prenet_out_units_if_accent=[224, 112]),
    "encoder": dict(
        encoder="ZoneoutEncoderV1", encoder_prenet_drop_rate=0.5, cbhg_out_units=256, conv_channels=128,
        max_filter_width=16, projection1_out_channels=128, projection2_out_channels=128, num_highway=4,
        encoder_prenet_out_units=[256, 128],
        encoder_v2_num_conv_layers=3, encoder_v2_kernel_size=5, encoder_v2_out_units=512,
        encoder_v2_drop_rate=0.5),
    "encoder_self_attention": dict(
        self_attention_out_units=32, self_attention_num_heads=2, self_attention_num_hop=1,
        self_attention_encoder_out_units=32, self_attention_drop_rate=0.05,
        self_attention_transformer_num_conv_layers=1, self_attention_transformer_kernel_size=5),
    "decoder": dict(
        decoder="ExtendedDecoder", attention="additive", forced_alignment_attention="teacher_forcing_forward",
        attention2="additive", forced_alignment_attention2="teacher_forcing_additive",
        attention1_out_units=224, attention2_out_units=32,
        decoder_prenet_drop_rate=0.5, apply_dropout_on_inference=False, decoder_prenet_out_units=[256, 128],
        attention_out_units=256, decoder_out_units=256,
        attention_kernel=31, attention_filters=32, cumulative_weights=False,
        use_forward_attention_transition_agent=False,
        decoder_self_attention_out_units=256, decoder_self_attention_num_heads=2,
        decoder_self_attention_num_hop=1, decoder_self_attention_drop_rate=0.05),
    "speaker": dict(
        use_speaker_embedding=False, use_external_speaker_embedding=False,
        speaker_embedding_projection_out_dim=-1, embedding_file="", num_speakers=1,
        speaker_embedding_dim=16, speaker_embedding_offset=0, speaker_for_synthesis=-1,
        speaker_embedd_to_prenet=True, speaker_embedd_to_decoder=False, speaker_embedd_to_postnet=False,
        channel_id_to_postnet=False, channel_id_file="", channel_id_dim=8,
        use_language_embedding=False, language_embedding_projection_out_dim=-1,
        language_embedding_file="", language_embedding_dim=16, language_embedd_to_input=False,
        language_embedd_to_decoder=False),
    "postnet": dict(
        post_net_cbhg_out_units=256, post_net_conv_channels=128, post_net_max_filter_width=8,
        post_net_projection1_out_channels=256, post_net_projection2_out_channels=80, post_net_num_highway=4,
        use_postnet_v2=False, num_postnet_v2_layers=5, postnet_v2_kernel_size=5,
        postnet_v2_out_channels=512, postnet_v2_drop_rate=0.5),
    "loss": dict(spec_loss_type="l1"),
    "training": dict(
        batch_size=32, adam_beta1=0.9, adam_beta2=0.999, adam_eps=1e-8, initial_learning_rate=0.002,
        decay_learning_rate=True, learning_rate_step_factor=1, use_l2_regularization=False,
        l2_regularization_weight=1e-7, save_summary_steps=100, save_checkpoints_steps=500,
        keep_checkpoint_max=200, keep_checkpoint_every_n_hours=1, log_step_count_steps=1,
        alignment_save_steps=10000, save_training_time_metrics=False, approx_min_target_length=100,
        suffle_buffer_size=64, batch_bucket_width=50, batch_num_buckets=50,
        interleave_cycle_length_cpu_factor=1.0, interleave_cycle_length_min=4,
        interleave_cycle_length_max=16, interleave_buffer_output_elements=200,
        interleave_prefetch_input_elements=200, prefetch_buffer_size=4, use_cache=False,
        cache_file_name="", logfile="log.txt", record_profile=False, profile_steps=50,
        warm_start=False, ckpt_to_initialize_from="", vars_to_warm_start=[".*"]),
    "eval": dict(
        max_iters=500, num_evaluation_steps=64, keep_eval_results_max_epoch=10,
        eval_start_delay_secs=120, eval_throttle_secs=600),
    "predict": dict(use_forced_alignment_mode=False, predicted_mel_extension="mfbsp"),
    "extension": dict(
        use_zoneout_at_encoder=False, decoder_version="v1", zoneout_factor_cell=0.1,
        zoneout_factor_output=0.1),
    "preprocess(out of scope)": dict(
        trim_top_db=30, trim_frame_length=1024, trim_hop_length=256, num_silent_frames=4),
}


def default_hparams() -> HParams:
    hp = HParams()
    for group in _DEFAULTS.values():
        for k, v in group.items():
            hp.add_hparam(k, copy.deepcopy(v))
    return hp


# Module-level singleton, mirroring ``from hparams import hparams`` in the reference CLIs
# (/root/reference/train.py:25, /root/reference/predict_mel.py:21).
hparams = default_hparams()


def hparams_debug_string(hp: HParams = None) -> str:
    """Same rendering as /root/reference/hparams.py:229-232."""
    values = (hp or hparams).values()
    lines = [f"  {name}: {values[name]}" for name in sorted(values)]
    return "Hyperparameters:\n" + "\n".join(lines)


def load_hparams(json_path: str = None, overrides: str = None) -> HParams:
    """defaults <- JSON file <- ``k=v`` string; the order used by train.py:111-116."""
    hp = default_hparams()
    if json_path:
        with open(json_path) as fh:
            hp.parse_json(fh.read())
    if overrides:
        hp.parse(overrides)
    return hp
