"""Warm start by variable-name regular expressions (/root/reference/train.py:76-78, hparams.py:200-202:
``tf.estimator.WarmStartSettings(ckpt_to_initialize_from, vars_to_warm_start)``) and the map from this store's parameter names
to the TF1 variable names the reference's graph would give them.

TensorFlow is absent here, so no TF checkpoint can be read or written and the map cannot be validated against one: it
is a best-effort restatement.  In-tree anchors: the ``decoder`` variable scope and ``out_projection`` / ``stop_token_projection``
(modules/module.py:717-723), ``proj1`` / ``proj2`` (module.py:60,67), ``ForwardAttention`` with ``location_features_convolution``,
``location_features_layer``, ``transition_factor_projection``, ``attention_variable``, ``attention_bias``
(modules/forward_attention.py:17-26,58,73,78,86), ``memory_layer`` / ``query_layer`` / ``attention_v`` of TF's BahdanauAttention
(SURVEY A.8).  Everything else (layer-class default names of tacotron2 / tf.layers: ``dense``, ``conv1d``, ``batch_normalization``,
``lstm_cell``, ``kernel`` / ``bias`` / ``gamma`` / ``beta``) is RECALLED, and the ``_<k>`` suffixes TF appends to repeated layer names
are not reproduced.  A regular expression is therefore matched against BOTH names of a tensor: the TF-style one and the store's own
(``embedding``, ``enc.prenet0.W``, ``cbhg.bank3.W``, ``att1.query.W``, ``dec.lstm2.W`` ...), which is exact.
"""
from __future__ import annotations

import re
from collections import namedtuple
from typing import Iterable, List

WarmStartSettings = namedtuple("WarmStartSettings", ["ckpt_to_initialize_from", "vars_to_warm_start"])


def tf_variable_name(n: str, d) -> str:
    """TF1-style variable name of trainable tensor `n` (see the module docstring for what is anchored and what is recalled)."""
    leaf = n.rsplit(".", 1)[-1]
    kind = {"W": "kernel", "b": "bias", "gamma": "gamma", "beta": "beta", "v": "attention_v"}.get(leaf, leaf)
    if n == "embedding":
        return "embedding/embedding"
    if n == "speaker_embedding":
        return "speaker_embedding/embedding"
    parts = n.split(".")
    head = parts[0]
    if head == "enc":
        sub = parts[1]
        if sub.startswith("prenet"):
            return f"encoder/prenet_{sub[6:]}/dense/{kind}"
        if sub.startswith("lstm"):                                   # enc.lstm_fw / enc.lstm_bw
            return f"encoder/cbhg/bidirectional_rnn/{'fw' if 'fw' in sub else 'bw'}/zoneout_lstm_cell/lstm_cell/{kind}"
        if sub.startswith("sa"):                                     # encoder self-attention (module.py:363-371,425-438)
            inner = ".".join(parts[2:-1])
            return f"encoder/self_attention/{sub}/{inner}/{kind}" if inner else f"encoder/self_attention/{sub}/{kind}"
        return "encoder/" + "/".join(parts[1:-1]) + f"/{kind}"
    if head == "cbhg":
        sub = parts[1]
        if leaf in ("gamma", "beta"):
            return f"encoder/cbhg/{sub}/batch_normalization/{leaf}"
        if sub.startswith("highway"):
            return f"encoder/cbhg/{sub}/{'.'.join(parts[2:-1]) or 'dense'}/{kind}"
        return f"encoder/cbhg/{sub}/conv1d/{kind}"
    if head in ("att1", "att2"):
        scope = "decoder/ForwardAttention" if (head == "att1" and d.attention == "forward") else \
                "decoder/LocationSensitiveAttention" if (head == "att1" and d.attention == "location_sensitive") else "decoder/BahdanauAttention"
        if head == "att2":
            scope += "_1"
        sub = ".".join(parts[1:-1])
        if n.endswith(".v"):
            return f"{scope}/{'attention_variable' if (head == 'att1' and d.attention != 'additive') else 'attention_v'}"
        if n == "att1.b":
            return f"{scope}/attention_bias"
        names = {"memory": "memory_layer", "query": "query_layer", "loc_conv": "location_features_convolution",
                 "loc_layer": "location_features_layer", "agent": "transition_factor_projection"}
        return f"{scope}/{names.get(sub, sub)}/{kind}"
    if head == "dec":
        sub = parts[1]
        if sub.startswith("lstm"):
            return f"decoder/decoder_rnn/cell_{int(sub[4:]) - 1}/zoneout_lstm_cell/lstm_cell/{kind}"
        if sub.startswith("prenet"):
            return f"decoder/prenet_{sub[6:]}/{'.'.join(parts[2:-1]) or 'dense'}/{kind}".replace(f"/{leaf}/", "/")
        if sub == "out_proj":
            return f"decoder/{'out_projection' if d.dual else 'output_and_stop_token_wrapper/dense'}/{kind}"
        if sub == "stop_proj":
            return f"decoder/{'stop_token_projection' if d.dual else 'output_and_stop_token_wrapper/dense_1'}/{kind}"
        if sub.startswith("sa"):
            inner = ".".join(parts[2:-1])
            return f"decoder/self_attention/{sub}/{inner}/{kind}" if inner else f"decoder/self_attention/{sub}/{kind}"
        return "decoder/" + "/".join(parts[1:-1]) + f"/{kind}"
    if head == "postnet":                                            # models/models.py:92-100,440-462 (layer names RECALLED)
        sub = parts[1]
        if sub == "proj":
            return f"postnet_v2/dense/{kind}"
        if leaf in ("gamma", "beta"):
            return f"postnet_v2/conv1d_{int(sub[4:]) + 1}/batch_normalization/{leaf}"
        return f"postnet_v2/conv1d_{int(sub[4:]) + 1}/conv1d/{kind}"
    return n.replace(".", "/")


def select_warm_start(names: Iterable[str], d, vars_to_warm_start) -> List[str]:
    """Trainable tensors selected by ``vars_to_warm_start`` (a regular expression or a list of them; tf.estimator.WarmStartSettings
    semantics: ``re.match`` from the start of the name, any pattern of a list).  A pattern is tried on the TF-style name and on the
    store's own name of each tensor."""
    pats = [vars_to_warm_start] if isinstance(vars_to_warm_start, str) else list(vars_to_warm_start or [".*"])
    rx = [re.compile(p) for p in pats]
    return [n for n in names if any(r.match(tf_variable_name(n, d)) or r.match(n) for r in rx)]
