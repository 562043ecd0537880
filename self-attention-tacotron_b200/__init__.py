"""satk — B200-native teacher-forced training hot path of Self-Attention Tacotron.

The directory name carries a hyphen (it mirrors the reference repository's name), so import it
with ``importlib.import_module("self-attention-tacotron_b200")``; ``tests/conftest.py`` and the
repo-root entry points register the alias ``satk`` in ``sys.modules``.
"""
from . import hparams as hparams_module            # noqa: F401
from .hparams import HParams, default_hparams, hparams, hparams_debug_string, load_hparams  # noqa: F401
from .params import ModelDims, ParamStore, dims_from_hparams, num_trainable, param_specs    # noqa: F401
from .data import (MelData, SourceData, SourceDataForPrediction, group_by_batch, make_masks, mask_shapes, padded_batch,  # noqa: F401
                   prepare_target, synthetic_batch, tfrecord_input_fn)
from . import tfrecord                               # noqa: F401
from .tf_names import WarmStartSettings, select_warm_start, tf_variable_name       # noqa: F401


def tacotron_model_factory(hparams, model_dir, run_config, warm_start_from=None, **kw):
    """models/models.py:1363-1381 (imported lazily: the model classes pull in the CUDA library binding)."""
    from .models import tacotron_model_factory as factory
    return factory(hparams, model_dir, run_config, warm_start_from, **kw)


__all__ = ["HParams", "default_hparams", "hparams", "load_hparams", "ModelDims", "ParamStore",
           "dims_from_hparams", "SourceData", "MelData", "synthetic_batch", "make_masks", "tacotron_model_factory", "tfrecord",
           "tfrecord_input_fn", "prepare_target", "padded_batch", "group_by_batch", "WarmStartSettings", "select_warm_start",
           "tf_variable_name"]
