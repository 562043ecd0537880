"""satk — B200-native teacher-forced training hot path of Self-Attention Tacotron.

The directory name carries a hyphen (it mirrors the reference repository's name), so import it
with ``importlib.import_module("self-attention-tacotron_b200")``; ``tests/conftest.py`` and the
repo-root entry points register the alias ``satk`` in ``sys.modules``.
"""
from . import hparams as hparams_module            # noqa: F401
from .hparams import HParams, default_hparams, hparams, hparams_debug_string, load_hparams  # noqa: F401
from .params import ModelDims, ParamStore, dims_from_hparams, num_trainable, param_specs    # noqa: F401
from .data import MelData, SourceData, SourceDataForPrediction, make_masks, mask_shapes, synthetic_batch  # noqa: F401

__all__ = ["HParams", "default_hparams", "hparams", "load_hparams", "ModelDims", "ParamStore",
           "dims_from_hparams", "SourceData", "MelData", "synthetic_batch", "make_masks"]
