"""cpmusic — B200-native compound-word linear-transformer agent.

Importable as ``cpmusic`` (repo-root shim) or via
``importlib.import_module("reinforcement-learning-in-music-generation_b200")``.
"""
from . import _lib, ops, encoder, model, rl, rollout, dist, data, graphs, midi  # noqa: F401
from .encoder import (TransformerEncoderBuilder, RecurrentEncoderBuilder, TriangularCausalMask, LengthMask, FullMask,  # noqa: F401
                      install_fast_transformers_shim)
from .model import (CPLinearTransformer, TransformerModel, LinearTransformer, Actor_Transformer,  # noqa: F401
                    Critic_Transformer)
from .rollout import RolloutEngine  # noqa: F401
from .graphs import GraphedTrainStep  # noqa: F401
from .ops import manual_seed  # noqa: F401

__version__ = "0.1.0"
