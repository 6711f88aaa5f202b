"""snt_b200 — B200-native Show-and-Tell caption-decoder hot path.

Import as `show_and_tell_b200` (the alias module at the repo root maps that name onto this directory,
whose own name is not a valid Python identifier).
"""
from . import _lib, ops, synthetic  # noqa: F401
from .models import DecoderRNN, EncoderCNN  # noqa: F401

from .trainer import CaptionModel, Trainer, evaluation  # noqa: F401

__all__ = ["EncoderCNN", "DecoderRNN", "CaptionModel", "Trainer", "evaluation", "ops", "synthetic"]
