"""B200-native TASTE speech-tokenization path (log-mel -> Whisper encoder -> text-aligned aggregator -> RVQ).

Drop-in behind the reference's audio-tower interface (taste_speech.modeling_taste.TasteAudioTower, MT:33-211).
"""
from .synth import TowerConfig, FULL, SMALL, TINY  # noqa: F401

__all__ = ["TowerConfig", "FULL", "SMALL", "TINY"]
