"""msa_b200 — B200-native (sm_100a) implementation of the audio-features -> fusion hot path of
Joaonic/multimodal-sentiment-analyzer, behind the reference's own Python interface.

    from msa_b200 import AudioAnalyzer, AdvancedFusionModel, AudioAnalysis

``AudioAnalyzer`` mirrors src/analyzers/audio_analyzer.py, ``AdvancedFusionModel`` mirrors
src/models/fusion_model.py; both call hand-written CUDA kernels through the C ABI declared in
include/msa_b200.h (libmsa_b200.so, loaded with ctypes).  There is no CPU fallback: importing
works anywhere, but any compute call without the built library and a CUDA device raises.
"""
from .structures import AudioAnalysis, DictMixin
from .audio_analyzer import AudioAnalyzer
from .ingest import (AudioFeatureNormalizer, FaceFeatureNormalizer, FeatureNormalizer, TextFeatureNormalizer, assemble_row,
                     resample)
from .fusion_model import AdvancedFusionModel, FusionModel
from .pipeline import SegmentPipeline, aggregate_speakers, shard_range
from .streaming import StreamingWindow

__all__ = ["AudioAnalysis", "DictMixin", "AudioAnalyzer", "AudioFeatureNormalizer", "FaceFeatureNormalizer", "TextFeatureNormalizer", "FeatureNormalizer", "assemble_row",
           "resample", "AdvancedFusionModel",
           "FusionModel", "SegmentPipeline", "aggregate_speakers", "shard_range", "StreamingWindow"]
