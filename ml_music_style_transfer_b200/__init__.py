"""B200-native (sm_100a) preprocessing / inversion hot path of silburt/ML_Music_Style_Transfer.

  preprocess  -- drop-in names of preprocessing/preprocess.py (STFT -> log1p power, chunking, piano roll)
  inference   -- AudioSynthesizer.griffinlim of model/inference.py
  features    -- librosa-shaped operators (stft, melspectrogram, griffinlim) + batched raw ops
  pianoroll   -- pretty_midi-shaped rasteriser, chunker, audio-rate upsampler
  pipeline    -- multi-stream host-buffer pipeline over the whole path (HostPipeline)
  sharding    -- one-process-per-GPU partitioning + optional NCCL gather

Everything computes in hand-written CUDA kernels behind ``torch.ops.mst_b200`` (C ABI: include/mst_b200.h).
There is no CPU fallback.
"""
from . import _lib, audio_io, dataset, features, pipeline, pianoroll, preprocess, inference, sharding, midi  # noqa: F401

__all__ = ["audio_io", "dataset", "features", "pianoroll", "preprocess", "inference", "sharding", "midi"]
