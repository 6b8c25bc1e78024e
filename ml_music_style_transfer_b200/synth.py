"""Seeded synthetic corpora for tests and benchmarks.  Host-side NumPy; not on the product path."""
import numpy as np


def piano_clip(clip_id, seconds, sr, seed=1234):
    """Piano-like clip: 1-6 decaying harmonic notes + -60 dB noise, peak 0.5, float32 mono."""
    rng = np.random.default_rng(seed + int(clip_id))
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    y = np.zeros(n, dtype=np.float64)
    for _ in range(int(rng.integers(1, 7))):
        pitch = int(rng.integers(21, 109))
        f0 = 440.0 * 2.0 ** ((pitch - 69) / 12.0)
        tau = rng.uniform(0.2, 2.0)
        onset = rng.uniform(0.0, max(1e-3, 0.6 * seconds))
        env = np.where(t >= onset, np.exp(-(t - onset) / tau), 0.0)
        for h in range(1, 9):
            if f0 * h < 0.45 * sr:
                y += env * np.sin(2 * np.pi * f0 * h * (t - onset) + rng.uniform(0, 2 * np.pi)) / h
    y += 1e-3 * rng.standard_normal(n)
    y *= 0.5 / max(np.abs(y).max(), 1e-9)
    return y.astype(np.float32)


def noise_clip(clip_id, n_samples, seed=1234):
    return (0.1 * np.random.default_rng(seed + int(clip_id)).standard_normal(n_samples)).astype(np.float32)


def midi_piece(piece_id, seconds=30.0, seed=99, notes_per_second=12.0):
    """Poisson onsets, pitch ~ N(64,14^2) clipped to [21,108], duration ~ Exp(0.4) in [0.03,4], velocity U{1..127}."""
    rng = np.random.default_rng(seed + int(piece_id))
    n = int(rng.poisson(notes_per_second * seconds))
    start = np.sort(rng.uniform(0.0, seconds, n))
    dur = np.clip(rng.exponential(0.4, n), 0.03, 4.0)
    end = np.minimum(start + dur, seconds)
    pitch = np.clip(np.rint(rng.normal(64, 14, n)), 21, 108).astype(np.int32)
    vel = rng.integers(1, 128, n).astype(np.int32)
    return pitch, vel, start.astype(np.float64), end.astype(np.float64)
