"""Robustness soak (not part of the test suite): long Griffin-Lim runs at scale, extreme amplitudes, NaN/Inf checks."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ml_music_style_transfer_b200 import features as F  # noqa: E402

dev = torch.device("cuda", 0)
n = 1024
audio = bench.make_audio_device(n, dev, 3)
batch = F.ClipBatch.uniform(n, bench.CLIP_LEN, 512, device=dev)
gb = F.ClipBatch.from_frames([bench.T_FRAMES] * n, 512, device=dev)
S = F.stft_batch(audio, batch, "magnitude", F.FRAME_MAJOR)
yb = F.ClipBatch.uniform(n, 512 * (bench.T_FRAMES - 1), 512, device=dev)
prev = None
for it in (0, 8, 32, 100, 300):
    y = F.griffinlim_batch(S, gb, n_iter=it, seed=11, layout=F.FRAME_MAJOR)
    assert torch.isfinite(y).all(), it
    sc = F.spectral_convergence_batch(y, yb, S, F.FRAME_MAJOR)
    print(f"n_iter {it:3d}: spectral convergence mean {sc.mean().item():.4f} max {sc.max().item():.4f}")
    if prev is not None:
        assert sc.mean().item() < prev + 1e-3
    prev = sc.mean().item()
plan = F.MelPlan.get(22050, device=dev)
for scale in (1e-6, 1.0, 1e3, 1e6):
    m = F.melspectrogram_batch(audio * scale, batch, plan, log1p=False, layout=F.BIN_MAJOR)
    p = F.stft_batch(audio * scale, batch, "log1p_power", F.FRAME_MAJOR)
    assert torch.isfinite(m).all() and torch.isfinite(p).all() and (m >= 0).all(), scale
    print(f"scale {scale:g}: mel max {m.max().item():.3e}, log1p-power max {p.max().item():.3f}")
y2 = F.griffinlim_batch(S, gb, n_iter=32, seed=11, layout=F.FRAME_MAJOR)
y3 = F.griffinlim_batch(S, gb, n_iter=32, seed=11, layout=F.FRAME_MAJOR)
print("Griffin-Lim bitwise reproducible across runs:", bool(torch.equal(y2, y3)))
print("soak ok")
