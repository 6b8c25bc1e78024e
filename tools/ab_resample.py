"""Time the kaiser_best resampler for a few ratios, per-phase weight tables vs per-output table interpolation
(MST_RS_NO_PHASES=1).  usage: python tools/ab_resample.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_music_style_transfer_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
x = torch.randn(64 * 8 * 44100, device=dev) * 0.1


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), out


for so, sn in ((48000, 44100), (44100, 16000), (22050, 44100), (44100, 48000), (44100, 22050)):
    os.environ.pop("MST_RS_NO_PHASES", None)
    t_a, a = timed(lambda: L.ops().resample(x, so, sn))
    os.environ["MST_RS_NO_PHASES"] = "1"
    t_b, b = timed(lambda: L.ops().resample(x, so, sn))
    os.environ.pop("MST_RS_NO_PHASES", None)
    print(f"{so} -> {sn}: default {t_a:.3f} ms ({x.numel() / so / (t_a * 1e-3) / 1e6:.2f} M audio-s/s), per-output interpolation "
          f"{t_b:.3f} ms, max diff {float((a - b).abs().max()):.2e}", flush=True)
