"""Times the reference's own preprocessing geometry (44.1 kHz, hop 256, 219 904-sample chunks every 131 072 samples,
bin-major log1p-power output = what process_audio_into_chunks returns) on the device."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_music_style_transfer_b200 import features as F  # noqa: E402

dev = torch.device("cuda", 0)
n_chunks, step, clen = 512, 131072, 219904
audio = 0.1 * torch.randn((n_chunks - 1) * step + clen, device=dev)
b = F.ClipBatch.uniform(n_chunks, clen, 256, clip_stride=step, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for layout, name in ((F.BIN_MAJOR, "bin-major (reference stack)"), (F.FRAME_MAJOR, "frame-major (librosa order)")):
    ms = timed(lambda: F.stft_batch(audio, b, "log1p_power", layout))
    frames = b.total_frames
    byts = n_chunks * 4 * clen + 4 * 1025 * frames
    print(f"{name}: {ms:.3f} ms for {n_chunks} chunks ({frames} frames) -> {n_chunks * clen / 44100 / ms * 1e3:.0f} audio-s/s, "
          f"{frames / ms * 1e3 / 1e6:.1f} Mframes/s, {byts / ms / 1e6:.0f} GB/s algorithmic")
