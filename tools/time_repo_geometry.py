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

# Griffin-Lim at the reference's geometry: (1025, 860) chunks, hop 256 (inference.py:105 defaults), 32 iterations
n_gl = 512
gb = F.ClipBatch.from_frames([860] * n_gl, 256, device=dev)
S = F.stft_batch(audio, F.ClipBatch.uniform(n_gl, clen, 256, clip_stride=step, device=dev), "magnitude", F.FRAME_MAJOR)
t32 = timed(lambda: F.griffinlim_batch(S, gb, n_iter=32, seed=1, layout=F.FRAME_MAJOR), 3)
t0 = timed(lambda: F.griffinlim_batch(S, gb, n_iter=0, seed=1, layout=F.FRAME_MAJOR), 3)
it_ms = (t32 - t0) / 32
L = 256 * 859
alg = n_gl * (36 * 1025 * 860 + 8 * L)
print(f"Griffin-Lim hop 256: {t32:.2f} ms for {n_gl} chunks x 32 it -> {n_gl * L / 44100 / t32 * 1e3:.0f} audio-s/s; "
      f"iteration {it_ms:.3f} ms = {alg / it_ms / 1e6:.0f} GB/s algorithmic ({alg / it_ms / 1e6 / 6545.6:.3f} of the copy peak)")
