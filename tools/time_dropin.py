"""Wall-clock of the NumPy-in / NumPy-out drop-in calls a user of the reference makes (host round trips included)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ml_music_style_transfer_b200 as mst  # noqa: E402
from oracle import preprocess as opp  # noqa: E402

pp = mst.preprocess
n_chunks = 100
audio = (0.1 * np.random.default_rng(0).standard_normal((n_chunks - 1) * 131072 + 219904)).astype(np.float32)


def wall(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


t, out = wall(lambda: pp.process_audio_into_chunks(audio, "cuba", 1, n_chunks))
print(f"process_audio_into_chunks (100 chunks, {audio.size / 44100:.0f} s of audio -> {out.nbytes / 1e6:.0f} MB): {t * 1e3:.1f} ms wall")
spec = out[0]
t, w = wall(lambda: mst.inference.AudioSynthesizer().griffinlim(spec, "x", n_iter=300), reps=3)
print(f"AudioSynthesizer.griffinlim (1025x860, 300 iterations): {t * 1e3:.1f} ms wall")
t0 = time.perf_counter()
ref = opp.process_audio_into_chunks(audio[:2 * 131072 + 219904], "cuba", 1, 3)
print(f"oracle (NumPy) process_audio_into_chunks: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per chunk")
