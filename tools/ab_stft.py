"""Time the log1p-power STFT in both layouts: C4 clip shape (hop 512, 173 frames) and the reference's chunk geometry
(44.1 kHz, hop 256, 860-frame chunks every 131 072 samples).  usage: [LD_PRELOAD=variant.so] python tools/ab_stft.py [n_clips]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ml_music_style_transfer_b200 import features as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
tag = os.environ.get("LD_PRELOAD", "/default/").split("/")[-2]


def timed(fn, reps=7):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(out.double().sum())


audio = bench.make_audio_device(n, dev, 0)
batch = F.ClipBatch.uniform(n, bench.CLIP_LEN, bench.HOP, device=dev)
for name, layout in (("frame-major", F.FRAME_MAJOR), ("bin-major", F.BIN_MAJOR)):
    t, cs = timed(lambda: F.stft_batch(audio, batch, "log1p_power", layout))
    print(f"{tag:8s} C4 x{n} {name}: {t:.3f} ms  checksum {cs:.6e}", flush=True)
del audio
n_ch, step_, clen = 512, 131072, 219904
a_repo = torch.randn((n_ch - 1) * step_ + clen, device=dev) * 0.1
b_repo = F.ClipBatch.uniform(n_ch, clen, 256, clip_stride=step_, device=dev)
for name, layout in (("frame-major", F.FRAME_MAJOR), ("bin-major", F.BIN_MAJOR)):
    t, cs = timed(lambda: F.stft_batch(a_repo, b_repo, "log1p_power", layout))
    print(f"{tag:8s} reference geometry x{n_ch} chunks {name}: {t:.3f} ms  checksum {cs:.6e}", flush=True)
