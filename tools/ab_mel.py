"""Time the log-mel stage alone (C4 clip shape).  usage: [MST_MEL_RING_WAVES=k] python tools/ab_mel.py [n_clips]"""
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
audio = bench.make_audio_device(n, dev, 0)
batch = F.ClipBatch.uniform(n, bench.CLIP_LEN, bench.HOP, device=dev)
plan = F.MelPlan.get(bench.SR, device=dev)
ref = None
for layout in (F.BIN_MAJOR, F.FRAME_MAJOR):
    fn = lambda: F.melspectrogram_batch(audio, batch, plan, log1p=True, layout=layout)
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"waves {os.environ.get('MST_MEL_RING_WAVES', 'default')}: clips {n} layout {layout}: log-mel {np.median(ts):.3f} ms, "
          f"checksum {float(out.double().sum()):.6e}", flush=True)
