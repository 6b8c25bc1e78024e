"""Time one Griffin-Lim iteration launch ((t[32 it] - t[0 it]) / 32) and the log-mel / log-power stages on the C4 clip shape.
Used for A/B runs of library variants:  LD_PRELOAD=.../variants/<name>/libmst_b200.so python tools/ab_gl.py [n_clips]"""
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
gb = F.ClipBatch.from_frames([bench.T_FRAMES] * n, bench.HOP, device=dev)
plan = F.MelPlan.get(bench.SR, device=dev)
S = F.stft_batch(audio, batch, "magnitude", F.FRAME_MAJOR)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


t0 = timed(lambda: F.griffinlim_batch(S, gb, n_iter=0, seed=1, layout=F.FRAME_MAJOR))
t32 = timed(lambda: F.griffinlim_batch(S, gb, n_iter=32, seed=1, layout=F.FRAME_MAJOR))
tm = timed(lambda: F.melspectrogram_batch(audio, batch, plan, log1p=True, layout=F.BIN_MAJOR))
tp = timed(lambda: F.stft_batch(audio, batch, "log1p_power", F.FRAME_MAJOR))
tb = timed(lambda: F.stft_batch(audio, batch, "log1p_power", F.BIN_MAJOR))
from ml_music_style_transfer_b200 import pianoroll as PR  # noqa: E402
notes = PR.NoteBatch(*bench.make_notes(n, 99), device=dev)
roll, onoff, row_off, _ = PR.rasterize(notes, bench.ROLL_FS)


def up():
    for s0 in range(0, n, 1024):
        PR.upsample_pair(roll, onoff, row_off[s0:min(n, s0 + 1024) + 1], bench.CLIP_LEN, bench.ROLL_FS, bench.SR, 21, 88, torch.int8)


tu = timed(up)
print(f"{os.environ.get('LD_PRELOAD', 'default').split('/')[-2] if os.environ.get('LD_PRELOAD') else 'default':8s} clips {n}: "
      f"GL iteration {(t32 - t0) / 32:.3f} ms (x{16384 / n:.0f} = {(t32 - t0) / 32 * 16384 / n:.2f} ms @16384), GL-32 {t32:.1f} ms, "
      f"log-mel {tm:.3f} ms, log1p-power frame-major {tp:.3f} ms, bin-major {tb:.3f} ms, GL init+final {t0:.2f} ms, "
      f"upsample pair {tu:.2f} ms = {2 * n * 88 * bench.CLIP_LEN / tu / 1e6:.0f} GB/s", flush=True)
