import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from ml_music_style_transfer_b200 import features as F
n = 8192
dev = torch.device("cuda", 0)
audio = bench.make_audio_device(n, dev, 0)
batch = F.ClipBatch.uniform(n, bench.CLIP_LEN, bench.HOP, device=dev)
gb = F.ClipBatch.from_frames([bench.T_FRAMES] * n, bench.HOP, device=dev)
plan = F.MelPlan.get(bench.SR, device=dev)
S = F.stft_batch(audio, batch, "magnitude", F.FRAME_MAJOR)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
t0 = timed(lambda: F.griffinlim_batch(S, gb, n_iter=0, seed=1, layout=F.FRAME_MAJOR))
t32 = timed(lambda: F.griffinlim_batch(S, gb, n_iter=32, seed=1, layout=F.FRAME_MAJOR))
tm = timed(lambda: F.melspectrogram_batch(audio, batch, plan, log1p=True, layout=F.BIN_MAJOR))
tp = timed(lambda: F.stft_batch(audio, batch, "log1p_power", F.FRAME_MAJOR))
print(f"GL iteration {(t32 - t0) / 32:.3f} ms, log-mel {tm:.3f} ms, log1p-power frame-major {tp:.3f} ms", flush=True)
