"""Small driver for ncu captures: runs each stage of the hot path a few times on a modest batch.
usage: python tools/prof_driver.py [n_clips] [stage ...]   stages: mel logpower gl roll"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ml_music_style_transfer_b200 import features as F, pianoroll as PR  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
stages = sys.argv[2:] or ["mel", "logpower", "gl", "roll"]   # also: logpower_bm (bin-major), repo (reference geometry, T = 860)
dev = torch.device("cuda", 0)
audio = bench.make_audio_device(n, dev, 0)
batch = F.ClipBatch.uniform(n, bench.CLIP_LEN, bench.HOP, device=dev)
gl_batch = F.ClipBatch.from_frames([bench.T_FRAMES] * n, bench.HOP, device=dev)
plan = F.MelPlan.get(bench.SR, device=dev)
notes = PR.NoteBatch(*bench.make_notes(n, 99), device=dev)
S = F.stft_batch(audio, batch, "magnitude", F.FRAME_MAJOR)
for rep in range(3):
    if "mel" in stages:
        F.melspectrogram_batch(audio, batch, plan, log1p=True, layout=F.BIN_MAJOR)
    if "logpower" in stages:
        F.stft_batch(audio, batch, "log1p_power", F.FRAME_MAJOR)
    if "logpower_bm" in stages:
        F.stft_batch(audio, batch, "log1p_power", F.BIN_MAJOR)
    if "repo" in stages:
        if rep == 0:
            n_ch, step_, clen = 256, 131072, 219904
            a_repo = torch.randn((n_ch - 1) * step_ + clen, device=dev) * 0.1
            b_repo = F.ClipBatch.uniform(n_ch, clen, 256, clip_stride=step_, device=dev)
        F.stft_batch(a_repo, b_repo, "log1p_power", F.BIN_MAJOR)
        F.stft_batch(a_repo, b_repo, "log1p_power", F.FRAME_MAJOR)
    if "gl" in stages:
        F.griffinlim_batch(S, gl_batch, n_iter=4, seed=1, layout=F.FRAME_MAJOR)
    if "roll" in stages:
        roll, onoff, ro, _ = PR.rasterize(notes, bench.ROLL_FS)
        PR.upsample_pair(roll, onoff, ro, bench.CLIP_LEN, bench.ROLL_FS, bench.SR, 21, 88, torch.int8)
    torch.cuda.synchronize()
print("ok")
