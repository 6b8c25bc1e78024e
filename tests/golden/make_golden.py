"""Regenerates the golden fixtures in this directory from implementations that are INDEPENDENT of oracle/ and of the
CUDA path: torch.stft / torch.istft (float64), torchaudio.functional.melscale_fbanks (Slaney/Slaney) and
torchaudio.functional.griffinlim (rand_init=False).  Run in the build container: ``python tests/golden/make_golden.py``.
librosa / pretty_midi (what the reference itself calls) are not installable here, so these are the pins.
"""
import os

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))


def signal(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 22050.0
    y = 0.3 * np.sin(2 * np.pi * 440.0 * t) * np.exp(-3 * t) + 0.2 * np.sin(2 * np.pi * 1567.98 * t + 0.3)
    return (y + 0.01 * rng.standard_normal(n)).astype(np.float32)


def main():
    y = signal(6000, 7)
    win = torch.hann_window(2048, periodic=True, dtype=torch.float64)
    out = {"y": y}
    for hop in (256, 512):
        for pad in ("reflect", "constant"):
            D = torch.stft(torch.from_numpy(y).double(), 2048, hop, window=win, center=True, pad_mode=pad,
                           return_complex=True).numpy()
            out[f"stft_{hop}_{pad}"] = D.astype(np.complex64)
        D = torch.stft(torch.from_numpy(y).double(), 2048, hop, window=win, center=True, pad_mode="reflect",
                       return_complex=True)
        out[f"istft_{hop}"] = torch.istft(D, 2048, hop, window=win, center=True).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "stft_torch.npz"), **out)

    mel = {}
    for sr in (22050, 44100):
        mel[f"fb_{sr}"] = torchaudio.functional.melscale_fbanks(1025, 0.0, sr / 2.0, 128, sr, norm="slaney",
                                                                 mel_scale="slaney").numpy().T.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "mel_torchaudio.npz"), **mel)

    mag = np.abs(out["stft_512_reflect"]).astype(np.float32)
    gl = {"mag": mag}
    for mom in (0.99, 0.0):
        w = torchaudio.functional.griffinlim(torch.from_numpy(mag).double(), win, 2048, 512, 2048, 1.0, 8, mom, None, False)
        gl[f"y_iter8_mom{mom}"] = w.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "griffinlim_torchaudio.npz"), **gl)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
