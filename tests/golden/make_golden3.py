"""Pins for the oracle at FFT sizes other than 2048 and at win_length < n_fft (round 2).  Independent of oracle/ and of
the CUDA path: torch.stft / torch.istft (float64), torchaudio.functional.melscale_fbanks, torchaudio.functional.griffinlim
(rand_init=False).  Run in the build container: ``python tests/golden/make_golden3.py`` -> other_nfft_pins.npz."""
import os

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))


def signal(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 22050.0
    y = 0.3 * np.sin(2 * np.pi * 523.25 * t) * np.exp(-2 * t) + 0.2 * np.sin(2 * np.pi * 2093.0 * t + 0.7)
    return (y + 0.01 * rng.standard_normal(n)).astype(np.float32)


def main():
    y = signal(5000, 13)
    out = {"y": y}
    yt = torch.from_numpy(y).double()
    for n_fft, hop, win_length in ((1024, 256, 1024), (4096, 1024, 4096), (512, 100, 512), (1024, 256, 800), (2048, 512, 1200)):
        win = torch.hann_window(win_length, periodic=True, dtype=torch.float64)   # torch centre-pads it to n_fft, like librosa
        tag = f"{n_fft}_{hop}_{win_length}"
        D = torch.stft(yt, n_fft, hop, win_length=win_length, window=win, center=True, pad_mode="reflect", return_complex=True)
        out[f"stft_{tag}"] = D.numpy().astype(np.complex64)
        if n_fft % hop == 0:
            out[f"istft_{tag}"] = torch.istft(D, n_fft, hop, win_length=win_length, window=win, center=True).numpy().astype(np.float32)
        if (n_fft, win_length) in ((1024, 1024), (1024, 800)):
            mag = D.abs()
            w = torchaudio.functional.griffinlim(mag, win, n_fft, hop, win_length, 1.0, 6, 0.99, None, False)
            out[f"gl6_{tag}"] = w.numpy().astype(np.float32)
    for sr, n_fft, n_mels in ((22050, 1024, 80), (44100, 4096, 128)):
        out[f"fb_{sr}_{n_fft}_{n_mels}"] = torchaudio.functional.melscale_fbanks(
            n_fft // 2 + 1, 0.0, sr / 2.0, n_mels, sr, norm="slaney", mel_scale="slaney").numpy().T.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "other_nfft_pins.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
