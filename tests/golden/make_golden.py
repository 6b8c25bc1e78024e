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
    # resampling: torchaudio's sinc_interp_kaiser with the three 'kaiser_best' parameters (what torchaudio documents as
    # the librosa-compatible configuration); independent of resampy's tabulated-window evaluation restated in oracle/
    t = np.arange(16000) / 44100.0
    x = (0.4 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 5000 * t + 1)
         + 0.01 * np.random.default_rng(5).standard_normal(16000)).astype(np.float32)
    rs = {"x": x}
    for so, sn in ((44100, 22050), (48000, 44100), (22050, 44100)):
        rs[f"y_{so}_{sn}"] = torchaudio.functional.resample(
            torch.from_numpy(x).double(), so, sn, lowpass_filter_width=64, rolloff=0.9475937167399596,
            resampling_method="sinc_interp_kaiser", beta=14.769656459379492).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "resample_torchaudio.npz"), **rs)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
