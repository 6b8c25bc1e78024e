"""Second, independent set of pins for the oracle (round 2).  Run in the build container:
``python tests/golden/make_golden2.py`` -> second_pins.npz.  Nothing here imports oracle/ or the CUDA path.

  * ``transformers.audio_utils.spectrogram`` / ``mel_filter_bank`` (HuggingFace's NumPy feature extractor front end;
    Slaney scale + Slaney norm): power spectrogram and mel spectrogram of a seeded signal;
  * ``scipy.signal.stft`` / ``istft`` (a third STFT implementation, different framing code): complex STFT of the
    reflect-padded signal, rescaled from SciPy's 1/sum(window) convention, and the inverse;
  * the 'kaiser_best' resampler evaluated DIRECTLY from its closed form -- two wings of
    h(u) = s rolloff sinc(rolloff u) kaiser_beta(u / 64), s = min(1, ratio) -- in float64, no interpolation table:
    a full-length pin with the same (zero) edge handling as resampy's wings.
"""
import os

import numpy as np
import scipy.signal
import scipy.special
from transformers import audio_utils as au

HERE = os.path.dirname(os.path.abspath(__file__))
BETA, ROLLOFF, ZEROS = 14.769656459379492, 0.9475937167399596, 64


def signal(n, seed, sr=22050.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    y = 0.3 * np.sin(2 * np.pi * 440.0 * t) * np.exp(-3 * t) + 0.2 * np.sin(2 * np.pi * 1567.98 * t + 0.3)
    return (y + 0.01 * rng.standard_normal(n)).astype(np.float32)


def kaiser_best_direct(x, sr_orig, sr_new):
    """resampy 0.2.2 resample_f semantics without its table: output n sums a left wing (input samples floor(t), floor(t)-1,
    ...) and a right wing (floor(t)+1, ...) of the continuous filter.  One documented quirk is kept because it changes
    the result at the 1e-4 level for ratios like 44100/48000: the tap spacing inside the window is the INTEGER table
    step floor(s * 512) / 512, not s itself (resampy: ``index_step = int(scale * num_table)``)."""
    ratio = sr_new / sr_orig
    s = min(1.0, ratio)
    s_step = np.floor(s * 512) / 512.0
    n_out = int(len(x) * ratio)
    out = np.zeros(n_out)
    xm = x.astype(np.float64)
    i0b = scipy.special.i0(BETA)

    def h(u):
        w = np.zeros_like(u)
        inside = u < ZEROS
        w[inside] = scipy.special.i0(BETA * np.sqrt(np.maximum(0.0, 1.0 - (u[inside] / ZEROS) ** 2))) / i0b
        return ROLLOFF * np.sinc(ROLLOFF * u) * w * s

    for j in range(n_out):
        t = j / ratio
        n = int(t)
        frac = t - n
        i = np.arange(0, n + 1)
        u = s * frac + i * s_step
        keep = u < ZEROS
        out[j] += np.sum(xm[n - i[keep]] * h(u[keep]))
        k = np.arange(0, len(x) - n - 1)
        u = s * (1.0 - frac) + k * s_step
        keep = u < ZEROS
        out[j] += np.sum(xm[n + 1 + k[keep]] * h(u[keep]))
    res = np.zeros(int(np.ceil(len(x) * ratio)), dtype=np.float32)
    res[:n_out] = out.astype(np.float32)
    return res


def main():
    out = {}
    y = signal(9000, 11)
    out["y"] = y
    win = au.window_function(2048, "hann", periodic=True)
    for sr, hop in ((22050, 512), (44100, 256)):
        fb = au.mel_filter_bank(num_frequency_bins=1025, num_mel_filters=128, min_frequency=0.0, max_frequency=sr / 2.0,
                                sampling_rate=sr, norm="slaney", mel_scale="slaney")
        out[f"hf_fb_{sr}"] = fb.T.astype(np.float32)                                      # (128, 1025)
        out[f"hf_power_{hop}"] = au.spectrogram(y.astype(np.float64), win, 2048, hop, fft_length=2048, power=2.0, center=True,
                                                pad_mode="reflect", dtype=np.float64).astype(np.float32)
        out[f"hf_mel_{sr}_{hop}"] = au.spectrogram(y.astype(np.float64), win, 2048, hop, fft_length=2048, power=2.0, center=True,
                                                   pad_mode="reflect", mel_filters=fb, mel_floor=0.0,
                                                   dtype=np.float64).astype(np.float32)
    w = scipy.signal.get_window("hann", 2048, fftbins=True)
    for hop in (256, 512):
        yp = np.pad(y.astype(np.float64), 1024, mode="reflect")
        _, _, Z = scipy.signal.stft(yp, window=w, nperseg=2048, noverlap=2048 - hop, nfft=2048, boundary=None, padded=False)
        Z = Z * w.sum()
        out[f"scipy_stft_{hop}"] = Z.astype(np.complex64)
        _, xr = scipy.signal.istft(Z / w.sum(), window=w, nperseg=2048, noverlap=2048 - hop, nfft=2048, boundary=False)
        out[f"scipy_istft_{hop}"] = xr[1024:1024 + hop * (Z.shape[1] - 1)].astype(np.float32)
    t = np.arange(6000) / 44100.0
    x = (0.4 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 5000 * t + 1)
         + 0.01 * np.random.default_rng(5).standard_normal(6000)).astype(np.float32)
    out["rs_x"] = x
    for so, sn in ((44100, 22050), (48000, 44100), (22050, 44100)):
        out[f"rs_direct_{so}_{sn}"] = kaiser_best_direct(x, so, sn)
    np.savez_compressed(os.path.join(HERE, "second_pins.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
