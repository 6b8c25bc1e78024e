"""librosa-0.8 ``stft`` / ``istft`` semantics restated in NumPy (oracle; test infrastructure).

Reference call sites: preprocessing/preprocess.py:48 (``librosa.stft(chunk, n_fft=2048,
hop_length=256)``), tests/plot_spec.py:14, model/inference.py:110 and
tests/test_griffinlim.py:23 (``librosa.griffinlim`` -> istft/stft pairs).

Upstream semantics followed (librosa 0.8.0 ``core/spectrum.py``; librosa is not vendored in
/root/reference and is absent from this image):
  * window = scipy.signal.get_window('hann', win_length, fftbins=True) (periodic Hann, float64),
    centre-padded to n_fft;
  * center=True -> np.pad(y, n_fft // 2, mode=pad_mode), default pad_mode='reflect';
  * frame t starts at padded index t*hop, T = 1 + len(y) // hop;
  * rfft of (float64 window * float32 frame) evaluated in float64, stored as complex64,
    result shape (1 + n_fft//2, T), Fortran order;
  * istft: irfft in float64, times window, overlap-added frame by frame into a float32 buffer
    of n_fft + hop*(T-1) samples, divided by the window sum-square envelope where that
    exceeds tiny(float32), trimmed by n_fft//2 on both sides.
"""
import numpy as np

__all__ = ["hann_window", "pad_signal", "stft", "istft", "window_sumsquare", "frame_count"]


def hann_window(win_length, n_fft=None):
    """Periodic Hann in float64, zero-padded symmetrically to n_fft (librosa util.pad_center)."""
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    if n_fft is not None and n_fft != win_length:
        lpad = (n_fft - win_length) // 2
        w = np.pad(w, (lpad, n_fft - win_length - lpad))
    return w


def frame_count(n_samples, hop_length):
    return 1 + int(n_samples) // int(hop_length)


def pad_signal(y, n_fft, pad_mode="reflect"):
    return np.pad(y, int(n_fft // 2), mode=pad_mode)


def stft(y, n_fft=2048, hop_length=None, win_length=None, center=True, pad_mode="reflect",
         out_dtype=np.complex64):
    y = np.asarray(y)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    w = hann_window(win_length, n_fft).reshape(-1, 1)
    if center:
        y = pad_signal(y, n_fft, pad_mode)
    if len(y) < n_fft:
        raise ValueError("input too short for n_fft")
    n_frames = 1 + (len(y) - n_fft) // hop_length
    # strided frame view, (n_fft, T)
    frames = np.lib.stride_tricks.as_strided(
        y, shape=(n_fft, n_frames), strides=(y.strides[0], y.strides[0] * hop_length), writeable=False)
    out = np.empty((1 + n_fft // 2, n_frames), dtype=out_dtype, order="F")
    blk = 256
    for s in range(0, n_frames, blk):
        out[:, s:s + blk] = np.fft.rfft(w * frames[:, s:s + blk], axis=0)
    return out


def window_sumsquare(n_frames, hop_length, win_length=None, n_fft=2048, dtype=np.float32):
    if win_length is None:
        win_length = n_fft
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = hann_window(win_length, n_fft) ** 2
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))]
    return x


def istft(D, hop_length=None, win_length=None, center=True, dtype=np.float32, length=None):
    D = np.asarray(D)
    n_fft = 2 * (D.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    w = hann_window(win_length, n_fft).reshape(-1, 1)
    if length:
        padded_length = length + int(n_fft) if center else length
        n_frames = min(D.shape[1], int(np.ceil(padded_length / hop_length)))
    else:
        n_frames = D.shape[1]
    expected = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected, dtype=dtype)
    blk = 256
    for s in range(0, n_frames, blk):
        ytmp = w * np.fft.irfft(D[:, s:min(s + blk, n_frames)], n=n_fft, axis=0)
        for f in range(ytmp.shape[1]):
            p = (s + f) * hop_length
            y[p:p + n_fft] += ytmp[:, f]
    wss = window_sumsquare(n_frames, hop_length, win_length, n_fft, dtype=dtype)
    nz = wss > np.finfo(dtype).tiny
    y[nz] /= wss[nz]
    if length is None:
        if center:
            y = y[n_fft // 2:-(n_fft // 2)]
    else:
        start = n_fft // 2 if center else 0
        y = y[start:start + length]
        if len(y) < length:
            y = np.pad(y, (0, length - len(y)))
    return y
