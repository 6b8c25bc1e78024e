"""librosa-0.8 ``filters.mel`` / ``feature.melspectrogram`` restated (oracle; test infrastructure).

Reference call site: tests/plot_spec.py:20 ``librosa.feature.melspectrogram(y=, sr=hp.sr,
n_fft=hp.n_fft, hop_length=hp.ws)`` (live) and preprocessing/preprocess.py:55 (commented).
Defaults followed: power=2.0, n_mels=128, fmin=0, fmax=sr/2, htk=False (Slaney scale),
norm='slaney' (area normalisation), filterbank dtype float32.  The reference applies no log to
the mel output; ``logmel`` below is this build's ``log1p`` convention (preprocess.py:49 uses
log1p on the linear power spectrogram).
"""
import numpy as np
from . import stft as _stft

__all__ = ["hz_to_mel", "mel_to_hz", "mel_frequencies", "mel_filterbank", "melspectrogram", "logmel"]

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    mels = f / _F_SP
    log_t = f >= _MIN_LOG_HZ
    safe = np.where(log_t, f, _MIN_LOG_HZ)
    return np.where(log_t, _MIN_LOG_MEL + np.log(safe / _MIN_LOG_HZ) / _LOGSTEP, mels)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    freqs = _F_SP * m
    log_t = m >= _MIN_LOG_MEL
    return np.where(log_t, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), freqs)


def mel_frequencies(n_mels, fmin, fmax):
    return mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels))


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, dtype=np.float32):
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=dtype)
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2, endpoint=True)
    mel_f = mel_frequencies(n_mels + 2, fmin, fmax)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def melspectrogram(y, sr, n_fft=2048, hop_length=512, n_mels=128, pad_mode="reflect"):
    """(n_mels, T) float32: mel_basis @ |stft|**2, contraction evaluated in float64."""
    S = _stft.stft(y, n_fft=n_fft, hop_length=hop_length, pad_mode=pad_mode)
    P = (np.abs(S) ** 2).astype(np.float32)
    W = mel_filterbank(sr, n_fft, n_mels)
    return (W.astype(np.float64) @ P.astype(np.float64)).astype(np.float32)


def logmel(y, sr, n_fft=2048, hop_length=512, n_mels=128, pad_mode="reflect"):
    return np.log1p(melspectrogram(y, sr, n_fft, hop_length, n_mels, pad_mode))
