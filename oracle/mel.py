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

__all__ = ["hz_to_mel", "mel_to_hz", "mel_frequencies", "mel_filterbank", "melspectrogram", "logmel", "nnls", "mel_to_stft",
           "mel_to_audio"]

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


# ---- mel inversion (librosa.feature.inverse.mel_to_stft / mel_to_audio; tests/test_griffinlim.py:24, commented) -------
_MAX_MEM_BLOCK = 2 ** 8 * 2 ** 10   # librosa.util.utils.MAX_MEM_BLOCK


def _nnls_obj(x, shape, A, B):
    """librosa.util._nnls._nnls_obj: 0.5/B.size * ||A x - B||^2 and its gradient (float32 einsum like the original)."""
    x = x.reshape(shape)
    diff = np.einsum("mf,...ft->...mt", A, x, optimize=True) - B
    value = (1 / B.size) * 0.5 * np.sum(diff ** 2)
    grad = (1 / B.size) * np.einsum("mf,...mt->...ft", A, diff, optimize=True)
    return value, grad.flatten()


def _nnls_lbfgs_block(A, B, x_init=None, **kwargs):
    import scipy.optimize
    if x_init is None:
        x_init = np.linalg.lstsq(A, B, rcond=None)[0]
        np.clip(x_init, 0, None, out=x_init)
    kwargs.setdefault("m", A.shape[1])
    bounds = [(0, None)] * x_init.size
    shape = x_init.shape
    x, obj_value, diagnostics = scipy.optimize.fmin_l_bfgs_b(_nnls_obj, x_init, args=(shape, A, B), bounds=bounds, **kwargs)
    return x.reshape(shape)


def nnls(A, B, **kwargs):
    """librosa.util.nnls (0.8): non-negative least squares min ||A X - B||, X >= 0, column blocks solved by L-BFGS-B
    started from the clipped least-squares solution."""
    import scipy.optimize
    if B.ndim == 1:
        return scipy.optimize.nnls(A, B)[0]
    n_columns = int(_MAX_MEM_BLOCK // (A.shape[-1] * A.itemsize))
    if B.shape[-1] <= n_columns:
        return _nnls_lbfgs_block(A, B, **kwargs).astype(A.dtype)
    x = np.linalg.lstsq(A, B, rcond=None)[0].astype(A.dtype)
    np.clip(x, 0, None, out=x)
    x_init = x
    for bl_s in range(0, x.shape[-1], n_columns):
        bl_t = min(bl_s + n_columns, B.shape[-1])
        x[:, bl_s:bl_t] = _nnls_lbfgs_block(A, B[:, bl_s:bl_t], x_init=x_init[:, bl_s:bl_t], **kwargs)
    return x


def mel_to_stft(M, sr=22050, n_fft=2048, power=2.0, **kwargs):
    """librosa.feature.inverse.mel_to_stft: nnls(mel_basis, M) ** (1 / power), (1 + n_fft/2, T)."""
    M = np.asarray(M)
    mel_basis = mel_filterbank(sr, n_fft, n_mels=M.shape[0], dtype=M.dtype if M.dtype.kind == "f" else np.float32, **kwargs)
    inverse = nnls(mel_basis, M)
    return np.power(inverse, 1.0 / power, out=inverse)


def mel_to_audio(M, sr=22050, n_fft=2048, hop_length=512, power=2.0, n_iter=32, init_phase=None, momentum=0.99):
    """librosa.feature.inverse.mel_to_audio = griffinlim(mel_to_stft(M)).  ``init_phase``: see oracle.griffinlim."""
    from . import griffinlim as _gl
    S = mel_to_stft(M, sr=sr, n_fft=n_fft, power=power)
    return _gl.griffinlim(S, n_iter, hop_length, momentum=momentum, init_phase=init_phase)
