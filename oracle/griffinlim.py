"""librosa-0.8 ``griffinlim`` restated in NumPy (oracle; test infrastructure).

Reference call sites: model/inference.py:105-110 (``sqrt(expm1(clip(spec,0,20)))`` then
``librosa.griffinlim(mag, n_iter=300, window='hann', win_length=2048, hop_length=256)``) and
tests/test_griffinlim.py:23.  ``momentum=0`` reproduces the classic loop kept as a comment at
model/inference.py:131-154, whose Frobenius loss (:149-150) is the convergence metric used for
parity (``spectral_convergence``).

Upstream recurrence followed (librosa 0.8.0 core/spectrum.py): complex64 ``angles`` initialised
to exp(2*pi*i*rand) (``init='random'``) or 1 (``init=None``); ``rebuilt = 0``; per iteration
``tprev = rebuilt; inverse = istft(S*angles); rebuilt = stft(inverse);
angles = rebuilt - momentum/(1+momentum)*tprev; angles /= |angles| + 1e-16``; returns
``istft(S*angles)``.  librosa draws the phase from an unseeded global RNG; here the uniform
[0,1) field is an explicit argument so that both sides of a parity test start identically.
"""
import numpy as np
from . import stft as _stft

__all__ = ["logpower_to_magnitude", "griffinlim", "spectral_convergence", "random_phase"]


def logpower_to_magnitude(spec):
    """model/inference.py:109."""
    return np.sqrt(np.expm1(np.clip(spec, 0, 20)))


def random_phase(shape, seed):
    """Uniform [0,1) field as librosa draws it with ``random_state=seed`` (RandomState.rand)."""
    return np.random.RandomState(seed).rand(*shape)


def griffinlim(S, n_iter=32, hop_length=None, win_length=None, momentum=0.99, init_phase=None,
               pad_mode="reflect", return_history=False):
    """S: (1+n_fft/2, T) magnitudes.  init_phase: uniform [0,1) array of S.shape, or None for angles=1."""
    S = np.asarray(S, dtype=np.float32)
    n_fft = 2 * (S.shape[0] - 1)
    angles = np.empty(S.shape, dtype=np.complex64)
    if init_phase is None:
        angles[:] = 1.0
    else:
        angles[:] = np.exp(2j * np.pi * np.asarray(init_phase))
    rebuilt = 0.0
    hist = []
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = _stft.istft(S * angles, hop_length=hop_length, win_length=win_length)
        rebuilt = _stft.stft(inverse, n_fft=n_fft, hop_length=hop_length, win_length=win_length, pad_mode=pad_mode)
        angles[:] = rebuilt - (momentum / (1 + momentum)) * tprev
        angles[:] /= np.abs(angles) + 1e-16
        if return_history:
            hist.append(float(np.linalg.norm(np.abs(rebuilt) - S) / np.linalg.norm(S)))
    y = _stft.istft(S * angles, hop_length=hop_length, win_length=win_length)
    return (y, hist) if return_history else y


def spectral_convergence(S, y, hop_length, n_fft=None, win_length=None, pad_mode="reflect"):
    """|| |STFT(y)| - S ||_F / || S ||_F  (normalised form of model/inference.py:149-150)."""
    S = np.asarray(S, dtype=np.float64)
    if n_fft is None:
        n_fft = 2 * (S.shape[0] - 1)
    R = np.abs(_stft.stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                          pad_mode=pad_mode, out_dtype=np.complex128))
    return float(np.linalg.norm(R - S) / np.linalg.norm(S))
