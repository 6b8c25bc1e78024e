"""The reference's own hot-path functions restated on top of the oracle primitives
(test infrastructure).  Same names / argument meaning as preprocessing/preprocess.py.
"""
import numpy as np
from . import stft as _stft
from . import pianoroll as _pr


class hyperparams(object):
    """preprocess.py:17-44 (numeric fields only)."""

    def __init__(self):
        self.sr = 44100
        self.n_fft = 2048
        self.stride = 512
        self.ws = 256
        self.wps = 44100 // self.ws
        self.spc = 5


hp = hyperparams()


def process_spectrum_from_chunk(audio_chunk):
    """preprocess.py:47-57: log1p(|stft|^2), (1025, T) float32."""
    spec = _stft.stft(audio_chunk, n_fft=hp.n_fft, hop_length=hp.ws)
    return np.log1p(np.abs(spec) ** 2)


def process_audio_into_chunks(audio, style, song_id, num_chunks, debug=False):
    """preprocess.py:60-77."""
    out = []
    n = (hp.spc * hp.wps - 1) * hp.ws
    for step in range(num_chunks):
        s = step * hp.ws * hp.stride
        out.append(process_spectrum_from_chunk(audio[s:s + n]))
    return np.array(out)


def process_pianoroll_into_chunks(pianoroll, onoff, song_id, num_chunks, debug=False):
    return _pr.process_pianoroll_into_chunks(pianoroll, onoff, num_chunks, hp.spc * hp.wps, hp.stride)


def get_num_song_chunks(pianoroll, offset_percentage=0.1, max_chunks=100):
    return _pr.get_num_song_chunks(pianoroll.shape[0], offset_percentage, max_chunks, hp.spc * hp.wps, hp.stride)


def midi_notes_to_pianoroll(pitch, velocity, start, end, fs=None, cc64=None, end_time=None):
    """preprocess.py:146-155 with the SMF parse replaced by explicit note arrays."""
    roll = _pr.get_piano_roll(pitch, velocity, start, end, hp.wps if fs is None else fs, end_time=end_time, cc64=cc64)
    return _pr.binarize_and_onoff(roll)
