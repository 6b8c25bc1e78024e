"""CPU oracle for the preprocessing / inversion hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy (float64-compute) restatement of the algorithms the reference
delegates to librosa 0.8 / pretty_midi 0.2.9 on the hot path named in BASELINE.json, plus the
few NumPy lines the reference itself contributes (preprocess.py:47-57,60-96,118-160,
model/inference.py:105-110).  It exists so that tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py have something to check the CUDA path
against.  Nothing in ``ml_music_style_transfer_b200/`` imports it, and nothing there falls back
to it: the product path raises when its CUDA extension is missing.

PARITY UNPINNED BY THE REFERENCE: /root/reference ships no golden vectors, no assertions
(tests/test_griffinlim.py:14-27 only writes a wav and cannot even be imported) and its
arithmetic lives in librosa / pretty_midi, neither of which is installed or installable here
(no requirements file pins a version either; era evidence says librosa 0.7.2-0.8.0,
pretty_midi 0.2.8-0.2.9).  The oracle is therefore pinned against *independent* implementations
that do exist in this image -- ``torch.stft`` / ``torch.istft``, ``scipy.signal.stft``,
``torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney')`` and
``torchaudio.functional.griffinlim`` -- and against closed-form known-answer vectors; the
generated fixtures live in tests/golden/ together with the script that made them
(tests/golden/make_golden.py).
"""
from . import stft, mel, pianoroll, griffinlim, preprocess, audio  # noqa: F401
