"""Drop-in for the hot-path methods of the reference's ``model/inference.py::AudioSynthesizer``.

Only the path members are reproduced: the Griffin-Lim inversion (inference.py:105-110) and the MIDI / audio
conditioning inputs (inference.py:37-57).  The PerformanceNet forward pass stays stock PyTorch and is out of scope.
"""
import os

import numpy as np
import torch

from . import _lib, features
from .preprocess import hyperparams, notes_to_pianoroll, process_spectrum_from_chunk
from .midi import read_midi_file

pp_hp = hyperparams()


class AudioSynthesizer():
    def __init__(self, checkpoint=None, exp_dir=None, midi_source=None, audio_source=None):
        # inference.py:23-29; the checkpoint is only read when somebody asks for it (model is out of scope)
        self.exp_dir = exp_dir
        self._checkpoint_name = checkpoint
        self.sample_rate = pp_hp.sr
        self.wps = pp_hp.wps
        self.midi_source = midi_source
        self.audio_source = audio_source

    @property
    def checkpoint(self):
        if getattr(self, "_checkpoint", None) is None:  # read once, like the attribute the reference sets in __init__
            self._checkpoint = torch.load(os.path.join(self.exp_dir, self._checkpoint_name))
        return self._checkpoint

    def process_custom_midi(self, midi_path):
        """inference.py:39-51: (pianoroll, onoff) transposed to (128, T)."""
        from . import pianoroll as _pr
        roll, onoff, _, _ = _pr.midi_to_pianoroll(read_midi_file(midi_path), self.wps)
        return (np.transpose(roll.to(torch.float64).cpu().numpy(), (1, 0)),
                np.transpose(onoff.to(torch.float64).cpu().numpy(), (1, 0)))

    def process_custom_midi_and_audio(self, midi_filename, audio_filename):
        """inference.py:37-71: the model's three inputs as CUDA float tensors with a leading batch axis --
        pianoroll (1,128,T), onoff (1,128,T), spec (1,1025,T') -- produced without leaving the device."""
        from . import audio_io, pianoroll as _pr
        midi_dir = os.path.join(self.exp_dir, 'midi') if self.exp_dir is not None else ''
        roll, onoff, _, _ = _pr.midi_to_pianoroll(read_midi_file(os.path.join(midi_dir, midi_filename)), self.wps)
        audio, _ = audio_io.load(audio_filename, sr=pp_hp.sr, as_numpy=False)
        spec = process_spectrum_from_chunk(audio)  # CUDA tensor in -> CUDA (1025, T') view out
        return (roll.t().to(torch.float32).unsqueeze(0), onoff.t().to(torch.float32).unsqueeze(0),
                spec.contiguous().unsqueeze(0))

    def process_custom_audio(self, audio):
        """inference.py:54-55: whole-song log1p-power spectrogram."""
        return process_spectrum_from_chunk(audio)

    def griffinlim(self, spectrogram, audio_id=None, n_iter=300, window='hann', n_fft=2048, hop_length=256,
                   verbose=False, random_state=None, init_phase=None):
        """inference.py:105-110.  ``spectrogram`` is the model's log1p-power output, (1025, T).

        magnitude = sqrt(expm1(clip(spectrogram, 0, 20))) is fused into the kernel that ingests the spectrogram;
        then librosa.griffinlim(magnitude, n_iter, window='hann', win_length=n_fft, hop_length) semantics.
        """
        was_np = not isinstance(spectrogram, torch.Tensor)
        device = _lib.require_cuda(None if was_np else spectrogram.device)
        S = torch.from_numpy(np.ascontiguousarray(spectrogram, dtype=np.float32)).to(device) if was_np \
            else spectrogram.to(torch.float32)
        if S.dim() != 2 or S.shape[0] != features.N_BINS:
            raise ValueError(f"spectrogram must be (1025, T); got {tuple(S.shape)}")
        features._check_window(window, n_fft, n_fft)
        T = int(S.shape[1])
        if init_phase is None and isinstance(random_state, (int, np.integer)):
            init_phase = np.random.RandomState(int(random_state)).rand(features.N_BINS, T)
        ph = None
        if init_phase is not None:
            ph = init_phase if isinstance(init_phase, torch.Tensor) else \
                torch.from_numpy(np.ascontiguousarray(init_phase, dtype=np.float32)).to(device)
            ph = ph.to(torch.float32).contiguous()
        seed = int(np.random.randint(0, 2 ** 31 - 1))  # librosa draws from NumPy's global generator too
        with features.ClipBatch.from_frames([T], int(hop_length), device=device) as b:
            y = features.griffinlim_batch(S.contiguous(), b, n_iter=n_iter, momentum=0.99, init_phase=ph, init="random",
                                          seed=seed, layout=features.BIN_MAJOR, is_log1p_power=True)
        return features.to_numpy(y) if was_np else y
