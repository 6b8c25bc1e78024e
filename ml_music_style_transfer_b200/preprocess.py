"""Drop-in for the hot-path functions of the reference's ``preprocessing/preprocess.py``.

Same names, argument meaning and return conventions; the arithmetic runs in the sm_100a kernels
behind ``torch.ops.mst_b200``.  What is *not* here is the reference's file I/O orchestration
(``get_data``, HDF5 writer, CLI: SURVEY section 8, out of scope).
"""
import glob

import numpy as np
import torch

from . import _lib, audio_io, features, pianoroll as _pr
from .midi import read_midi, read_midi_file, read_midi_notes


class hyperparams(object):
    """preprocess.py:17-44."""

    def __init__(self):
        self.sr = 44100  # Sampling rate (samples per second)
        self.n_fft = 2048  # fft points (samples)
        self.stride = 512  # number of windows of separation between chunks/data points
        self.piano_scores = {
            'train': [2240, 2530, 1763, 2308, 2533, 1772, 2444, 2478, 2509, 1776, 1749, 2486, 2487, 2678, 2490, 2492,
                      2527],
            'test': [2533, 1760],
        }
        self.styles = ['cuba', 'aliciakeys', 'gentleman', 'harpsichord', 'upright']
        self.ws = 256  # window size (audio samples per window) == hop
        self.wps = 44100 // self.ws  # ~172 windows/second
        self.spc = 5  # seconds per chunk


hp = hyperparams()
VERBOSE = False


def process_spectrum_from_chunk(audio_chunk):
    """preprocess.py:47-57: log1p(|librosa.stft(chunk, n_fft=2048, hop_length=256)|^2) -> float32 (1025, T).

    The result is Fortran-ordered exactly like the reference's (librosa allocates its STFT matrix order='F').
    """
    return features.spectrogram(audio_chunk, hp.ws, out="log1p_power")


def process_audio_into_chunks(audio, style, song_id, num_chunks, debug=False):
    """preprocess.py:60-77: (num_chunks, 1025, 860) float32, C-contiguous like ``np.array(spec_list)``.

    All chunks (overlapping windows of 219 904 samples every 131 072) go through ONE batched launch; each chunk is
    reflect-padded independently, as the per-chunk librosa.stft calls of the reference do.
    """
    if VERBOSE:
        print(f"processing {style} style for song_id {song_id}")
    n_samples_per_chunk = (hp.spc * hp.wps - 1) * hp.ws
    step = hp.ws * hp.stride
    a, was_np = features._to_device_audio(audio)
    if num_chunks <= 0:
        out = torch.empty((0,), dtype=torch.float32, device=a.device)
        return out.cpu().numpy() if was_np else out
    last_end = (num_chunks - 1) * step + n_samples_per_chunk
    if last_end > a.numel():
        # the reference silently produces ragged chunks here and np.array() then fails / makes an object array
        raise ValueError(f"audio has {a.numel()} samples but {num_chunks} chunks need {last_end}")
    with features.ClipBatch.uniform(num_chunks, n_samples_per_chunk, hp.ws, clip_stride=step, device=a.device) as b:
        T = b.total_frames // num_chunks
        out = features.stft_batch(a, b, "log1p_power", features.BIN_MAJOR).view(num_chunks, features.N_BINS, T)
    return features.to_numpy(out) if was_np else out


def process_pianoroll_into_chunks(pianoroll, onoff, song_id, num_chunks, debug=False):
    """preprocess.py:80-96: two (num_chunks, 860, 128) arrays (float64 for NumPy inputs, like the reference)."""
    if VERBOSE:
        print(f"processing pianoroll for song_id {song_id}")
    n_windows_per_chunk = hp.spc * hp.wps
    was_np = not isinstance(pianoroll, torch.Tensor)
    device = _lib.require_cuda(None if was_np else pianoroll.device)

    def dev(x):
        if isinstance(x, torch.Tensor):
            return x.to(torch.int8).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x).astype(np.int8)).to(device)

    out_dtype = torch.float64 if was_np else torch.int8
    score = _pr.chunks(dev(pianoroll), num_chunks, n_windows_per_chunk, hp.stride, out_dtype)
    oo = _pr.chunks(dev(onoff), num_chunks, n_windows_per_chunk, hp.stride, out_dtype)
    if was_np:
        return features.to_numpy(score), features.to_numpy(oo)
    return score, oo


def load_audio(data_dir, song_id, style, debug=False):
    """preprocess.py:99-115: glob the style's wav, librosa.load(path, sr=hp.sr) (decode, mono, kaiser_best resample)."""
    audio_file = glob.glob(f"{data_dir}/{song_id}*{style}.wav")
    if len(audio_file) == 0:
        raise ValueError("couldnt find audio track!")
    elif len(audio_file) > 1:
        raise ValueError("multiple files picked up, issue:", audio_file)
    y, sr = audio_io.load(audio_file[0], sr=hp.sr)
    if debug is True:
        print("length of audio clip / sr: ", len(y), sr)
        print("audio files picked up:", audio_file)
    return y


def get_num_song_chunks(pianoroll, offset_percentage=0.1, max_chunks=100):
    """preprocess.py:118-136 (host integer glue)."""
    n_windows_per_chunk = hp.spc * hp.wps
    num_chunks = (pianoroll.shape[0] - n_windows_per_chunk) // hp.stride
    offset = int(offset_percentage * num_chunks)
    num_chunks -= offset
    if num_chunks > max_chunks:
        if VERBOSE:
            print(f"song has more than max_chunks={max_chunks}, reducing")
        num_chunks = max_chunks
    if VERBOSE:
        print('song has {} chunks'.format(num_chunks))
    return num_chunks


def notes_to_pianoroll(pitch, velocity, start, end, fs=None, as_numpy=True, cc64=None, end_time=None, pedal_threshold=64):
    """preprocess.py:147-155 for an explicit note list (+ optional CC64 sustain events): (pianoroll, onoff), both (T,128)."""
    nb = _pr.NoteBatch(pitch, velocity, start, end, [0, len(pitch)], end_times=None if end_time is None else [end_time],
                       pedals=None if cc64 is None else [cc64])
    roll, onoff, _, _ = _pr.rasterize(nb, hp.wps if fs is None else fs, pedal_threshold=pedal_threshold)
    if as_numpy:
        return roll.to(torch.float64).cpu().numpy(), onoff.to(torch.float64).cpu().numpy()
    return roll, onoff


def load_midi(data_dir, song_id, ext='mixcraft', debug=False):
    """preprocess.py:139-160: glob the MIDI file, rasterise at hp.wps, binarise, on/off.  Returns float64 (T,128) x2."""
    midi_file = glob.glob(f"{data_dir}/{song_id}*{ext}.mid")
    if len(midi_file) == 0:
        raise ValueError("couldnt find midi track!")
    elif len(midi_file) > 1:
        raise ValueError("multiple files picked up, issue:", midi_file)
    # PrettyMIDI(file).get_piano_roll(fs=hp.wps).T with every instrument's own pedal / bends, then binarise + on/off
    roll, oo, _, _ = _pr.midi_to_pianoroll(read_midi_file(midi_file[0]), hp.wps)
    pianoroll, onoff = roll.to(torch.float64).cpu().numpy(), oo.to(torch.float64).cpu().numpy()
    if debug is True:
        print("length of pianoroll: ", pianoroll.shape)
        print("midi files picked up:", midi_file)
    return pianoroll, onoff


def get_data(data_dir, dataset_outpath, data_type, debug=False, piano_scores=None, styles=None):
    """preprocess.py:163-200 with the HDF5 file replaced by a shard directory ``{dataset_outpath}_{data_type}``
    (dataset.ShardManager mirrors io_manager.h5pyManager).  ``piano_scores`` / ``styles`` default to ``hp``'s."""
    from .dataset import ShardManager
    data_manager = ShardManager(f"{dataset_outpath}_{data_type}", mode="w")   # h5py.File(outfile, 'w') truncates too
    for song_id in (hp.piano_scores[data_type] if piano_scores is None else piano_scores):
        pianoroll, onoff = load_midi(data_dir, song_id, debug=debug)
        num_chunks = get_num_song_chunks(pianoroll)
        pianoroll_list, onoff_list = process_pianoroll_into_chunks(pianoroll, onoff, song_id, num_chunks, debug=debug)
        data_manager.write_pianoroll(pianoroll_list, onoff_list)
        for style in (hp.styles if styles is None else styles):
            try:
                audio = load_audio(data_dir, song_id, style, debug=debug)
            except ValueError:
                # not all styles exist for all midi...  (the reference uses a bare except here, preprocess.py:185-190)
                print(f"Couldnt load audio for song={song_id}, style={style}, skipping...")
                continue
            spec_list = process_audio_into_chunks(audio, style, song_id, num_chunks, debug=debug)
            data_manager.write_spectrum(spec_list, style)
            if debug is True:
                assert pianoroll_list.shape[0] == spec_list.shape[0]
                assert pianoroll_list.shape == onoff_list.shape
    return data_manager
