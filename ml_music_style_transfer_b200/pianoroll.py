"""MIDI notes -> piano roll / on-off / chunks / audio-rate planes on the GPU (P3).

Mirrors ``pretty_midi.PrettyMIDI.get_piano_roll(fs)`` (notes only) and the NumPy lines of the
reference (preprocessing/preprocess.py:146-155, :80-96; model/inference.py:40-51).
"""
import numpy as np
import torch

from . import _lib

DTYPE_CODES = {torch.int8: 0, torch.float32: 1, torch.float64: 2}


class NoteBatch:
    """Structure-of-arrays notes for a batch of pieces, resident on the device.

    pitch / velocity: int32; start / end: float64 seconds; note_offsets: int64 [n_pieces + 1].
    """

    def __init__(self, pitch, velocity, start, end, note_offsets, device=None):
        device = _lib.require_cuda(device)
        self.device = device
        self.pitch = torch.as_tensor(np.asarray(pitch, dtype=np.int32)).to(device)
        self.velocity = torch.as_tensor(np.asarray(velocity, dtype=np.int32)).to(device)
        self.start = torch.as_tensor(np.asarray(start, dtype=np.float64)).to(device)
        self.end = torch.as_tensor(np.asarray(end, dtype=np.float64)).to(device)
        self.note_offsets = torch.as_tensor(np.asarray(note_offsets, dtype=np.int64)).to(device)
        self.n_pieces = int(self.note_offsets.numel()) - 1
        if self.n_pieces < 1:
            raise ValueError("note_offsets needs at least two entries")

    @classmethod
    def from_pieces(cls, pieces, device=None):
        """pieces: iterable of (pitch, velocity, start, end) array tuples."""
        ps, vs, ss, es, offs = [], [], [], [], [0]
        for p, v, s, e in pieces:
            ps.append(np.asarray(p, dtype=np.int32)); vs.append(np.asarray(v, dtype=np.int32))
            ss.append(np.asarray(s, dtype=np.float64)); es.append(np.asarray(e, dtype=np.float64))
            offs.append(offs[-1] + len(ps[-1]))
        cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
        return cls(cat(ps, np.int32), cat(vs, np.int32), cat(ss, np.float64), cat(es, np.float64), offs, device)


def rasterize(notes, fs, want_velsum=False):
    """-> (roll uint8 [sum T,128] in {0,1}, onoff int8 [sum T,128] in {-1,0,1}, row_offsets int64 [n_pieces+1], velsum|None).

    T_p = int(fs * max note end) per piece (pretty_midi roll width); rows are time-major (the reference's ``.T``).
    """
    o = _lib.ops()
    rows = o.pianoroll_count_rows(notes.end, notes.note_offsets, int(fs))
    row_offsets = torch.zeros(notes.n_pieces + 1, dtype=torch.int64, device=notes.device)
    torch.cumsum(rows, 0, out=row_offsets[1:])
    total_rows = int(row_offsets[-1].item())  # one small D2H: the roll has to be allocated
    roll, onoff, velsum = o.pianoroll_rasterize(notes.pitch, notes.velocity, notes.start, notes.end, notes.note_offsets,
                                                row_offsets, total_rows, int(fs), bool(want_velsum))
    return roll, onoff, row_offsets, (velsum if want_velsum else None)


def get_piano_roll(pitch, velocity, start, end, fs=100):
    """pretty_midi Instrument.get_piano_roll(fs) for one note list: float64 (128, T) velocity sums (NumPy)."""
    nb = NoteBatch(pitch, velocity, start, end, [0, len(pitch)])
    _, _, _, velsum = rasterize(nb, fs, want_velsum=True)
    return velsum.t().to(torch.float64).cpu().numpy()


def chunks(plane, num_chunks, chunk_rows, stride_rows, dtype=torch.float64):
    """preprocess.py:80-96 on the device: (num_chunks, chunk_rows, 128)."""
    return _lib.ops().pianoroll_chunks(plane, int(num_chunks), int(chunk_rows), int(stride_rows), DTYPE_CODES[dtype])


def upsample(plane, row_offsets, samples_per_piece, fs, sr, pitch_lo=21, n_keys=88, dtype=torch.int8):
    """Hold-replicate a frame-rate plane to the audio rate: per piece (n_keys, N_p), col(n) = (n*fs)//sr.

    Returns (flat tensor, sample_offsets); piece p's block is flat[n_keys*off[p] : n_keys*off[p+1]].view(n_keys, N_p).
    """
    n_pieces = int(row_offsets.numel()) - 1
    spp = np.broadcast_to(np.asarray(samples_per_piece, dtype=np.int64), (n_pieces,))
    so = np.zeros(n_pieces + 1, dtype=np.int64)
    np.cumsum(spp, out=so[1:])
    sample_offsets = torch.from_numpy(so).to(plane.device)
    out = _lib.ops().pianoroll_upsample(plane, row_offsets, sample_offsets, int(so[-1]), int(fs), int(sr), int(pitch_lo),
                                        int(n_keys), DTYPE_CODES[dtype])
    return out, so
