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

    def __init__(self, pitch, velocity, start, end, note_offsets, device=None, end_times=None, pedals=None):
        device = _lib.require_cuda(device)
        self.device = device
        self.pitch = torch.as_tensor(np.asarray(pitch, dtype=np.int32)).to(device)
        self.velocity = torch.as_tensor(np.asarray(velocity, dtype=np.int32)).to(device)
        self.start = torch.as_tensor(np.asarray(start, dtype=np.float64)).to(device)
        self.end = torch.as_tensor(np.asarray(end, dtype=np.float64)).to(device)
        h_off = np.asarray(note_offsets, dtype=np.int64)
        h_end = np.asarray(end, dtype=np.float64)
        self.note_offsets = torch.as_tensor(h_off).to(device)
        self.n_pieces = int(self.note_offsets.numel()) - 1
        # host copies of what the roll geometry needs (latest note end per piece), so that rasterize() never has to
        # read a size back from the device
        self.h_max_end = np.array([h_end[a:b].max() if b > a else 0.0 for a, b in zip(h_off[:-1], h_off[1:])], dtype=np.float64)
        if self.n_pieces < 1:
            raise ValueError("note_offsets needs at least two entries")
        # optional per-piece extras of a parsed MIDI file: latest event time (pretty_midi get_end_time also counts
        # control changes) and the CC64 event list [(time, value), ...] for the sustain rule
        self.end_times = None if end_times is None else [float(t) for t in end_times]
        self.pedals = pedals

    @classmethod
    def from_host_tensors(cls, pitch, velocity, start, end, note_offsets, h_max_end, device=None):
        """Asynchronous construction from (pinned) host tensors on the current stream; ``h_max_end`` is the latest note
        end of every piece (host float64 array), which spares rasterize() a device read-back."""
        self = cls.__new__(cls)
        self.device = _lib.require_cuda(device)
        self.pitch, self.velocity, self.start, self.end = [t.to(self.device, non_blocking=True)
                                                           for t in (pitch, velocity, start, end)]
        self.note_offsets = note_offsets.to(self.device, non_blocking=True)
        self.n_pieces = int(note_offsets.numel()) - 1
        self.end_times, self.pedals = None, None
        self.h_max_end = np.asarray(h_max_end, dtype=np.float64)
        return self

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


def pedal_spans(cc64_events, fs, pedal_threshold=64):
    """pretty_midi's sustain state machine (get_piano_roll, pedal_threshold): [(time, value), ...] in file order ->
    list of (start_col, end_col) pedal-down spans; a pedal still down at the end of the piece sustains nothing."""
    spans, on, t_on = [], False, 0
    for t, v in cc64_events:
        now = int(float(t) * fs)
        cur = v >= pedal_threshold
        if not on and cur:
            t_on, on = now, True
        elif on and not cur:
            spans.append((t_on, now))
            on = False
    return spans


def rasterize(notes, fs, want_velsum=False, pedal_threshold=64):
    """-> (roll uint8 [sum T,128] in {0,1}, onoff int8 [sum T,128] in {-1,0,1}, row_offsets int64 [n_pieces+1], velsum|None).

    T_p = int(fs * end_time) per piece (pretty_midi roll width; end_time = latest note end, or ``notes.end_times``);
    rows are time-major (the reference's ``.T``).  When ``notes.pedals`` holds CC64 events the sustain rule of
    pretty_midi >= 0.2.9 (``pedal_threshold``, None = off) is applied before binarising.
    """
    o = _lib.ops()
    if int(fs) != fs or fs <= 0:
        # the kernels take an integer rate (the reference's hp.wps = sr // ws is one); a fractional fs would size the rows
        # on the host with one value and place the columns on the device with another
        raise ValueError(f"fs={fs!r} must be a positive integer")
    fs = int(fs)
    h_end = notes.end_times if notes.end_times is not None else getattr(notes, "h_max_end", None)
    if h_end is not None:
        # int(fs * end_time) on the host is the same IEEE double product the device kernel evaluates
        h_rows = np.array([max(0, int(fs * float(t))) for t in h_end], dtype=np.int64)
        h_ro = np.zeros(notes.n_pieces + 1, dtype=np.int64)
        np.cumsum(h_rows, out=h_ro[1:])
        row_offsets = torch.from_numpy(h_ro).pin_memory().to(notes.device, non_blocking=True)
        total_rows = int(h_ro[-1])
    else:  # device-only note arrays: count on the device, one small read-back to size the roll
        rows = o.pianoroll_count_rows(notes.end, notes.note_offsets, int(fs))
        row_offsets = torch.zeros(notes.n_pieces + 1, dtype=torch.int64, device=notes.device)
        torch.cumsum(rows, 0, out=row_offsets[1:])
        total_rows = int(row_offsets[-1].item())
    sp = ss = se = None
    if notes.pedals is not None and pedal_threshold is not None:
        spans = [(i, a, b) for i, ev in enumerate(notes.pedals) for a, b in pedal_spans(ev, fs, pedal_threshold) if b > a]
        if spans:
            sp = torch.tensor([x[0] for x in spans], dtype=torch.int32, device=notes.device)
            ss = torch.tensor([x[1] for x in spans], dtype=torch.int64, device=notes.device)
            se = torch.tensor([x[2] for x in spans], dtype=torch.int64, device=notes.device)
    roll, onoff, velsum = o.pianoroll_rasterize(notes.pitch, notes.velocity, notes.start, notes.end, notes.note_offsets,
                                                row_offsets, total_rows, int(fs), bool(want_velsum), sp, ss, se)
    return roll, onoff, row_offsets, (velsum if (want_velsum or sp is not None) else None)


def get_piano_roll(pitch, velocity, start, end, fs=100, cc64=None, end_time=None, pedal_threshold=64):
    """pretty_midi Instrument.get_piano_roll(fs, pedal_threshold) for one note list (+ optional CC64 events):
    float64 (128, T) velocity sums (NumPy)."""
    nb = NoteBatch(pitch, velocity, start, end, [0, len(pitch)], end_times=None if end_time is None else [end_time],
                   pedals=None if cc64 is None else [cc64])
    _, _, _, velsum = rasterize(nb, fs, want_velsum=True, pedal_threshold=pedal_threshold)
    return velsum.t().to(torch.float64).cpu().numpy()


def bend_segments(pitch_bends, end_time, fs):
    """pretty_midi's bend loop (Instrument.get_piano_roll) reduced to its active segments: for consecutive bends
    (sorted by time, a zero bend appended at ``end_time``) with |pitch| >= 1 and a non-empty column range ->
    (c0, c1, d, 1 - d, bend_int, positive).  All scalars are evaluated here exactly as Python evaluates them."""
    ordered = sorted(pitch_bends, key=lambda b: b[1])
    segs = []
    for (pitch, t0), (_, t1) in zip(ordered, ordered[1:] + [(0, end_time)]):
        if abs(pitch) < 1:
            continue
        c0, c1 = int(t0 * fs), int(t1 * fs)
        if c1 <= c0:
            continue
        semis = 2.0 * pitch / 8192.0
        bend_int = int(np.sign(semis) * np.floor(np.abs(semis)))
        d = float(np.abs(semis - bend_int))
        segs.append((c0, c1, d, 1 - d, bend_int, 1 if pitch >= 0 else 0))
    return segs


_SEG_DTYPE = np.dtype([("c0", "<i8"), ("c1", "<i8"), ("d", "<f8"), ("m1", "<f8"), ("bi", "<i4"), ("pos", "<i4")])


def midi_to_pianoroll(midi_files, fs, pedal_threshold=64, want_f64=False, device=None):
    """``PrettyMIDI.get_piano_roll(fs).T`` -> binarise -> on/off for parsed MIDI files (``midi.read_midi_file``), with
    pretty_midi's full semantics: one roll per instrument (its own CC64 sustain spans and pitch bends, drums zero),
    summed in instrument order into the widest roll (preprocess.py:146-155, inference.py:40-51).

    Returns (roll uint8 [sum T, 128], onoff int8, row_offsets int64 [n_files + 1], velsum float64 | None).
    """
    device = _lib.require_cuda(device)
    if not isinstance(midi_files, (list, tuple)):
        midi_files = [midi_files]
    insts, file_inst_off = [], [0]
    for mf in midi_files:
        insts += list(mf.instruments)
        file_inst_off.append(len(insts))
    n_files = len(midi_files)
    file_rows = []
    for f in range(n_files):
        mine = insts[file_inst_off[f]:file_inst_off[f + 1]]
        file_rows.append(max([int(fs * i.get_end_time()) if i.n_notes else 0 for i in mine], default=0))
    file_row_off = np.zeros(n_files + 1, dtype=np.int64)
    np.cumsum(file_rows, out=file_row_off[1:])
    total_rows = int(file_row_off[-1])
    if not insts or total_rows == 0:
        z = torch.zeros((total_rows, 128), dtype=torch.uint8, device=device)
        return z, z.to(torch.int8), torch.from_numpy(file_row_off).to(device), (z.to(torch.float64) if want_f64 else None)
    # one "piece" per instrument through the note rasteriser + CC64 rule (drums keep their width but get no notes)
    empty = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64), np.zeros(0, np.float64))
    nb = NoteBatch.from_pieces([empty if i.is_drum else i.arrays() for i in insts], device=device)
    nb.end_times = [i.get_end_time() if i.n_notes else 0.0 for i in insts]
    nb.pedals = [[] if i.is_drum else i.cc64() for i in insts]
    _, _, inst_row_off, velsum = rasterize(nb, fs, want_velsum=True, pedal_threshold=pedal_threshold)
    seg_off, segs = [0], []
    for i in insts:
        if not i.is_drum and i.n_notes:
            segs += bend_segments(i.pitch_bends, i.get_end_time(), fs)
        seg_off.append(len(segs))
    seg_np = np.array(segs, dtype=_SEG_DTYPE) if segs else np.zeros(0, dtype=_SEG_DTYPE)
    dev_i32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int32)).to(device)
    seg_bytes = torch.from_numpy(np.frombuffer(seg_np.tobytes() or b"\0" * _SEG_DTYPE.itemsize, dtype=np.uint8).copy()).to(device)
    roll, onoff, out = _lib.ops().pianoroll_merge_instruments(
        velsum.contiguous(), inst_row_off, dev_i32([1 if i.is_drum else 0 for i in insts]), dev_i32(file_inst_off),
        torch.from_numpy(file_row_off).to(device), total_rows, dev_i32(seg_off), seg_bytes, bool(want_f64))
    return roll, onoff, torch.from_numpy(file_row_off).to(device), (out if want_f64 else None)


def chunks(plane, num_chunks, chunk_rows, stride_rows, dtype=torch.float64):
    """preprocess.py:80-96 on the device: (num_chunks, chunk_rows, 128)."""
    return _lib.ops().pianoroll_chunks(plane, int(num_chunks), int(chunk_rows), int(stride_rows), DTYPE_CODES[dtype])


def _sample_offsets(n_pieces, samples_per_piece, device):
    if np.ndim(samples_per_piece) == 0:  # uniform pieces: offsets built on the device, nothing crosses PCIe
        so = np.arange(n_pieces + 1, dtype=np.int64) * int(samples_per_piece)
        return so, torch.arange(n_pieces + 1, dtype=torch.int64, device=device) * int(samples_per_piece)
    spp = np.asarray(samples_per_piece, dtype=np.int64)
    so = np.zeros(n_pieces + 1, dtype=np.int64)
    np.cumsum(spp, out=so[1:])
    return so, torch.from_numpy(so).pin_memory().to(device, non_blocking=True)


def upsample_pair(roll, onoff, row_offsets, samples_per_piece, fs, sr, pitch_lo=21, n_keys=88, dtype=torch.int8):
    """Roll AND on/off to the audio rate in ONE launch (shared index arithmetic): -> (up_roll, up_onoff, sample_offsets),
    each laid out like ``upsample``'s result."""
    n_pieces = int(row_offsets.numel()) - 1
    so, sample_offsets = _sample_offsets(n_pieces, samples_per_piece, roll.device)
    a, b = _lib.ops().pianoroll_upsample_pair(roll, onoff, row_offsets, sample_offsets, int(so[-1]), int(fs), int(sr),
                                              int(pitch_lo), int(n_keys), DTYPE_CODES[dtype])
    return a, b, so


def upsample(plane, row_offsets, samples_per_piece, fs, sr, pitch_lo=21, n_keys=88, dtype=torch.int8):
    """Hold-replicate a frame-rate plane to the audio rate: per piece (n_keys, N_p), col(n) = (n*fs)//sr.

    Returns (flat tensor, sample_offsets); piece p's block is flat[n_keys*off[p] : n_keys*off[p+1]].view(n_keys, N_p).
    """
    n_pieces = int(row_offsets.numel()) - 1
    so, sample_offsets = _sample_offsets(n_pieces, samples_per_piece, plane.device)
    out = _lib.ops().pianoroll_upsample(plane, row_offsets, sample_offsets, int(so[-1]), int(fs), int(sr), int(pitch_lo),
                                        int(n_keys), DTYPE_CODES[dtype])
    return out, so
