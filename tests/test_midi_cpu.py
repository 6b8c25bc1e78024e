"""CPU suite: the SMF reader against hand-derived pretty_midi behaviour (note pairing, tempo map, instrument
membership), and hand-derived known answers for the oracle's restatement of PrettyMIDI.get_piano_roll (per-instrument
pedal, pitch bends, drums, instrument sum).  pretty_midi itself is absent from the image: every expected value below is
derived by hand from its 0.2.9 source semantics and spelled out in the test."""
import numpy as np

from ml_music_style_transfer_b200 import midi
from oracle import pianoroll as opr

ON, OFF = 0x90, 0x80


def test_zero_length_note_does_not_dangle(tmp_path):
    """pretty_midi: a note-off that closes nothing (the only open note-on has the same tick) FORGETS the key.  So
    on@0, off@0, on@480, off@960 is exactly one note [0.5 s, 1.0 s] -- not an extra phantom note from tick 0."""
    path = str(tmp_path / "z.mid")
    midi.write_midi_tracks(path, [[('raw', 0, ON, 60, 90), ('raw', 0, OFF, 60, 0), ('raw', 480, ON, 60, 70),
                                   ('raw', 960, OFF, 60, 0)]])
    p, v, s, e, cc, end_time = midi.read_midi(path)
    assert list(p) == [60] and list(v) == [70] and list(s) == [0.5] and list(e) == [1.0] and end_time == 1.0


def test_same_tick_note_on_survives_when_something_closed(tmp_path):
    """on@0, on@480, off@480: the off closes the first note; the note-on of tick 480 continues and is closed at 960."""
    path = str(tmp_path / "k.mid")
    midi.write_midi_tracks(path, [[('raw', 0, ON, 60, 90), ('raw', 480, ON, 60, 70), ('raw', 480, OFF, 60, 0),
                                   ('raw', 960, OFF, 60, 0), ('raw', 1200, OFF, 60, 0)]])
    p, v, s, e, _, _ = midi.read_midi(path)
    assert list(zip(p, v, s, e)) == [(60, 90, 0.0, 0.5), (60, 70, 0.5, 1.0)]   # the off at 1200 finds no open key


def test_tempo_map_uses_track_zero_only(tmp_path):
    """Tempo changes: 120 bpm, then 60 bpm from tick 480 (track 0).  A set_tempo hidden in track 1 is ignored."""
    path = str(tmp_path / "t.mid")
    midi.write_midi_tracks(path, [[('raw', 0, ON, 60, 90), ('raw', 960, OFF, 60, 0)],
                                  [('raw', 0, ON | 1, 64, 80), ('raw', 480, OFF | 1, 64, 0)]],
                           tempo_changes=[(480, 60.0)])
    with open(path, 'rb') as f:
        raw = bytearray(f.read())
    # splice a set_tempo (240 bpm) at tick 0 into track 1: find the second MTrk and insert after its header
    i = raw.index(b'MTrk', raw.index(b'MTrk') + 4)
    ev = b'\x00\xFF\x51\x03' + int(6e7 / 240).to_bytes(3, 'big')
    ln = int.from_bytes(raw[i + 4:i + 8], 'big') + len(ev)
    raw[i + 4:i + 8] = ln.to_bytes(4, 'big')
    raw[i + 8:i + 8] = ev
    with open(path, 'wb') as f:
        f.write(bytes(raw))
    mf = midi.read_midi_file(path)
    assert len(mf.instruments) == 2
    a, b = mf.instruments
    # 480 ticks at 120 bpm = 0.5 s, then 480 ticks at 60 bpm = 1.0 s
    assert a.start == [0.0] and a.end == [1.5]
    assert b.start == [0.0] and b.end == [0.5]


def test_instrument_membership_and_event_attachment(tmp_path):
    """(program, channel, track) instruments; CCs before the first note travel via the straggler; a channel that never
    plays a note contributes nothing (its CC at 5 s does not widen the roll); channel 9 is a drum instrument."""
    path = str(tmp_path / "m.mid")
    midi.write_midi_tracks(path, [
        [('cc', 0, 64, 127, 0.1), ('note', 0, 60, 100, 0.2, 0.5), ('cc', 0, 64, 0, 0.8), ('bend', 0, 1000, 0.3)],
        [('program', 1, 5, 0.0), ('note', 1, 72, 50, 0.0, 0.25), ('note', 9, 36, 127, 0.0, 2.0), ('cc', 3, 64, 127, 5.0)],
    ])
    mf = midi.read_midi_file(path)
    assert [(i.program, i.is_drum, i.n_notes) for i in mf.instruments] == [(0, False, 1), (5, False, 1), (0, True, 1)]
    a, b, d = mf.instruments
    assert [(n, v) for n, v, _ in a.control_changes] == [(64, 127), (64, 0)] and a.pitch_bends[0][0] == 1000
    assert np.allclose([t for _, _, t in a.control_changes], [0.1, 0.8], atol=2e-3)
    assert b.control_changes == [] and d.control_changes == []
    assert abs(a.get_end_time() - 0.8) < 2e-3 and abs(d.get_end_time() - 2.0) < 2e-3 and abs(mf.get_end_time() - 2.0) < 2e-3


class _Inst:
    def __init__(self, notes, ccs=(), bends=(), is_drum=False):
        self.pitch = [n[0] for n in notes]; self.velocity = [n[1] for n in notes]
        self.start = [n[2] for n in notes]; self.end = [n[3] for n in notes]
        self.control_changes, self.pitch_bends, self.is_drum = list(ccs), list(bends), is_drum


def test_per_instrument_pedal_known_answer():
    """fs = 100.  A: pitch 60 v100 [0, 0.5], pedal down at 0.25, up at 1.0 -> width 100, pitch 60 held to column 99.
    B: pitch 72 v50 [0, 0.5], no pedal -> width 50, ends at column 50 (a merged single instrument would have sustained
    it through A's pedal to column 99).  Sum: width 100."""
    A = _Inst([(60, 100, 0.0, 0.5)], ccs=[(64, 127, 0.25), (64, 0, 1.0)])
    B = _Inst([(72, 50, 0.0, 0.5)])
    roll = opr.prettymidi_piano_roll([A, B], fs=100)
    assert roll.shape == (128, 100)
    assert np.array_equal(roll[60], np.full(100, 100.0))
    assert np.array_equal(roll[72], np.r_[np.full(50, 50.0), np.zeros(50)])
    assert roll.sum() == 100 * 100 + 50 * 50
    merged = opr.get_piano_roll([60, 72], [100, 50], [0.0, 0.0], [0.5, 0.5], 100, end_time=1.0, cc64=[(0.25, 127), (1.0, 0)])
    assert np.array_equal(merged[72], np.full(100, 50.0))      # what merging the instruments would (wrongly) give


def test_pitch_bend_known_answer():
    """fs = 100, one note pitch 60 v80 [0, 1.0].  Bends: +4096 (one semitone: int 1, decimal 0) at 0.25;
    +2048 (half a semitone: int 0, decimal 0.5) at 0.5; -8192 (two semitones down: int -2, decimal 0) at 0.75; the zero
    bend pretty_midi appends at end_time = 1.0 closes the last segment."""
    I = _Inst([(60, 80, 0.0, 1.0)], bends=[(4096, 0.25), (2048, 0.5), (-8192, 0.75)])
    roll = opr.instrument_piano_roll(I, fs=100)
    assert roll.shape == (128, 100)
    want = np.zeros((128, 100))
    want[60, 0:25] = 80                       # un-bent
    want[61, 25:50] = 80                      # whole row shifted up by one, row 60 empty
    want[60, 50:75] = 40; want[61, 50:75] = 40   # (1 - 0.5) * row + 0.5 * row below
    want[58, 75:100] = 80                     # shifted down by two
    assert np.array_equal(roll, want)
    # drums: zero roll of the right width; an instrument without notes: width 0 even if it has control changes
    D = _Inst([(36, 127, 0.0, 2.0)], is_drum=True)
    E = _Inst([], ccs=[(64, 127, 9.0)])
    assert opr.instrument_piano_roll(D, 100).shape == (128, 200) and opr.instrument_piano_roll(D, 100).sum() == 0
    assert opr.instrument_piano_roll(E, 100).shape == (128, 0)
    total = opr.prettymidi_piano_roll([I, D, E], 100)
    assert total.shape == (128, 200) and np.array_equal(total[:, :100], want) and total[:, 100:].sum() == 0


def test_bend_segments_host_logic():
    from ml_music_style_transfer_b200.pianoroll import bend_segments
    segs = bend_segments([(2048, 0.5), (4096, 0.25), (0, 0.6), (-8192, 0.75), (-100, 0.9)], 1.0, 100)
    # sorted by time; the zero bend at 0.6 ends the +2048 segment early and is itself inactive
    assert [(s[0], s[1], s[4], s[5]) for s in segs] == [(25, 50, 1, 1), (50, 60, 0, 1), (75, 90, -2, 0), (90, 100, 0, 0)]
    assert segs[0][2] == 0.0 and segs[1][2] == 0.5 and segs[2][2] == 0.0 and segs[3][2] == 100 * 2.0 / 8192.0
    assert all(s[3] == 1 - s[2] for s in segs)


def test_reader_on_hand_assembled_file(tmp_path):
    """A type-1 file assembled byte by byte from the SMF specification (not by the repo's writer): running status,
    note-on with velocity 0 as note-off, a tempo change in track 0, program change, CC64, pitch bend, a sysex and an
    unknown meta event in between, a two-byte variable-length delta.  480 ticks per quarter; 120 bpm until tick 960,
    then 60 bpm: tick 480 = 0.5 s, 960 = 1.0 s, 1440 = 2.0 s, 1920 = 3.0 s."""
    from ml_music_style_transfer_b200 import midi

    def chunk(tag, body):
        return tag + len(body).to_bytes(4, 'big') + body

    track0 = (b'\x00\xFF\x03\x05tempo'                      # track name
              b'\x00\xFF\x58\x04\x04\x02\x18\x08'           # time signature (ignored)
              b'\x87\x40\xFF\x51\x03\x0F\x42\x40'           # delta 960 (0x87 0x40): tempo 1 000 000 us/quarter = 60 bpm
              b'\x00\xFF\x2F\x00')
    track1 = (b'\x00\xC0\x05'                               # program 5 on channel 0
              b'\x00\xF0\x03\x7E\x7F\xF7'                   # sysex, 3 data bytes
              b'\x00\x90\x3C\x64'                           # tick 0: note-on 60 vel 100
              b'\x00\x40\x50'                               # running status: note-on 64 vel 80
              b'\x83\x60\x3C\x00'                           # delta 480 (0x83 0x60): running status note-on 60 vel 0 = note-off at 0.5 s
              b'\x00\xB0\x40\x7F'                           # CC64 = 127 at tick 480
              b'\x83\x60\x80\x40\x00'                       # tick 960: note-off 64 -> 1.0 s
              b'\x00\xFF\x7F\x02\x00\x01'                   # sequencer-specific meta (unknown to the reader)
              b'\x83\x60\xE0\x00\x50'                       # tick 1440 = 2.0 s: pitch bend 0x50 << 7 = 10240 -> +2048
              b'\x00\x90\x43\x7F'                           # note-on 67 vel 127 at 2.0 s
              b'\x83\x60\x80\x43\x40'                       # tick 1920 = 3.0 s: note-off 67
              b'\x00\xB0\x40\x00'                           # CC64 = 0 at 3.0 s
              b'\x00\xFF\x2F\x00')
    data = chunk(b'MThd', (1).to_bytes(2, 'big') + (2).to_bytes(2, 'big') + (480).to_bytes(2, 'big')) + \
        chunk(b'MTrk', track0) + chunk(b'MTrk', track1)
    path = str(tmp_path / "hand.mid")
    with open(path, 'wb') as f:
        f.write(data)
    mf = midi.read_midi_file(path)
    assert mf.resolution == 480 and len(mf.instruments) == 1
    ins = mf.instruments[0]
    assert ins.program == 5 and not ins.is_drum
    notes = sorted(zip(ins.pitch, ins.velocity, ins.start, ins.end))
    assert notes == [(60, 100, 0.0, 0.5), (64, 80, 0.0, 1.0), (67, 127, 2.0, 3.0)]
    assert ins.control_changes == [(64, 127, 0.5), (64, 0, 3.0)]
    assert ins.pitch_bends == [(2048, 2.0)]
    assert mf.get_end_time() == 3.0
    # the oracle's roll of that instrument: the pedal goes down at 0.5 s, exactly when 60 ends (not sustained: the
    # running maximum starts inside the span), while 64 is sounding and is held; the bend only acts from 2.0 s on
    from oracle import pianoroll as opr
    roll = opr.instrument_piano_roll(ins, 100)
    assert roll.shape == (128, 300)
    assert (roll[60, 0:50] == 100).all() and (roll[60, 50:200] == 0).all() and (roll[64, 0:200] == 80).all()
    assert roll[67, 200:].sum() > 0 and roll[68, 200:].sum() > 0   # +0.5 semitone: 67 is split between rows 67 and 68
