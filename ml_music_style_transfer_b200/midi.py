"""Standard MIDI File reader (host side): the part of ``pretty_midi.PrettyMIDI(path)`` that ``get_piano_roll`` needs.

Stands in for ``pretty_midi.PrettyMIDI(path)`` at preprocessing/preprocess.py:146 and model/inference.py:40.  Followed
from pretty_midi 0.2.9 (``pretty_midi.py``: ``_load_tempo_changes``, ``_update_tick_to_time``, ``_load_instruments``):

  * type 0/1 files, PPQ time division; tempo map from the set_tempo events of TRACK 0 only, first tempo 120 bpm;
    ``time(tick) = time(segment start) + (tick - segment start) * tick_scale`` in float64;
  * instruments are keyed by (program, channel, track); a note creates its instrument, control changes and pitch
    bends attach to the existing instrument of that key, else to the (channel, track) "straggler" whose event lists a
    later note-created instrument on that (channel, track) adopts; events on a (channel, track) that never plays a note
    are dropped (they are not part of ``PrettyMIDI.instruments``);
  * note-on / note-off pairing per (channel, pitch) within a track: a note-off closes every open note-on of that key
    whose start tick differs from the off tick; note-ons of the SAME tick survive only if something was closed,
    otherwise the key is forgotten (so a zero-length note never leaves a dangling note-on);
  * channel 9 makes a drum instrument (its roll is all zeros, but its width counts).

CC64 sustain, pitch bends and the per-instrument sum are applied by ``pianoroll.midi_to_pianoroll`` on the device.
"""
import struct

import numpy as np


def _read_vlq(data, pos):
    val = 0
    while True:
        b = data[pos]
        pos += 1
        val = (val << 7) | (b & 0x7F)
        if not b & 0x80:
            return val, pos


def _parse_track(data):
    """-> list of (abs_tick, kind, a, b, channel); kind in {'on','off','tempo','cc','bend','program'}."""
    events = []
    pos, tick, status = 0, 0, None
    n = len(data)
    while pos < n:
        delta, pos = _read_vlq(data, pos)
        tick += delta
        b = data[pos]
        if b == 0xFF:  # meta
            mtype = data[pos + 1]
            length, pos = _read_vlq(data, pos + 2)
            if mtype == 0x51 and length == 3:
                events.append((tick, 'tempo', int.from_bytes(data[pos:pos + 3], 'big'), 0, 0))
            pos += length
            if mtype == 0x2F:
                break
            continue
        if b in (0xF0, 0xF7):  # sysex
            length, pos = _read_vlq(data, pos + 1)
            pos += length
            continue
        if b & 0x80:
            status = b
            pos += 1
        if status is None:
            raise ValueError("running status without a status byte")
        hi, ch = status & 0xF0, status & 0x0F
        if hi in (0x80, 0x90, 0xA0, 0xB0, 0xE0):
            a, v = data[pos], data[pos + 1]
            pos += 2
            if hi == 0x90 and v > 0:
                events.append((tick, 'on', a, v, ch))
            elif hi == 0x80 or hi == 0x90:
                events.append((tick, 'off', a, v, ch))
            elif hi == 0xB0:
                events.append((tick, 'cc', a, v, ch))
            elif hi == 0xE0:
                events.append((tick, 'bend', ((v << 7) | a) - 8192, 0, ch))  # 14-bit value, centre 8192 (mido's .pitch)
        elif hi == 0xC0:
            events.append((tick, 'program', data[pos], 0, ch))
            pos += 1
        elif hi == 0xD0:
            pos += 1
        else:
            raise ValueError(f"unexpected status byte {status:#x}")
    return events


class Instrument:
    """pretty_midi.Instrument as far as get_piano_roll reads it."""

    def __init__(self, program, is_drum=False):
        self.program, self.is_drum = int(program), bool(is_drum)
        self.pitch, self.velocity, self.start, self.end = [], [], [], []
        self.control_changes = []  # [(number, value, time)] in file order
        self.pitch_bends = []      # [(pitch in [-8192, 8191], time)] in file order

    @property
    def n_notes(self):
        return len(self.pitch)

    def get_end_time(self):
        """pretty_midi Instrument.get_end_time: latest note end, pitch bend or control change."""
        events = list(self.end) + [t for _, t in self.pitch_bends] + [t for _, _, t in self.control_changes]
        return float(max(events)) if events else 0.0

    def cc64(self):
        return [(t, v) for num, v, t in self.control_changes if num == 64]

    def arrays(self):
        return (np.array(self.pitch, dtype=np.int32), np.array(self.velocity, dtype=np.int32),
                np.array(self.start, dtype=np.float64), np.array(self.end, dtype=np.float64))


class MidiFile:
    def __init__(self, resolution, instruments):
        self.resolution, self.instruments = resolution, instruments

    def get_end_time(self):
        return max([i.get_end_time() for i in self.instruments], default=0.0)


def read_midi_file(path):
    """-> MidiFile whose ``instruments`` are pretty_midi's (same membership and order, same note lists and event lists)."""
    with open(path, 'rb') as f:
        data = f.read()
    if data[:4] != b'MThd':
        raise ValueError("not a Standard MIDI File")
    hlen, fmt, ntrks, division = struct.unpack('>IHHH', data[4:14])
    if division & 0x8000:
        raise ValueError("SMPTE time division is not supported")
    pos = 8 + hlen
    tracks = []
    for _ in range(ntrks):
        if data[pos:pos + 4] != b'MTrk':
            raise ValueError("bad track chunk")
        tlen = struct.unpack('>I', data[pos + 4:pos + 8])[0]
        tracks.append(_parse_track(data[pos + 8:pos + 8 + tlen]))
        pos += 8 + tlen
    # tempo map: set_tempo events of track 0 only (pretty_midi _load_tempo_changes), first tempo 120 bpm at tick 0
    seg_tick, seg_scale = [0], [60.0 / (120.0 * division)]
    for t, kind, us, _, _ in (tracks[0] if tracks else []):
        if kind != 'tempo':
            continue
        if t == 0:
            seg_tick, seg_scale = [0], [60.0 / ((6e7 / us) * division)]
        else:
            scale = 60.0 / ((6e7 / us) * division)
            if scale != seg_scale[-1]:
                seg_tick.append(t)
                seg_scale.append(scale)
    # _update_tick_to_time: each segment restarts from the time of the previous segment's last tick
    seg_time = [0.0]
    for i in range(1, len(seg_tick)):
        seg_time.append(seg_time[-1] + seg_scale[i - 1] * (seg_tick[i] - seg_tick[i - 1]))
    seg_tick_a, seg_time_a, seg_scale_a = np.array(seg_tick), np.array(seg_time), np.array(seg_scale)

    def to_time(tick):
        i = int(np.searchsorted(seg_tick_a, tick, side='right')) - 1
        return float(seg_time_a[i] + seg_scale_a[i] * (tick - seg_tick_a[i]))

    instrument_map, stragglers = {}, {}   # dicts keep insertion order, like pretty_midi's OrderedDict

    def get_instrument(program, channel, track, create_new):
        if (program, channel, track) in instrument_map:
            return instrument_map[(program, channel, track)]
        if not create_new and (channel, track) in stragglers:
            return stragglers[(channel, track)]
        if create_new:
            inst = Instrument(program, channel == 9)
            if (channel, track) in stragglers:  # adopt (share) the straggler's event lists
                inst.control_changes = stragglers[(channel, track)].control_changes
                inst.pitch_bends = stragglers[(channel, track)].pitch_bends
            instrument_map[(program, channel, track)] = inst
        else:
            inst = Instrument(program)
            stragglers[(channel, track)] = inst
        return inst

    for track_idx, tr in enumerate(tracks):
        last_note_on = {}
        current_instrument = [0] * 16
        for tick, kind, a, v, ch in tr:
            if kind == 'program':
                current_instrument[ch] = a
            elif kind == 'on':
                last_note_on.setdefault((ch, a), []).append((tick, v))
            elif kind == 'off':
                key = (ch, a)
                if key in last_note_on:
                    open_notes = last_note_on[key]
                    to_close = [(st, vv) for st, vv in open_notes if st != tick]
                    to_keep = [(st, vv) for st, vv in open_notes if st == tick]
                    for st, vv in to_close:
                        inst = get_instrument(current_instrument[ch], ch, track_idx, 1)
                        inst.pitch.append(a); inst.velocity.append(vv)
                        inst.start.append(to_time(st)); inst.end.append(to_time(tick))
                    if to_close and to_keep:
                        last_note_on[key] = to_keep   # same-tick note-on continues after the notes just closed
                    else:
                        del last_note_on[key]
            elif kind == 'bend':
                get_instrument(current_instrument[ch], ch, track_idx, 0).pitch_bends.append((a, to_time(tick)))
            elif kind == 'cc':
                get_instrument(current_instrument[ch], ch, track_idx, 0).control_changes.append((a, v, to_time(tick)))
    return MidiFile(division, list(instrument_map.values()))


def read_midi(path):
    """Flat view for single-instrument files: (pitch, velocity, start, end, cc64 [(time, value), ...], end_time).

    Notes of every non-drum instrument are concatenated in instrument order, the CC64 events of all instruments merged
    in time order and ``end_time`` is the latest event of any instrument.  This equals pretty_midi's roll exactly when
    the file has ONE pitched instrument and no pitch bends (the reference's piano MIDI); ``pianoroll.midi_to_pianoroll``
    is the general path (per-instrument pedal, pitch bends, instrument sum)."""
    mf = read_midi_file(path)
    pitched = [i for i in mf.instruments if not i.is_drum]
    cat = lambda j, dt: np.concatenate([i.arrays()[j] for i in pitched]) if pitched else np.zeros(0, dt)
    cc = sorted((t, n, v) for n, inst in enumerate(mf.instruments) for t, v in inst.cc64())
    return (cat(0, np.int32), cat(1, np.int32), cat(2, np.float64), cat(3, np.float64),
            [(t, v) for t, _, v in cc], float(mf.get_end_time()))


def read_midi_notes(path):
    """-> (pitch int32[], velocity int32[], start float64[], end float64[])."""
    return read_midi(path)[:4]


def _vlq(x):
    out = [x & 0x7F]
    x >>= 7
    while x:
        out.append((x & 0x7F) | 0x80)
        x >>= 7
    return bytes(reversed(out))


def write_midi_tracks(path, tracks, ticks_per_beat=480, bpm=120.0, tempo_changes=()):
    """Type-1 writer for tests and synthetic corpora.  ``tracks``: list of event lists with events
    ('note', ch, pitch, vel, start_s, end_s) | ('cc', ch, number, value, t_s) | ('bend', ch, pitch14, t_s) |
    ('program', ch, program, t_s) | ('raw', tick, status, a, b).  Times are quantised to ticks at the initial tempo
    (``tempo_changes`` = [(tick, bpm), ...] go into track 0 and only matter to the reader)."""
    scale = ticks_per_beat * bpm / 60.0
    chunks = []
    for ti, events in enumerate(tracks):
        ev = []
        order = 0
        for e in events:
            order += 1
            if e[0] == 'note':
                _, ch, p, v, s, t = e
                ev.append((int(round(s * scale)), 2, order, bytes([0x90 | ch, int(p), int(v)])))
                ev.append((int(round(t * scale)), 0, order, bytes([0x80 | ch, int(p), 0])))
            elif e[0] == 'cc':
                _, ch, num, val, t = e
                ev.append((int(round(t * scale)), 1, order, bytes([0xB0 | ch, int(num), int(val)])))
            elif e[0] == 'bend':
                _, ch, pitch, t = e
                u = int(pitch) + 8192
                ev.append((int(round(t * scale)), 1, order, bytes([0xE0 | ch, u & 0x7F, (u >> 7) & 0x7F])))
            elif e[0] == 'program':
                _, ch, prog, t = e
                ev.append((int(round(t * scale)), 0, order, bytes([0xC0 | ch, int(prog)])))
            elif e[0] == 'raw':
                _, tick, st, a, b = e
                ev.append((int(tick), 1, order, bytes([st, a, b])))
            else:
                raise ValueError(e[0])
        if ti == 0:
            ev.append((0, -1, 0, b'\xFF\x51\x03' + int(round(6e7 / bpm)).to_bytes(3, 'big')))
            for tick, b2 in tempo_changes:
                ev.append((int(tick), -1, 0, b'\xFF\x51\x03' + int(round(6e7 / b2)).to_bytes(3, 'big')))
        ev.sort(key=lambda x: (x[0], x[1], x[2]))
        body = bytearray()
        last = 0
        for t, _, _, payload in ev:
            body += _vlq(t - last) + payload
            last = t
        body += b'\x00\xFF\x2F\x00'
        chunks.append(b'MTrk' + struct.pack('>I', len(body)) + bytes(body))
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 1 if len(tracks) > 1 else 0, len(tracks), ticks_per_beat))
        for c in chunks:
            f.write(c)


def write_midi_notes(path, pitch, velocity, start, end, ticks_per_beat=480, bpm=120.0, cc64=None):
    """Single-track, channel-0 file from a note list (+ optional CC64 events)."""
    ev = [('note', 0, int(p), int(v), float(s), float(e)) for p, v, s, e in zip(pitch, velocity, start, end)]
    ev += [('cc', 0, 64, int(v), float(t)) for t, v in (cc64 or [])]
    write_midi_tracks(path, [ev], ticks_per_beat, bpm)
