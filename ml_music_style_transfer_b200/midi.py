"""Minimal Standard MIDI File reader (host side): note list in seconds, the input of the rasteriser.

Stands in for ``pretty_midi.PrettyMIDI(path)`` at preprocessing/preprocess.py:146 and
model/inference.py:40 as far as the *notes* are concerned: type 0/1 files, tempo map -> seconds,
note-on / note-off pairing per (channel, pitch) with pretty_midi's rule (a note-off closes every open note-on
of that key whose start tick differs from the off tick), drum channel (9) skipped.  CC64 sustain and pitch
bends do not enter the piano roll in this build.
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
    """-> list of (abs_tick, kind, a, b, channel) with kind in {'on','off','tempo'}."""
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
                events.append((tick, 'bend', a, v, ch))
        elif hi in (0xC0, 0xD0):
            pos += 1
        else:
            raise ValueError(f"unexpected status byte {status:#x}")
    return events


def read_midi_notes(path):
    """-> (pitch int32[], velocity int32[], start float64[], end float64[]) in file order of note-offs per track."""
    return read_midi(path)[:4]


def read_midi(path):
    """-> (pitch, velocity, start, end, cc64 [(time, value), ...], end_time).

    ``end_time`` is pretty_midi's get_end_time(): the latest note end, control change or pitch bend of any non-drum
    channel; ``cc64`` are the sustain-pedal events in time order (all non-drum channels merged)."""
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
    # tempo map (pretty_midi: tempo changes of every track, first tempo 120 bpm at tick 0)
    tempi = sorted((t, us) for tr in tracks for (t, kind, us, _, _) in tr if kind == 'tempo')
    seg_tick, seg_time, seg_scale = [0], [0.0], [60.0 / (120.0 * division)]
    for t, us in tempi:
        scale = 60.0 / ((6e7 / us) * division)
        if t == 0:
            seg_scale[0] = scale
            continue
        if scale == seg_scale[-1]:
            continue
        seg_time.append(seg_time[-1] + (t - seg_tick[-1]) * seg_scale[-1])
        seg_tick.append(t)
        seg_scale.append(scale)
    seg_tick_a, seg_time_a, seg_scale_a = np.array(seg_tick), np.array(seg_time), np.array(seg_scale)

    def to_time(tick):
        i = int(np.searchsorted(seg_tick_a, tick, side='right')) - 1
        return float(seg_time_a[i] + (tick - seg_tick_a[i]) * seg_scale_a[i])

    pitch, vel, start, end = [], [], [], []
    cc64, other_times = [], []
    for tr in tracks:
        open_notes = {}
        for tick, kind, a, v, ch in tr:
            if kind == 'cc':
                if ch != 9:
                    other_times.append(to_time(tick))
                    if a == 64:
                        cc64.append((tick, to_time(tick), v))
                continue
            if kind == 'bend':
                if ch != 9:
                    other_times.append(to_time(tick))
                continue
            if kind == 'on':
                open_notes.setdefault((ch, a), []).append((tick, v))
            elif kind == 'off':
                key = (ch, a)
                if key in open_notes:
                    keep = []
                    for st, vv in open_notes[key]:
                        if st != tick:
                            if ch != 9:  # drums contribute zeros to the roll
                                pitch.append(a); vel.append(vv); start.append(to_time(st)); end.append(to_time(tick))
                        else:
                            keep.append((st, vv))
                    if keep:
                        open_notes[key] = keep
                    else:
                        del open_notes[key]
    cc64.sort(key=lambda e: e[0])
    end_time = max(list(end) + other_times) if (len(end) or other_times) else 0.0
    return (np.array(pitch, dtype=np.int32), np.array(vel, dtype=np.int32), np.array(start, dtype=np.float64),
            np.array(end, dtype=np.float64), [(t, v) for _, t, v in cc64], float(end_time))


def write_midi_notes(path, pitch, velocity, start, end, ticks_per_beat=480, bpm=120.0, cc64=None):
    """Tiny type-0 writer used by tests and synthetic corpora (times quantised to ticks)."""
    scale = ticks_per_beat * bpm / 60.0
    ev = []
    for p, v, s, e in zip(pitch, velocity, start, end):
        ev.append((int(round(s * scale)), 1, 0x90, int(p), int(v)))
        ev.append((int(round(e * scale)), 0, 0x80, int(p), 0))
    for t, v in (cc64 or []):
        ev.append((int(round(t * scale)), 2, 0xB0, 64, int(v)))
    ev.sort()

    def vlq(x):
        out = [x & 0x7F]
        x >>= 7
        while x:
            out.append((x & 0x7F) | 0x80)
            x >>= 7
        return bytes(reversed(out))

    body = bytearray()
    body += b'\x00\xFF\x51\x03' + int(round(6e7 / bpm)).to_bytes(3, 'big')
    last = 0
    for t, _, st, a, b in ev:
        body += vlq(t - last) + bytes([st, a, b])
        last = t
    body += b'\x00\xFF\x2F\x00'
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 0, 1, ticks_per_beat))
        f.write(b'MTrk' + struct.pack('>I', len(body)) + bytes(body))
