"""pretty_midi-0.2.9 ``get_piano_roll`` (notes only) + the reference's binarise / on-off /
chunking lines restated in NumPy (oracle; test infrastructure).

Reference: preprocessing/preprocess.py:146-155 (roll -> binarise -> on/off), :80-96 (chunking),
:118-136 (chunk count); duplicated at model/inference.py:40-51.  The audio-rate upsampling is
README-only in the reference (README.md:19-20); its definition here is SURVEY.md section 8a row P3d:
``up[k, n] = x[col(n), pitch_lo + k]`` with ``col(n) = (n * fs) // sr`` in 64-bit integers and
zero where ``col(n) >= T``.

pretty_midi semantics followed (pretty_midi is absent from this image):
  Instrument.get_piano_roll(fs): ``roll = zeros((128, int(fs * end_time)))`` with end_time the
  latest note end; for each note ``roll[pitch, int(start*fs):int(end*fs)] += velocity`` (float64
  products truncated toward zero by ``int``); PrettyMIDI.get_piano_roll sums the instruments
  into the widest roll; CC64 sustain spans keep a running maximum; pitch bends shift / interpolate rows
  (``instrument_piano_roll`` / ``prettymidi_piano_roll`` restate those statement by statement).
"""
import numpy as np

__all__ = ["get_piano_roll", "binarize_and_onoff", "onoff_reference_loop", "upsample_to_audio_rate",
           "process_pianoroll_into_chunks", "get_num_song_chunks", "instrument_piano_roll", "prettymidi_piano_roll"]


def get_piano_roll(pitch, velocity, start, end, fs, end_time=None, cc64=None, pedal_threshold=64):
    """(128, T) float64 velocity-sum roll.  ``start`` / ``end`` are float64 seconds.  ``cc64``: [(time, value), ...]
    sustain-pedal events -> pretty_midi >= 0.2.9's running-maximum rule inside pedal-down spans."""
    pitch = np.asarray(pitch, dtype=np.int64)
    velocity = np.asarray(velocity, dtype=np.int64)
    start = np.asarray(start, dtype=np.float64)
    end = np.asarray(end, dtype=np.float64)
    if len(pitch) == 0:
        return np.zeros((128, 0))
    if end_time is None:
        end_time = float(end.max())
    roll = np.zeros((128, int(fs * end_time)))
    for p, v, s, e in zip(pitch, velocity, start, end):
        roll[int(p), int(float(s) * fs):int(float(e) * fs)] += int(v)
    if cc64 is not None and pedal_threshold is not None:
        time_pedal_on, is_pedal_on = 0, False
        for t, val in cc64:
            time_now = int(float(t) * fs)
            is_current_pedal_on = val >= pedal_threshold
            if not is_pedal_on and is_current_pedal_on:
                time_pedal_on, is_pedal_on = time_now, True
            elif is_pedal_on and not is_current_pedal_on:
                subpr = roll[:, time_pedal_on:time_now]
                roll[:, time_pedal_on:time_now] = np.maximum.accumulate(subpr, axis=1)
                is_pedal_on = False
    return roll


def instrument_piano_roll(inst, fs=100, pedal_threshold=64):
    """pretty_midi 0.2.9 ``Instrument.get_piano_roll(fs, times=None, pedal_threshold)`` restated statement by
    statement: notes, CC64 running maximum, pitch bends (integer row shift + linear interpolation by the fractional
    part).  ``inst`` exposes is_drum, pitch / velocity / start / end (parallel sequences), control_changes
    [(number, value, time)] and pitch_bends [(pitch, time)] in file order."""
    if len(inst.pitch) == 0:
        return np.array([[]] * 128)
    events = [float(e) for e in inst.end] + [t for _, t in inst.pitch_bends] + [t for _, _, t in inst.control_changes]
    end_time = max(events)
    piano_roll = np.zeros((128, int(fs * end_time)))
    if inst.is_drum:
        return piano_roll
    for p, v, s, e in zip(inst.pitch, inst.velocity, inst.start, inst.end):
        piano_roll[int(p), int(float(s) * fs):int(float(e) * fs)] += int(v)
    if pedal_threshold is not None:
        time_pedal_on, is_pedal_on = 0, False
        for num, val, t in inst.control_changes:
            if num != 64:
                continue
            time_now = int(t * fs)
            is_current_pedal_on = val >= pedal_threshold
            if not is_pedal_on and is_current_pedal_on:
                time_pedal_on, is_pedal_on = time_now, True
            elif is_pedal_on and not is_current_pedal_on:
                subpr = piano_roll[:, time_pedal_on:time_now]
                piano_roll[:, time_pedal_on:time_now] = np.maximum.accumulate(subpr, axis=1)
                is_pedal_on = False
    ordered_bends = sorted(inst.pitch_bends, key=lambda bend: bend[1])
    for (start_pitch_raw, start_t), (_, end_t) in zip(ordered_bends, ordered_bends[1:] + [(0, end_time)]):
        if np.abs(start_pitch_raw) < 1:
            continue
        start_pitch = 2.0 * start_pitch_raw / 8192.0            # pitch_bend_to_semitones, semitone_range=2
        bend_int = int(np.sign(start_pitch) * np.floor(np.abs(start_pitch)))
        bend_decimal = np.abs(start_pitch - bend_int)
        bend_range = np.r_[int(start_t * fs):int(end_t * fs)]
        bent_roll = np.zeros(piano_roll[:, bend_range].shape)
        if start_pitch_raw >= 0:
            if bend_int != 0:
                bent_roll[bend_int:] = piano_roll[:-bend_int, bend_range]
            else:
                bent_roll = piano_roll[:, bend_range]
            bent_roll[1:] = ((1 - bend_decimal) * bent_roll[1:] + bend_decimal * bent_roll[:-1])
        else:
            if bend_int != 0:
                bent_roll[:bend_int] = piano_roll[-bend_int:, bend_range]
            else:
                bent_roll = piano_roll[:, bend_range]
            bent_roll[:-1] = ((1 - bend_decimal) * bent_roll[:-1] + bend_decimal * bent_roll[1:])
        piano_roll[:, bend_range] = bent_roll
    return piano_roll


def prettymidi_piano_roll(instruments, fs=100, pedal_threshold=64):
    """pretty_midi 0.2.9 ``PrettyMIDI.get_piano_roll(fs)``: per-instrument rolls summed into the widest one."""
    if len(instruments) == 0:
        return np.zeros((128, 0))
    piano_rolls = [instrument_piano_roll(i, fs, pedal_threshold) for i in instruments]
    piano_roll = np.zeros((128, np.max([p.shape[1] for p in piano_rolls])))
    for roll in piano_rolls:
        piano_roll[:, :roll.shape[1]] += roll
    return piano_roll


def onoff_reference_loop(pianoroll):
    """The literal loop of preprocess.py:149-155 (setdiff1d form); O(T) Python iterations."""
    onoff = np.zeros(pianoroll.shape)
    for i in range(pianoroll.shape[0]):
        if i == 0:
            onoff[i][pianoroll[i].nonzero()] = 1
        else:
            onoff[i][np.setdiff1d(pianoroll[i - 1].nonzero(), pianoroll[i].nonzero())] = -1
            onoff[i][np.setdiff1d(pianoroll[i].nonzero(), pianoroll[i - 1].nonzero())] = 1
    return onoff


def binarize_and_onoff(roll_128_T):
    """preprocess.py:147-155: transpose to (T,128), binarise, first difference with a zero row."""
    pianoroll = np.array(roll_128_T, dtype=np.float64).T.copy()
    pianoroll[pianoroll.nonzero()] = 1
    onoff = np.empty_like(pianoroll)
    if pianoroll.shape[0]:
        onoff[0] = pianoroll[0]
        onoff[1:] = pianoroll[1:] - pianoroll[:-1]
    return pianoroll, onoff


def upsample_to_audio_rate(x_T_128, fs, sr, n_samples, pitch_lo=21, n_keys=88, dtype=np.int8):
    """(n_keys, n_samples) hold-replication of a (T,128) frame-rate plane."""
    x = np.asarray(x_T_128)
    T = x.shape[0]
    n = np.arange(n_samples, dtype=np.int64)
    col = (n * int(fs)) // int(sr)
    valid = col < T
    out = np.zeros((n_keys, n_samples), dtype=dtype)
    out[:, valid] = x[col[valid], pitch_lo:pitch_lo + n_keys].T.astype(dtype)
    return out


def process_pianoroll_into_chunks(pianoroll, onoff, num_chunks, n_windows_per_chunk=860, stride=512):
    """preprocess.py:80-96."""
    score, oo = [], []
    for step in range(num_chunks):
        score.append(pianoroll[step * stride:step * stride + n_windows_per_chunk])
        oo.append(onoff[step * stride:step * stride + n_windows_per_chunk])
    return np.array(score), np.array(oo)


def get_num_song_chunks(n_rows, offset_percentage=0.1, max_chunks=100, n_windows_per_chunk=860, stride=512):
    """preprocess.py:118-136."""
    num_chunks = (n_rows - n_windows_per_chunk) // stride
    num_chunks -= int(offset_percentage * num_chunks)
    return min(num_chunks, max_chunks)
