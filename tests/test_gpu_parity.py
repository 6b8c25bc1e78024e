"""GPU parity tests (-m gpu): the CUDA path, called through torch.ops.mst_b200 -> C ABI, against the NumPy oracle
on the same seeded inputs.  Tolerances are BASELINE.json's: STFT / mel <= 1e-4 relative (float32), piano roll
bit-exact, Griffin-Lim spectral convergence within 1e-3 of the oracle at equal iteration count."""
import numpy as np
import pytest
import torch

from oracle import audio as oaudio, griffinlim as ogl, mel as omel, pianoroll as opr, preprocess as opp, stft as ostft

pytestmark = pytest.mark.gpu

TOL = 1e-4


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) else np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rel_max(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def assert_close(a, b, tol=TOL):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert rel_l2(a, b) <= tol and rel_max(a, b) <= tol, (rel_l2(a, b), rel_max(a, b))


@pytest.fixture(scope="module")
def pkg(gpu):
    import ml_music_style_transfer_b200 as p
    return p


def clip(seed, n, kind="piano"):
    from ml_music_style_transfer_b200 import synth
    if kind == "piano":
        return synth.piano_clip(seed, n / 22050.0, 22050)[:n]
    return synth.noise_clip(seed, n)


# ---- P1: STFT ---------------------------------------------------------------------------------
@pytest.mark.parametrize("hop", [256, 512])
@pytest.mark.parametrize("pad", ["reflect", "constant"])
@pytest.mark.parametrize("n", [88200, 30001, 2049, 1025])
def test_stft_complex(pkg, hop, pad, n):
    y = clip(n, n, "noise" if n == 30001 else "piano")
    D = pkg.features.stft(y, hop_length=hop, pad_mode=pad)
    ref = ostft.stft(y, 2048, hop, pad_mode=pad)
    assert D.dtype == np.complex64 and D.flags.f_contiguous
    assert_close(D, ref)


def test_stft_known_answers(pkg):
    n = np.arange(8192)
    D = pkg.features.stft(np.cos(2 * np.pi * 100 * n / 2048).astype(np.float32), hop_length=512)[:, 8]
    assert abs(abs(D[100]) - 512.0) < 0.05 and abs(abs(D[99]) - 256.0) < 0.05 and np.abs(D[104:150]).max() < 0.05
    D = pkg.features.stft(np.ones(8192, dtype=np.float32), hop_length=512)[:, 8]
    assert abs(D[0].real - 1024.0) < 0.01 and abs(abs(D[1]) - 512.0) < 0.01 and np.abs(D[2:]).max() < 0.01
    x = np.zeros(8192, dtype=np.float32)
    x[4096 + 300] = 1.0
    D = pkg.features.stft(x, hop_length=512)[:, 8]
    assert np.allclose(np.abs(D), ostft.hann_window(2048)[1324], atol=1e-5)


@pytest.mark.parametrize("out", ["magnitude", "power", "log1p_power"])
def test_stft_epilogues(pkg, out):
    y = clip(5, 44100)
    ref = np.abs(ostft.stft(y, 2048, 256, out_dtype=np.complex128))
    ref = {"magnitude": ref, "power": ref ** 2, "log1p_power": np.log1p(ref ** 2)}[out]
    got = pkg.features.spectrogram(y, 256, out=out)
    assert got.dtype == np.float32
    assert_close(got, ref)


def test_process_spectrum_from_chunk_dropin(pkg):
    """preprocess.py:47-57 at the repo's own geometry (44.1 kHz, hop 256, 219 904-sample chunk -> 860 frames)."""
    y = clip(11, 219904)
    got = pkg.preprocess.process_spectrum_from_chunk(y)
    ref = opp.process_spectrum_from_chunk(y)
    assert got.shape == (1025, 860) and got.dtype == np.float32 and got.flags.f_contiguous
    assert_close(got, ref.astype(np.float64))


def test_process_audio_into_chunks_dropin(pkg):
    """preprocess.py:60-77: overlapping chunks, each reflect-padded independently, C-contiguous stack."""
    n = 2 * 131072 + 219904
    y = clip(12, n, "noise")
    got = pkg.preprocess.process_audio_into_chunks(y, "cuba", 2240, 3)
    ref = opp.process_audio_into_chunks(y, "cuba", 2240, 3)
    assert got.shape == (3, 1025, 860) and got.flags.c_contiguous
    assert_close(got, ref.astype(np.float64))
    with pytest.raises(ValueError):
        pkg.preprocess.process_audio_into_chunks(y[:300000], "cuba", 2240, 3)


def test_stft_ragged_batch_device_tensors(pkg, gpu):
    """Ragged, overlapping clips in one launch; CUDA tensor in -> CUDA tensor out; both layouts."""
    F = pkg.features
    y = clip(13, 120000, "noise")
    offs = np.array([0, 1000, 50001, 90000], dtype=np.int64)
    lens = np.array([40000, 2049, 33333, 30000], dtype=np.int64)
    a = torch.from_numpy(y).to(gpu)
    b = F.ClipBatch.from_clips(offs, lens, 512, device=gpu)
    fm = F.stft_batch(a, b, "log1p_power", F.FRAME_MAJOR).view(-1, 1025).cpu().numpy()
    bm = F.stft_batch(a, b, "log1p_power", F.BIN_MAJOR).cpu().numpy()
    f0 = 0
    for o, l in zip(offs, lens):
        ref = np.log1p(np.abs(ostft.stft(y[o:o + l], 2048, 512, out_dtype=np.complex128)) ** 2)
        T = ref.shape[1]
        assert_close(fm[f0:f0 + T].T, ref)
        assert_close(bm[f0 * 1025:(f0 + T) * 1025].reshape(1025, T), ref)
        f0 += T
    assert f0 == b.total_frames


def test_raw_c_abi_through_ctypes(pkg, gpu):
    """The boundary itself: include/mst_b200.h bound with ctypes (no torch op in between), as INTEGRATION.md shows."""
    import ctypes
    lib = pkg._lib.cdll()
    lib.mst_batch_total_frames.restype = ctypes.c_int64
    vp = ctypes.c_void_p
    y = clip(14, 50000, "noise")
    a = torch.from_numpy(y).to(gpu)
    offs = (ctypes.c_int64 * 2)(0, 20000)
    lens = (ctypes.c_int64 * 2)(30000, 30000)
    batch = vp()
    assert lib.mst_batch_create(2, offs, lens, 2048, 256, 0, ctypes.byref(batch)) == 0, lib.mst_last_error()
    frames = lib.mst_batch_total_frames(batch)
    assert frames == 2 * (1 + 30000 // 256)
    out = torch.empty(frames * 1025, dtype=torch.float32, device=gpu)
    stream = vp(torch.cuda.current_stream().cuda_stream)
    n0 = lib.mst_launch_count()
    assert lib.mst_stft_f32(vp(a.data_ptr()), batch, 3, 1, vp(out.data_ptr()), stream) == 0, lib.mst_last_error()
    assert lib.mst_launch_count() == n0 + 1
    torch.cuda.synchronize()
    T = frames // 2
    for c, o in enumerate((0, 20000)):
        ref = np.log1p(np.abs(ostft.stft(y[o:o + 30000], 2048, 256, out_dtype=np.complex128)) ** 2)
        assert_close(out[c * T * 1025:(c + 1) * T * 1025].view(1025, T).cpu().numpy(), ref)
    # error reporting: bad layout, null pointer
    assert lib.mst_stft_f32(vp(a.data_ptr()), batch, 3, 7, vp(out.data_ptr()), stream) == -1
    assert b"layout" in lib.mst_last_error()
    assert lib.mst_stft_f32(None, batch, 3, 1, vp(out.data_ptr()), stream) == -1
    lib.mst_batch_destroy(batch)


def test_edge_cases_empty_and_short_inputs(pkg, gpu):
    """Empty / degenerate inputs behave like the reference's NumPy code, or raise the same exception type."""
    pp = pkg.preprocess
    y = clip(15, 300000, "noise")
    out = pp.process_audio_into_chunks(y, "cuba", 1, 0)            # np.array([]) in the reference
    assert out.shape == (0,)
    roll = np.zeros((2000, 128)); roll[10:50, 60] = 1
    a, b = pp.process_pianoroll_into_chunks(roll, roll.copy(), 1, 0)
    assert a.shape[0] == 0 and b.shape[0] == 0
    with pytest.raises(ValueError):                                 # np.pad(reflect) on a too-short signal raises too
        pkg.features.stft(np.zeros(1000, dtype=np.float32), hop_length=256)
    with pytest.raises(ValueError):
        pkg.features.griffinlim(np.ones((1025, 2), dtype=np.float32), n_iter=1, hop_length=256)
    # a piece whose only note rounds to zero columns -> empty roll, and chunking past the end pads with zeros
    r, o = pp.notes_to_pianoroll([60], [5], [0.0], [0.001])
    assert r.shape == (0, 128) and o.shape == (0, 128)
    r, o = pp.notes_to_pianoroll([60, 60], [5, 7], [0.0, 0.1], [1.0, 0.5])   # collision: overlapping same-pitch notes
    assert r[:, 60].sum() == 172 and o[0, 60] == 1 and o[172 - 1, 60] == 0
    vs = pkg.pianoroll.get_piano_roll([60, 60], [5, 7], [0.0, 0.1], [1.0, 0.5], 172)
    assert vs[60, 20] == 12 and vs[60, 10] == 5                     # velocities add where notes overlap
    # silence in, finite zeros out (log1p(0) = 0; Griffin-Lim of an all-zero spectrogram is silence)
    z = pp.process_spectrum_from_chunk(np.zeros(8192, dtype=np.float32))
    assert np.array_equal(z, np.zeros_like(z))
    w = pkg.features.griffinlim(np.zeros((1025, 20), dtype=np.float32), n_iter=4, hop_length=256)
    assert np.isfinite(w).all() and np.abs(w).max() == 0.0


@pytest.mark.parametrize("hop,pad", [(256, "reflect"), (512, "constant")])
def test_stft_all_tile_remainders(pkg, gpu, hop, pad):
    """Ragged batch whose clips cover every frame count mod 8 (partial tiles) and odd sample offsets (unaligned
    loads), in both output layouts."""
    F = pkg.features
    lens = [1025 + 37 * i + hop * (i % 11) for i in range(30)]
    offs = np.concatenate([[0], np.cumsum([l + (i % 3) for i, l in enumerate(lens)])[:-1]])  # some odd offsets
    y = clip(16, int(offs[-1] + lens[-1] + 8), "noise")
    a = torch.from_numpy(y).to(gpu)
    b = F.ClipBatch.from_clips(offs, lens, hop, pad_mode=pad, device=gpu)
    fm = F.stft_batch(a, b, "complex").cpu().numpy()
    bm = F.stft_batch(a, b, "log1p_power", F.BIN_MAJOR).cpu().numpy()
    f0 = 0
    seen = set()
    for o, l in zip(offs, lens):
        ref = ostft.stft(y[o:o + l], 2048, hop, pad_mode=pad, out_dtype=np.complex128)
        T = ref.shape[1]
        seen.add(T % 8)
        assert_close(fm[f0:f0 + T].T, ref)
        assert_close(bm[f0 * 1025:(f0 + T) * 1025].reshape(1025, T), np.log1p(np.abs(ref) ** 2))
        f0 += T
    assert f0 == b.total_frames and len(seen) >= 6


# ---- P2: mel ----------------------------------------------------------------------------------
@pytest.mark.parametrize("sr,hop", [(22050, 512), (44100, 256)])
def test_melspectrogram_and_logmel(pkg, sr, hop):
    y = clip(21, 66150)
    ref = omel.melspectrogram(y, sr, 2048, hop)
    got = pkg.features.melspectrogram(y=y, sr=sr, n_fft=2048, hop_length=hop)
    assert got.shape == ref.shape == (128, 1 + len(y) // hop) and got.dtype == np.float32
    assert_close(got, ref.astype(np.float64))
    assert_close(pkg.features.logmel(y, sr=sr, hop_length=hop), np.log1p(ref.astype(np.float64)))


def test_mel_frame_major_batch(pkg, gpu):
    F = pkg.features
    y = clip(22, 4 * 88200, "noise")
    a = torch.from_numpy(y).to(gpu)
    b = F.ClipBatch.uniform(4, 88200, 512, device=gpu)
    plan = F.MelPlan.get(22050, device=gpu)
    got = F.melspectrogram_batch(a, b, plan, log1p=True, layout=F.FRAME_MAJOR).view(4, 173, 128).cpu().numpy()
    for c in range(4):
        assert_close(got[c].T, omel.logmel(y[c * 88200:(c + 1) * 88200], 22050, 2048, 512).astype(np.float64))


def test_mel_multi_chunk_ragged_and_80_mels(pkg, gpu):
    """More frames than one ring chunk (8 waves of 148 x 128 rows), ragged clips, frame tiles straddling projection
    tiles, and a filterbank with a different band structure (80 mels)."""
    F = pkg.features
    lens = [88200, 30001, 2049, 66150] * 420  # 1 680 clips, 154 140 frames > 151 552
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    y = clip(23, int(sum(lens)), "noise")
    a = torch.from_numpy(y).to(gpu)
    b = F.ClipBatch.from_clips(offs, lens, 512, device=gpu)
    assert b.total_frames > 8 * 148 * 128
    for n_mels in (128, 80):
        plan = F.MelPlan.get(22050, n_mels=n_mels, device=gpu)
        bm = F.melspectrogram_batch(a, b, plan, log1p=False, layout=F.BIN_MAJOR).cpu().numpy()
        fm = F.melspectrogram_batch(a, b, plan, log1p=True, layout=F.FRAME_MAJOR).view(-1, n_mels).cpu().numpy()
        f0 = 0
        for i, (o, l) in enumerate(zip(offs, lens)):
            T = 1 + l // 512
            if i in (0, 1, 2, 3, 840, 1575, 1576, 1577, 1578, 1579, 1580, 1581, 1582, 1679):  # incl. the chunk boundary (clip 1579)
                ref = omel.melspectrogram(y[o:o + l], 22050, 2048, 512, n_mels).astype(np.float64)
                assert_close(bm[f0 * n_mels:(f0 + T) * n_mels].reshape(n_mels, T), ref)
                assert_close(fm[f0:f0 + T].T, np.log1p(ref))
            f0 += T


# ---- P3: piano roll (bit-exact) -------------------------------------------------------------
def test_pianoroll_known_answer(pkg):
    roll, onoff = pkg.preprocess.notes_to_pianoroll([60, 60, 64, 21], [100, 90, 80, 1], [0, 0.5, 0.25, 0.999],
                                                    [0.5, 1.0, 0.2501, 2.0])
    assert roll.shape == (344, 128) and roll.dtype == np.float64
    on = np.nonzero(onoff[:, 60])[0]
    assert list(on) == [0, 172] and list(onoff[on, 60]) == [1.0, -1.0]
    assert roll[:, 64].sum() == 0
    # int(0.29*100) == 28: truncation of the float64 product
    r2, _ = pkg.preprocess.notes_to_pianoroll([50], [10], [0.29], [0.5], fs=100)
    assert r2[28, 50] == 1 and r2[27, 50] == 0
    vs = pkg.pianoroll.get_piano_roll([60, 60, 64, 21], [100, 90, 80, 1], [0, 0.5, 0.25, 0.999], [0.5, 1.0, 0.2501, 2.0], 172)
    assert np.array_equal(vs, opr.get_piano_roll([60, 60, 64, 21], [100, 90, 80, 1], [0, 0.5, 0.25, 0.999], [0.5, 1.0, 0.2501, 2.0], 172))


@pytest.mark.parametrize("fs,sr,pitch_lo,n_keys", [(250, 22050, 21, 88), (172, 44100, 0, 128)])
def test_pianoroll_batch_bit_exact(pkg, gpu, fs, sr, pitch_lo, n_keys):
    from ml_music_style_transfer_b200 import synth
    P = pkg.pianoroll
    pieces = [synth.midi_piece(i, seconds=4.0 + i) for i in range(5)]
    pieces.append((np.array([60], np.int32), np.array([5], np.int32), np.array([0.0]), np.array([0.001])))  # empty roll
    nb = P.NoteBatch.from_pieces(pieces, device=gpu)
    roll, onoff, row_off, velsum = P.rasterize(nb, fs, want_velsum=True)
    ro = row_off.cpu().numpy()
    n_samp = [int(sr * 1.5) + 7 * i for i in range(len(pieces))]
    up, so = P.upsample(roll, row_off, n_samp, fs, sr, pitch_lo, n_keys, torch.int8)
    up_oo, _ = P.upsample(onoff, row_off, n_samp, fs, sr, pitch_lo, n_keys, torch.float32)
    # roll + on/off in ONE launch (shared index arithmetic) == the two single-plane launches, int8 and float32
    pa, pb, so2 = P.upsample_pair(roll, onoff, row_off, n_samp, fs, sr, pitch_lo, n_keys, torch.int8)
    up_oo8, _ = P.upsample(onoff, row_off, n_samp, fs, sr, pitch_lo, n_keys, torch.int8)
    assert np.array_equal(so2, so) and torch.equal(pa, up) and torch.equal(pb, up_oo8)
    fa, fb, _ = P.upsample_pair(roll, onoff, row_off, n_samp, fs, sr, pitch_lo, n_keys, torch.float32)
    assert torch.equal(fb, up_oo) and torch.equal(fa, up.to(torch.float32))
    up, up_oo = up.cpu().numpy(), up_oo.cpu().numpy()
    for i, (p, v, s, e) in enumerate(pieces):
        ref_v = opr.get_piano_roll(p, v, s, e, fs)
        ref_r, ref_o = opr.binarize_and_onoff(ref_v)
        T = ref_v.shape[1]
        assert ro[i + 1] - ro[i] == T
        assert np.array_equal(velsum[ro[i]:ro[i + 1]].cpu().numpy().T, ref_v)
        assert np.array_equal(roll[ro[i]:ro[i + 1]].cpu().numpy(), ref_r)
        assert np.array_equal(onoff[ro[i]:ro[i + 1]].cpu().numpy(), ref_o)
        N = n_samp[i]
        blk = up[n_keys * so[i]:n_keys * so[i + 1]].reshape(n_keys, N)
        assert np.array_equal(blk, opr.upsample_to_audio_rate(ref_r, fs, sr, N, pitch_lo, n_keys, np.int8))
        blk = up_oo[n_keys * so[i]:n_keys * so[i + 1]].reshape(n_keys, N)
        assert np.array_equal(blk, opr.upsample_to_audio_rate(ref_o, fs, sr, N, pitch_lo, n_keys, np.float32))


def test_pianoroll_sustain_pedal(pkg, gpu, tmp_path):
    """CC64 sustain: device running-max over pedal spans == pretty_midi rule (oracle), bit exact, batch + file."""
    from ml_music_style_transfer_b200 import midi, synth
    P = pkg.pianoroll
    rng = np.random.default_rng(4)
    pieces, pedals, ends = [], [], []
    for i in range(4):
        p, v, s, e = synth.midi_piece(20 + i, seconds=6.0)
        times = np.sort(rng.uniform(0, 6.5, 14))
        cc = [(float(t), int(rng.choice([0, 20, 64, 100, 127]))) for t in times]
        pieces.append((p, v, s, e)); pedals.append(cc); ends.append(max(float(e.max()), cc[-1][0]))
    pedals[3] = []  # a piece without pedal events
    nb = P.NoteBatch.from_pieces(pieces, device=gpu)
    nb.pedals, nb.end_times = pedals, ends
    roll, onoff, row_off, velsum = P.rasterize(nb, 172)
    ro = row_off.cpu().numpy()
    for i, (p, v, s, e) in enumerate(pieces):
        ref_v = opr.get_piano_roll(p, v, s, e, 172, end_time=ends[i], cc64=pedals[i])
        ref_r, ref_o = opr.binarize_and_onoff(ref_v)
        assert ro[i + 1] - ro[i] == ref_v.shape[1]
        assert np.array_equal(velsum[ro[i]:ro[i + 1]].cpu().numpy().T, ref_v)
        assert np.array_equal(roll[ro[i]:ro[i + 1]].cpu().numpy(), ref_r)
        assert np.array_equal(onoff[ro[i]:ro[i + 1]].cpu().numpy(), ref_o)
    # through a MIDI file and the load_midi drop-in
    p, v, s, e = pieces[0]
    midi.write_midi_notes(str(tmp_path / "2308_y_mixcraft.mid"), p, v, s, e, cc64=pedals[0])
    got_r, got_o = pkg.preprocess.load_midi(str(tmp_path), 2308)
    rp, rv, rs, re_, cc, end_time = midi.read_midi(str(tmp_path / "2308_y_mixcraft.mid"))
    ref_r, ref_o = opp.midi_notes_to_pianoroll(rp, rv, rs, re_, cc64=cc, end_time=end_time)
    assert np.array_equal(got_r, ref_r) and np.array_equal(got_o, ref_o)
    assert got_r.sum() > opp.midi_notes_to_pianoroll(rp, rv, rs, re_)[0].sum()  # the pedal really sustained something


def _random_midi_file(path, seed, n_tracks=3):
    """Multi-track file: pitched instruments on distinct channels with their own pedal and pitch-bend streams, a program
    change mid-track, a drum track, and a channel that only sends control changes."""
    from ml_music_style_transfer_b200 import midi, synth
    rng = np.random.default_rng(seed)
    tracks = []
    for tr in range(n_tracks):
        ch = tr
        p, v, s, e = synth.midi_piece(500 + 10 * seed + tr, seconds=5.0 + tr, notes_per_second=6.0)
        ev = [('note', ch, int(a), int(b), float(c), float(d)) for a, b, c, d in zip(p, v, s, e)]
        for t in np.sort(rng.uniform(0, 6.5 + tr, 10)):
            ev.append(('cc', ch, 64, int(rng.choice([0, 30, 64, 90, 127])), float(t)))
        for t in np.sort(rng.uniform(0, 6.0 + tr, 12)):
            ev.append(('bend', ch, int(rng.choice([0, 0, 1, -1, 700, -700, 2048, 4096, -4096, 5000, -6000, 8191, -8192])), float(t)))
        ev.append(('cc', ch, 7, 100, 0.0))
        if tr == 1:
            ev.append(('program', ch, 40, 2.5))          # notes closed after 2.5 s land in a second instrument
        tracks.append(ev)
    tracks.append([('note', 9, 36, 120, 0.0, 9.5), ('note', 9, 38, 90, 1.0, 1.2)])   # drums: zeros, but 9.5 s wide
    tracks[0].append(('cc', 12, 64, 127, 30.0))                                         # note-less channel: ignored
    midi.write_midi_tracks(path, tracks)


@pytest.mark.parametrize("fs", [172, 250])
def test_midi_full_get_piano_roll(pkg, gpu, tmp_path, fs):
    """PrettyMIDI(file).get_piano_roll(fs) semantics on whole files: per-instrument CC64 sustain, pitch bends
    (float64, NumPy's rounding sequence), drums, instrument sum -- bit exact against the oracle, two files in one batch."""
    from ml_music_style_transfer_b200 import midi
    P = pkg.pianoroll
    files = []
    for i in range(2):
        path = str(tmp_path / f"{2308 + i}_multi_mixcraft.mid")
        _random_midi_file(path, i)
        files.append(midi.read_midi_file(path))
    assert len(files[0].instruments) == 5 and any(i.is_drum for i in files[0].instruments)
    roll, onoff, row_off, vs = P.midi_to_pianoroll(files, fs, want_f64=True)
    ro = row_off.cpu().numpy()
    bent = False
    for i, mf in enumerate(files):
        ref_v = opr.prettymidi_piano_roll(mf.instruments, fs)
        ref_r, ref_o = opr.binarize_and_onoff(ref_v)
        assert ro[i + 1] - ro[i] == ref_v.shape[1] == int(fs * 9.5)
        assert np.array_equal(vs[ro[i]:ro[i + 1]].cpu().numpy().T, ref_v)          # float64 values, bit for bit
        assert np.array_equal(roll[ro[i]:ro[i + 1]].cpu().numpy(), ref_r)
        assert np.array_equal(onoff[ro[i]:ro[i + 1]].cpu().numpy(), ref_o)
        bent |= bool((ref_v != np.floor(ref_v)).any())
        # per-instrument pedals matter: merging all channels into one instrument gives a different roll
        rp, rv, rs, re_, cc, end_time = midi.read_midi(str(tmp_path / f"{2308 + i}_multi_mixcraft.mid"))
        merged = opr.get_piano_roll(rp, rv, rs, re_, fs, end_time=end_time, cc64=cc)
        assert merged.shape == ref_v.shape and not np.array_equal(merged != 0, ref_v != 0)
    assert bent                                                                      # fractional bends were exercised
    # the drop-in reads the same file through the same path (fs = hp.wps = 172)
    got_r, got_o = pkg.preprocess.load_midi(str(tmp_path), 2308)
    ref_r, ref_o = opr.binarize_and_onoff(opr.prettymidi_piano_roll(files[0].instruments, 172))
    assert got_r.dtype == np.float64 and np.array_equal(got_r, ref_r) and np.array_equal(got_o, ref_o)


def test_process_pianoroll_into_chunks_dropin(pkg):
    from ml_music_style_transfer_b200 import synth
    p, v, s, e = synth.midi_piece(3, seconds=14.0)
    roll, onoff = pkg.preprocess.notes_to_pianoroll(p, v, s, e)
    ref_r, ref_o = opp.midi_notes_to_pianoroll(p, v, s, e)
    assert np.array_equal(roll, ref_r) and np.array_equal(onoff, ref_o)
    n = pkg.preprocess.get_num_song_chunks(roll)
    assert n == opp.get_num_song_chunks(ref_r) and n >= 2
    a, b = pkg.preprocess.process_pianoroll_into_chunks(roll, onoff, 1, n)
    ra, rb = opp.process_pianoroll_into_chunks(ref_r, ref_o, 1, n)
    assert a.dtype == np.float64 and np.array_equal(a, ra) and np.array_equal(b, rb)


def test_load_midi_dropin(pkg, tmp_path):
    from ml_music_style_transfer_b200 import midi, synth
    p, v, s, e = synth.midi_piece(4, seconds=8.0)
    midi.write_midi_notes(str(tmp_path / "2240_x_mixcraft.mid"), p, v, s, e)
    roll, onoff = pkg.preprocess.load_midi(str(tmp_path), 2240)
    rp, rv, rs, re_ = midi.read_midi_notes(str(tmp_path / "2240_x_mixcraft.mid"))
    ref_r, ref_o = opp.midi_notes_to_pianoroll(rp, rv, rs, re_)
    assert np.array_equal(roll, ref_r) and np.array_equal(onoff, ref_o)
    with pytest.raises(ValueError):
        pkg.preprocess.load_midi(str(tmp_path), 9999)


# ---- next row 1: librosa.load = WAV decode + mono + kaiser_best resample ------------------------------
@pytest.mark.parametrize("so,sn", [(44100, 22050), (48000, 44100), (22050, 44100), (44100, 44100)])
def test_resample_kaiser_best(pkg, so, sn):
    x = clip(51, 30011, "noise") + clip(52, 30011)
    got = pkg.audio_io.resample(x, so, sn)
    ref = oaudio.resample(x, so, sn)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert_close(got, ref.astype(np.float64))


@pytest.mark.parametrize("so,sn", [(48000, 44100), (22050, 44100), (44100, 16000), (8000, 22050), (44100, 48000), (44101, 44100)])
def test_resample_phase_tables_match_per_output_interpolation(pkg, so, sn, monkeypatch):
    """Rational ratios run resample_phase_kernel (tap weights precomputed per phase); MST_RS_NO_PHASES=1 forces the
    per-output table interpolation.  Same taps, same float32 fma, same summation order: the two agree to rounding of the
    phase position (1e-6), and both match the oracle; awkward lengths exercise the wing limits at both signal ends.
    44101 -> 44100 has 44100 phases (> 2048): it always takes the per-output kernel."""
    for n in (5, 130, 1000, 20011):
        x = (clip(57, max(n, 1025), "noise")[:n] + 0.5).astype(np.float32)
        monkeypatch.delenv("MST_RS_NO_PHASES", raising=False)
        a = pkg.audio_io.resample(x, so, sn)
        monkeypatch.setenv("MST_RS_NO_PHASES", "1")
        b = pkg.audio_io.resample(x, so, sn)
        monkeypatch.delenv("MST_RS_NO_PHASES", raising=False)
        assert a.shape == b.shape == (int(np.ceil(n * sn / so)),)
        scale = max(1.0, float(np.abs(b).max()))
        assert np.abs(a - b).max() <= 2e-6 * scale, (n, np.abs(a - b).max())
        if n <= 1000:
            ref = oaudio.resample(x, so, sn)
            assert np.abs(a - ref).max() <= 1e-5 * scale, (n, np.abs(a - ref).max())


def test_resample_halving_edge_cases_and_mono_mix(pkg, gpu):
    """The decimate-by-2 FIR path (one table phase for every output) at awkward lengths, and librosa.to_mono."""
    for n in (2, 3, 100, 255, 1023, 4097):   # (n = 1 gives zero output samples: resampy itself raises there)
        x = clip(55, max(n, 1025), "noise")[:n] * 3.0
        got = pkg.audio_io.resample(x, 44100, 22050)
        ref = oaudio.resample(x, 44100, 22050)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), (n, np.abs(got - ref).max())
    for ch in (2, 3, 6):
        fr = np.random.default_rng(ch).standard_normal((1000, ch)).astype(np.float32)
        mono = pkg._lib.ops().mono_mix(torch.from_numpy(fr).to(gpu)).cpu().numpy()
        assert np.array_equal(mono, np.mean(fr.T, axis=0))


def test_load_audio_dropin(pkg, tmp_path):
    """preprocess.py:99-115 on a stereo 48 kHz PCM16 file -> mono float32 at hp.sr = 44100."""
    x = np.stack([clip(53, 24000), clip(54, 24000, "noise")], axis=1)
    oaudio.write_wav(str(tmp_path / "2240_prelude_cuba.wav"), x, 48000, bits=16)
    got = pkg.preprocess.load_audio(str(tmp_path), 2240, "cuba")
    ref, sr = oaudio.load(str(tmp_path / "2240_prelude_cuba.wav"), sr=44100)
    assert sr == 44100 and got.shape == ref.shape == (22050,) and got.dtype == np.float32
    assert_close(got, ref.astype(np.float64))
    with pytest.raises(ValueError):
        pkg.preprocess.load_audio(str(tmp_path), 2240, "upright")


def test_process_custom_midi_and_audio_dropin(pkg, tmp_path):
    """inference.py:37-71: device-resident model inputs from a MIDI file and a wav file."""
    from ml_music_style_transfer_b200 import midi, synth
    os_ = __import__("os")
    os_.makedirs(str(tmp_path / "midi"))
    p, v, s, e = synth.midi_piece(9, seconds=3.0)
    midi.write_midi_notes(str(tmp_path / "midi" / "song.mid"), p, v, s, e, cc64=[(0.5, 127), (1.5, 0)])
    x = clip(61, 44100 * 2, "noise")
    oaudio.write_wav(str(tmp_path / "song.wav"), x, 44100, bits=32)
    synth_ = pkg.inference.AudioSynthesizer("ckpt.tar", str(tmp_path), "song.mid", str(tmp_path / "song.wav"))
    roll, onoff, spec = synth_.process_custom_midi_and_audio("song.mid", str(tmp_path / "song.wav"))
    assert roll.is_cuda and roll.dtype == torch.float32 and roll.shape[:2] == (1, 128) and onoff.shape == roll.shape
    assert spec.shape == (1, 1025, 1 + len(x) // 256)
    rp, rv, rs, re_, cc, end_time = midi.read_midi(str(tmp_path / "midi" / "song.mid"))
    ref_r, ref_o = opp.midi_notes_to_pianoroll(rp, rv, rs, re_, cc64=cc, end_time=end_time)
    assert np.array_equal(roll[0].cpu().numpy(), ref_r.T) and np.array_equal(onoff[0].cpu().numpy(), ref_o.T)
    assert_close(spec[0].cpu().numpy(), opp.process_spectrum_from_chunk(x).astype(np.float64))


def test_get_data_end_to_end(pkg, tmp_path):
    """preprocess.py:163-200 on one synthetic song: midi + two style wavs -> shard dataset -> training items."""
    from ml_music_style_transfer_b200 import midi, synth
    from ml_music_style_transfer_b200.dataset import ShardDataset
    seconds = 13.0
    p, v, s, e = synth.midi_piece(7, seconds=seconds)
    midi.write_midi_notes(str(tmp_path / "1749_x_mixcraft.mid"), p, v, s, e)
    audio = {st: clip(60 + i, int(44100 * (seconds + 1)), "noise") for i, st in enumerate(("cuba", "upright"))}
    for st, x in audio.items():
        oaudio.write_wav(str(tmp_path / f"1749_x_{st}.wav"), x, 44100, bits=32)
    mgr = pkg.preprocess.get_data(str(tmp_path), str(tmp_path / "out"), "train", debug=True, piano_scores=[1749],
                                  styles=["cuba", "gentleman", "upright"])  # 'gentleman' is missing -> skipped
    assert sorted(mgr.keys()) == ["onoff", "pianoroll", "spec_cuba", "spec_upright"]
    ds = ShardDataset(str(tmp_path / "out_train"))
    rp, rv, rs, re_ = midi.read_midi_notes(str(tmp_path / "1749_x_mixcraft.mid"))
    ref_r, ref_o = opp.midi_notes_to_pianoroll(rp, rv, rs, re_)
    n = opp.get_num_song_chunks(ref_r)
    assert len(ds) == n and n >= 2
    ra, rb = opp.process_pianoroll_into_chunks(ref_r, ref_o, 1749, n)
    X, Xc, y = ds[1]
    assert np.array_equal(X.numpy(), np.concatenate((ra[1], rb[1]), axis=-1).T.astype(np.float32))
    refs = [opp.process_audio_into_chunks(audio[st], st, 1749, n)[1] for st in audio]
    assert min(rel_l2(y.numpy(), r.astype(np.float64)) for r in refs) <= TOL
    # device-resident dataset: same items (same RNG stream), served from GPU memory without a host hop
    ds_h = ShardDataset(str(tmp_path / "out_train"), seed=7)
    ds_d = ShardDataset(str(tmp_path / "out_train"), seed=7, device="cuda:0")
    assert ds_d.resident and ds_d.pianoroll.is_cuda and ds_d.pianoroll.dtype == torch.int8
    items_h = [ds_h[i] for i in range(len(ds_h))]
    ds_d.__init__(str(tmp_path / "out_train"), seed=7, device="cuda:0")   # restart the RNG stream
    for i, (Xh, Ch, yh) in enumerate(items_h):
        Xd, Cd, yd = ds_d[i]
        assert Xd.is_cuda and Xd.shape == (256, 860) and Xd.dtype == torch.float32
        assert torch.equal(Xd.cpu(), Xh) and torch.equal(Cd.cpu(), Ch) and torch.equal(yd.cpu(), yh)
    # running get_data again truncates (h5py.File(..., 'w') semantics) instead of duplicating every song
    mgr2 = pkg.preprocess.get_data(str(tmp_path), str(tmp_path / "out"), "train", piano_scores=[1749], styles=["cuba"])
    assert mgr2.n_rows("pianoroll") == n and sorted(mgr2.keys()) == ["onoff", "pianoroll", "spec_cuba"]


# ---- next row 4: mel inversion (librosa.feature.inverse.mel_to_audio, tests/test_griffinlim.py:24) -----------------------
def test_mel_to_stft_and_mel_to_audio(pkg, gpu):
    """mel_to_stft: same start point as librosa.util.nnls, every frame's NNLS problem solved to convergence.  The
    oracle restates librosa (L-BFGS-B from the clipped least-squares point); L-BFGS-B stops early on its scaled
    projected-gradient test, so the two are not the same point of the (non-unique) solution set.  Checked: start point,
    feasibility, residual no worse than librosa's, closeness to librosa's answer, and the audio's consistency with M."""
    F = pkg.features
    sr, hop = 22050, 512
    y = clip(91, 30000)
    M = omel.melspectrogram(y, sr, 2048, hop)                      # (128, 59) float32
    W = omel.mel_filterbank(sr, 2048, 128)
    Wd = W.astype(np.float64)
    res = lambda X: float(np.linalg.norm(Wd @ X.astype(np.float64) - M) / np.linalg.norm(M))
    # start point: max(pinv(A) M, 0) == clipped least squares
    X0 = F.mel_to_stft(M, sr=sr, power=1.0, max_iter=0)
    ref0 = np.clip(np.linalg.lstsq(Wd, M.astype(np.float64), rcond=None)[0], 0, None)
    assert X0.shape == ref0.shape == (1025, M.shape[1])
    assert rel_l2(X0, ref0) < 1e-4, rel_l2(X0, ref0)
    # converged solution vs librosa's
    S_ref = omel.mel_to_stft(M, sr=sr, n_fft=2048, power=2.0)
    S = F.mel_to_stft(M, sr=sr, n_fft=2048, power=2.0)
    assert S.shape == S_ref.shape and S.dtype == np.float32 and (S >= 0).all() and np.isfinite(S).all()
    r_gpu, r_ref, r0 = res(S.astype(np.float64) ** 2), res(S_ref.astype(np.float64) ** 2), res(ref0)
    assert r_gpu <= r_ref + 1e-6 and r_gpu < 0.2 * r0, (r_gpu, r_ref, r0)
    # the NNLS problem is under-determined (1025 unknowns, 128 equations): librosa's early-stopped L-BFGS-B point and the
    # converged one are different members of (the neighbourhood of) the solution set -- measured distance ~0.35 in
    # magnitude; only a sanity bound is asserted, the binding checks are the start point and the residual above
    assert rel_l2(S, S_ref.astype(np.float64)) < 0.6, rel_l2(S, S_ref.astype(np.float64))
    # 80 mels, frame-major batched op on a ragged batch == per-clip calls
    frames = [20, 59, 7]
    Ms = [omel.melspectrogram(clip(92 + i, hop * (T - 1)), sr, 2048, hop, 80) for i, T in enumerate(frames)]
    plan = F.MelInversePlan.get(sr, n_mels=80, device=gpu)
    gb = F.ClipBatch.from_frames(frames, hop, device=gpu)
    flat = torch.from_numpy(np.concatenate([m.ravel() for m in Ms])).to(gpu)
    Sb = F.mel_to_stft_batch(flat, gb, plan, 2.0, F.BIN_MAJOR).cpu().numpy()
    o = 0
    for m, T in zip(Ms, frames):
        one = F.mel_to_stft(m, sr=sr, power=2.0)
        assert np.array_equal(Sb[o:o + T].T, one)
        o += T
    # mel_to_audio: Griffin-Lim on the inverted magnitudes; the audio's mel spectrogram must be as close to M as the oracle's
    u = ogl.random_phase((1025, M.shape[1]), 4)
    a_ref = omel.mel_to_audio(M, sr=sr, n_fft=2048, hop_length=hop, n_iter=32, init_phase=u)
    a_gpu = F.mel_to_audio(M, sr=sr, n_fft=2048, hop_length=hop, n_iter=32, init_phase=u)
    assert a_gpu.shape == a_ref.shape == (hop * (M.shape[1] - 1),)
    mel_sc = lambda a: float(np.linalg.norm(omel.melspectrogram(a, sr, 2048, hop).astype(np.float64) - M) / np.linalg.norm(M))
    assert mel_sc(a_gpu) <= mel_sc(a_ref) + 2e-2, (mel_sc(a_gpu), mel_sc(a_ref))
    # and Griffin-Lim itself is the same op as before: SC against the magnitudes it was given, gpu vs oracle loop on the SAME S
    ref_same = ogl.griffinlim(S, 32, hop, init_phase=u)
    got_same = F.griffinlim(S, n_iter=32, hop_length=hop, init_phase=u)
    assert abs(_sc(S, ref_same, hop) - _sc(S, got_same, hop)) <= 1e-3


# ---- P4: Griffin-Lim ------------------------------------------------------------------------
def _sc(S, y, hop):
    return ogl.spectral_convergence(S, y, hop)


@pytest.mark.parametrize("hop,n_iter,mom", [(512, 32, 0.99), (256, 32, 0.99), (512, 16, 0.0)])
def test_griffinlim_spectral_convergence(pkg, hop, n_iter, mom):
    y = clip(31, 44100)
    S = np.abs(ostft.stft(y, 2048, hop)).astype(np.float32)
    u = ogl.random_phase(S.shape, 0)
    ref = ogl.griffinlim(S, n_iter, hop, momentum=mom, init_phase=u)
    got = pkg.features.griffinlim(S, n_iter=n_iter, hop_length=hop, momentum=mom, init_phase=u)
    assert got.shape == ref.shape == (hop * (S.shape[1] - 1),) and got.dtype == np.float32
    sc_ref, sc_got = _sc(S, ref, hop), _sc(S, got, hop)
    assert abs(sc_ref - sc_got) <= 1e-3, (sc_ref, sc_got)
    assert sc_got < 0.5


def test_griffinlim_reference_configuration_300_iterations(pkg):
    """The reference's own call: n_iter=300, hop 256 (inference.py:105, tests/test_griffinlim.py:23)."""
    y = clip(35, 22050)
    S = np.abs(ostft.stft(y, 2048, 256)).astype(np.float32)
    u = ogl.random_phase(S.shape, 3)
    ref = ogl.griffinlim(S, 300, 256, init_phase=u)
    got = pkg.features.griffinlim(S, n_iter=300, hop_length=256, init_phase=u)
    sc_ref, sc_got = _sc(S, ref, 256), _sc(S, got, 256)
    assert abs(sc_ref - sc_got) <= 1e-3, (sc_ref, sc_got)
    # random_state=int reproduces librosa's RandomState(seed).rand(*S.shape) phase draw
    got2 = pkg.features.griffinlim(S, n_iter=8, hop_length=256, random_state=3)
    ref2 = ogl.griffinlim(S, 8, 256, init_phase=u)
    assert rel_l2(got2, ref2.astype(np.float64)) < 5e-3


def test_griffinlim_ragged_batch_matches_per_clip_oracle(pkg, gpu):
    """Ragged batch in one launch sequence, bin-major log1p-power input (what the model emits), shared phases."""
    F = pkg.features
    frames = [60, 173, 45]
    hop = 512
    S_list, u_list = [], []
    for i, T in enumerate(frames):
        yy = clip(40 + i, hop * (T - 1))
        S_list.append(np.abs(ostft.stft(yy, 2048, hop)).astype(np.float32))
        u_list.append(ogl.random_phase(S_list[-1].shape, 10 + i).astype(np.float32))
        assert S_list[-1].shape == (1025, T)
    logp = np.concatenate([np.log1p(S.astype(np.float64) ** 2).astype(np.float32).ravel() for S in S_list])
    ph = np.concatenate([u.ravel() for u in u_list])
    gb = F.ClipBatch.from_frames(frames, hop, device=gpu)
    out = F.griffinlim_batch(torch.from_numpy(logp).to(gpu), gb, n_iter=16, init_phase=torch.from_numpy(ph).to(gpu),
                             layout=F.BIN_MAJOR, is_log1p_power=True).cpu().numpy()
    o = 0
    for S, u, T in zip(S_list, u_list, frames):
        L = hop * (T - 1)
        mag = ogl.logpower_to_magnitude(np.log1p(S.astype(np.float64) ** 2).astype(np.float32))
        ref = ogl.griffinlim(mag, 16, hop, init_phase=u)
        assert abs(_sc(mag, ref, hop) - _sc(mag, out[o:o + L], hop)) <= 1e-3
        o += L
    assert o == out.shape[0]


@pytest.mark.parametrize("hop", [128, 300, 441, 1024])
def test_other_hops(pkg, hop):
    """Hops the reference never uses still follow librosa: odd hops take the unaligned frame loader, hop 128 / 1024 the
    R = 16 / R = 2 overlap-add, non-power-of-two hops the generic overlap-add and the per-frame envelope path."""
    y = clip(36, 20000 + hop)
    ref = ostft.stft(y, 2048, hop)
    got = pkg.features.stft(y, hop_length=hop)
    assert_close(got, ref)
    S = np.abs(ref).astype(np.float32)
    u = ogl.random_phase(S.shape, 2)
    for n_iter in (0, 3):
        w_ref = ogl.griffinlim(S, n_iter, hop, init_phase=u)
        w_got = pkg.features.griffinlim(S, n_iter=n_iter, hop_length=hop, init_phase=u)
        assert w_got.shape == w_ref.shape
        assert rel_l2(w_got, w_ref.astype(np.float64)) < 5e-4, (hop, n_iter, rel_l2(w_got, w_ref.astype(np.float64)))


@pytest.mark.parametrize("win_length,hop", [(1024, 256), (1500, 375), (2047, 512), (400, None)])
def test_win_length_shorter_than_n_fft(pkg, win_length, hop):
    """librosa.stft / griffinlim with win_length < n_fft: periodic Hann of that length, centre-padded to 2048 -- analysis
    window, synthesis window and window-sum-square envelope all follow (the reference passes win_length = n_fft)."""
    y = clip(37, 24000)
    ref = ostft.stft(y, 2048, hop, win_length=win_length)
    got = pkg.features.stft(y, hop_length=hop, win_length=win_length)
    assert got.shape == ref.shape
    assert_close(got, ref)
    h = win_length // 4 if hop is None else hop
    S = np.abs(ref).astype(np.float32)
    u = ogl.random_phase(S.shape, 6)
    for n_iter in (0, 2):
        w_ref = ogl.griffinlim(S, n_iter, h, win_length=win_length, init_phase=u)
        w_got = pkg.features.griffinlim(S, n_iter=n_iter, hop_length=hop, win_length=win_length, init_phase=u)
        assert w_got.shape == w_ref.shape
        assert rel_l2(w_got, w_ref.astype(np.float64)) < 5e-4, (win_length, n_iter, rel_l2(w_got, w_ref.astype(np.float64)))
    w_ref = ogl.griffinlim(S, 16, h, win_length=win_length, init_phase=u)
    w_got = pkg.features.griffinlim(S, n_iter=16, hop_length=hop, win_length=win_length, init_phase=u)
    sc = lambda w: ogl.spectral_convergence(S, w, h, win_length=win_length)
    assert abs(sc(w_ref) - sc(w_got)) <= 1e-3, (sc(w_ref), sc(w_got))
    with pytest.raises(ValueError):
        pkg.features.stft(y, hop_length=256, win_length=4096)


@pytest.mark.parametrize("hop", [128, 256, 512, 1024])
def test_griffinlim_all_tile_remainders(pkg, gpu, hop):
    """Every frame count from the shortest legal clip up to 3+ tiles in ONE ragged batch: exercises first / last / only
    tile flags, partial tiles, clips with no interior frame, and the shared / exclusive overlap-add blocks."""
    F = pkg.features
    t_min = 1024 // hop + 2                      # hop * (T - 1) must exceed n_fft / 2 (reflect padding)
    frames = list(range(t_min, t_min + 27))
    rng = np.random.default_rng(hop)
    S_list = [np.abs(rng.standard_normal((1025, T))).astype(np.float32) for T in frames]
    u_list = [rng.random((1025, T)).astype(np.float32) for T in frames]
    gb = F.ClipBatch.from_frames(frames, hop, device=gpu)
    S = torch.from_numpy(np.concatenate([s.ravel() for s in S_list])).to(gpu)
    ph = torch.from_numpy(np.concatenate([u.ravel() for u in u_list])).to(gpu)
    for n_iter in (0, 2):
        out = F.griffinlim_batch(S, gb, n_iter=n_iter, init_phase=ph, layout=F.BIN_MAJOR).cpu().numpy()
        o = 0
        for Sm, u, T in zip(S_list, u_list, frames):
            L = hop * (T - 1)
            ref = ogl.griffinlim(Sm, n_iter, hop, init_phase=u)
            err = rel_l2(out[o:o + L], ref.astype(np.float64))
            assert err < 5e-4, (hop, T, n_iter, err)
            o += L
        assert o == out.shape[0]
    # the same clips replicated past 2 x n_SM tiles take the multi-launch path (one kernel per iteration) instead of the
    # cooperative persistent kernel: both must give the same waveforms, replica after replica
    reps = 14
    gb2 = F.ClipBatch.from_frames(frames * reps, hop, device=gpu)
    out2 = F.griffinlim_batch(S.repeat(reps), gb2, n_iter=2, init_phase=ph.repeat(reps), layout=F.BIN_MAJOR).cpu().numpy()
    assert out2.shape[0] == reps * out.shape[0]
    for r in (0, reps // 2, reps - 1):
        seg = out2[r * out.shape[0]:(r + 1) * out.shape[0]]
        assert rel_l2(seg, out.astype(np.float64)) < 1e-5, (hop, r)


def test_spectral_convergence_fused(pkg, gpu):
    """mst_spectral_convergence_f32 (fused STFT epilogue + per-clip reduction) == the oracle's Frobenius ratio."""
    F = pkg.features
    frames, hop = [60, 173, 45], 512
    ys, Ss = [], []
    for i, T in enumerate(frames):
        yy = clip(70 + i, hop * (T - 1))
        S = np.abs(ostft.stft(clip(80 + i, hop * (T - 1), "noise"), 2048, hop)).astype(np.float32)  # unrelated target
        ys.append(yy); Ss.append(S)
    y = torch.from_numpy(np.concatenate(ys)).to(gpu)
    lens = [len(v) for v in ys]
    b = F.ClipBatch.from_clips(np.concatenate([[0], np.cumsum(lens)[:-1]]), lens, hop, device=gpu)
    S_bm = torch.from_numpy(np.concatenate([S.ravel() for S in Ss])).to(gpu)
    S_fm = torch.from_numpy(np.concatenate([S.T.ravel() for S in Ss])).to(gpu)
    got_bm = F.spectral_convergence_batch(y, b, S_bm, F.BIN_MAJOR).cpu().numpy()
    got_fm = F.spectral_convergence_batch(y, b, S_fm, F.FRAME_MAJOR).cpu().numpy()
    for i in range(3):
        ref = ogl.spectral_convergence(Ss[i], ys[i], hop)
        assert abs(got_bm[i] - ref) < 1e-5 * ref and abs(got_fm[i] - ref) < 1e-5 * ref
    assert abs(F.spectral_convergence(Ss[1], ys[1], hop) - ogl.spectral_convergence(Ss[1], ys[1], hop)) < 1e-5


def test_griffinlim_early_iterations_match_waveform(pkg):
    """Before the chaotic phase dynamics amplify float32 rounding, the waveform itself must agree."""
    y = clip(32, 30000)
    S = np.abs(ostft.stft(y, 2048, 512)).astype(np.float32)
    u = ogl.random_phase(S.shape, 1)
    for n_iter in (0, 1, 2):
        ref = ogl.griffinlim(S, n_iter, 512, init_phase=u)
        got = pkg.features.griffinlim(S, n_iter=n_iter, hop_length=512, init_phase=u)
        assert rel_l2(got, ref.astype(np.float64)) < 2e-4, (n_iter, rel_l2(got, ref.astype(np.float64)))
    ref = ogl.griffinlim(S, 2, 512, init_phase=None)
    got = pkg.features.griffinlim(S, n_iter=2, hop_length=512, init=None)
    assert rel_l2(got, ref.astype(np.float64)) < 2e-4


def test_audiosynthesizer_griffinlim_dropin(pkg):
    """inference.py:105-110: log1p-power in, sqrt(expm1(clip)) fused, hop 256."""
    y = clip(33, 30000)
    spec = opp.process_spectrum_from_chunk(y)  # (1025, T) log1p power, what the model emits
    mag = ogl.logpower_to_magnitude(spec)
    u = ogl.random_phase(spec.shape, 5)
    synth_ = pkg.inference.AudioSynthesizer(None, None, None, None)
    got = synth_.griffinlim(spec, "a", n_iter=24, init_phase=u)
    ref = ogl.griffinlim(mag, 24, 256, init_phase=u)
    assert got.shape == ref.shape
    assert abs(_sc(mag, ref, 256) - _sc(mag, got, 256)) <= 1e-3
    # device RNG path: different phase draw, same convergence behaviour
    got2 = synth_.griffinlim(spec, "b", n_iter=24)
    assert _sc(mag, got2, 256) < 1.5 * _sc(mag, ref, 256) + 0.05


def test_griffinlim_batch_full_size_round_trip(pkg, gpu):
    """Size-independent property at scale: with the TRUE phase as initial phase and n_iter=0 the synthesis launch is
    istft(stft(y)) and must return y (COLA); a ragged batch exercises the packed layout."""
    F = pkg.features
    frames = [173, 200, 57, 173]
    hop = 512
    lens = [hop * (t - 1) for t in frames]
    y = clip(34, sum(lens), "noise")
    a = torch.from_numpy(y).to(gpu)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    b = F.ClipBatch.from_clips(offs, lens, hop, device=gpu)
    D = F.stft_batch(a, b, "complex")  # (total_frames, 1025) complex64, frame-major
    S = D.abs().contiguous()
    ph = (torch.angle(D) / (2 * np.pi)) % 1.0
    gb = F.ClipBatch.from_frames(frames, hop, device=gpu)
    out = F.griffinlim_batch(S, gb, n_iter=0, init_phase=ph.float().contiguous(), layout=F.FRAME_MAJOR).cpu().numpy()
    o = 0
    for l in lens:
        seg, ref = out[o:o + l], y[o:o + l]
        assert np.abs(seg - ref)[1024:-1024].max() < 5e-5
        o += l
    # and a few iterations from there must keep the consistent spectrogram (fixed point of the projection)
    out2 = F.griffinlim_batch(S, gb, n_iter=3, init_phase=ph.float().contiguous(), layout=F.FRAME_MAJOR).cpu().numpy()
    assert np.abs(out2 - y)[2048:l - 2048].max() < 5e-3


# ---- size-independent properties at benchmark scale (C4 shapes) ------------------------------------------------------
def test_full_size_properties(pkg, gpu):
    """2048 clips x 4 s (an eighth of C4 per GPU): properties that need no oracle run at that size."""
    import bench
    F, PR = pkg.features, pkg.pianoroll
    n = 2048
    audio = bench.make_audio_device(n, gpu, 5)
    batch = F.ClipBatch.uniform(n, bench.CLIP_LEN, bench.HOP, device=gpu)
    plan = F.MelPlan.get(bench.SR, device=gpu)
    # (1) power-mel is quadratic in the signal: mel(2x) == 4 mel(x) (exact powers of two survive every rounding)
    m1 = F.melspectrogram_batch(audio, batch, plan, log1p=False, layout=F.BIN_MAJOR)
    m2 = F.melspectrogram_batch(audio * 2.0, batch, plan, log1p=False, layout=F.BIN_MAJOR)
    assert torch.equal(m2, 4.0 * m1)
    assert torch.isfinite(m1).all() and (m1 >= 0).all()
    # (2) layouts agree: bin-major == transpose of frame-major, clip by clip
    fm = F.melspectrogram_batch(audio, batch, plan, log1p=False, layout=F.FRAME_MAJOR).view(n, bench.T_FRAMES, 128)
    assert torch.equal(m1.view(n, 128, bench.T_FRAMES), fm.transpose(1, 2))
    # (3) Parseval per frame on the log-power path: sum_k c_k |X_k|^2 == N * sum (w x)^2 for interior frames
    P = F.stft_batch(audio[:64 * bench.CLIP_LEN], F.ClipBatch.uniform(64, bench.CLIP_LEN, bench.HOP, device=gpu), "power",
                     F.FRAME_MAJOR).view(64, bench.T_FRAMES, 1025).double()
    wts = torch.full((1025,), 2.0, dtype=torch.float64, device=gpu)
    wts[0] = wts[1024] = 1.0
    lhs = (P * wts).sum(-1)[:, 10]
    win = torch.hann_window(2048, periodic=True, dtype=torch.float64, device=gpu)
    x = audio[:64 * bench.CLIP_LEN].view(64, -1)[:, 10 * 512 - 1024:10 * 512 + 1024].double()
    rhs = 2048.0 * ((x * win) ** 2).sum(-1)
    assert torch.allclose(lhs, rhs, rtol=1e-5)
    # (4) piano roll at scale: every audio-rate row is the hold-replication of its frame-rate row, checked by an exact
    # integer identity (sum over samples == sum over columns of value x samples-per-column)
    notes = PR.NoteBatch(*bench.make_notes(256, 99), device=gpu)
    roll, onoff, row_off, _ = PR.rasterize(notes, bench.ROLL_FS)
    up, so = PR.upsample(roll, row_off, bench.CLIP_LEN, bench.ROLL_FS, bench.SR, 21, 88, torch.int8)
    up = up.view(256, 88, bench.CLIP_LEN)
    cols = (torch.arange(bench.CLIP_LEN, device=gpu) * bench.ROLL_FS) // bench.SR
    ro = row_off.cpu().numpy()
    for i in (0, 17, 255):
        T = int(ro[i + 1] - ro[i])
        per_col = torch.bincount(cols[cols < T], minlength=T).to(torch.int64)
        want = (roll[ro[i]:ro[i + 1], 21:109].to(torch.int64) * per_col[:, None]).sum(0)
        assert torch.equal(up[i].to(torch.int64).sum(-1), want)
        # on/off is the time derivative of the roll: integrating it gives the roll back (roll -> events -> roll round trip,
        # the property behind the reference's utils/pretty_midi_roll_to_midi.py)
        assert torch.equal(torch.cumsum(onoff[ro[i]:ro[i + 1]].to(torch.int32), 0), roll[ro[i]:ro[i + 1]].to(torch.int32))
    # (5) Griffin-Lim at scale is a fixed point on consistent spectrograms: true phase in, 2 iterations, signal back
    nb = 256
    gb = F.ClipBatch.from_frames([bench.T_FRAMES] * nb, bench.HOP, device=gpu)
    b2 = F.ClipBatch.uniform(nb, bench.HOP * (bench.T_FRAMES - 1), bench.HOP, clip_stride=bench.CLIP_LEN, device=gpu)
    D = F.stft_batch(audio, b2, "complex")
    ph = ((torch.angle(D) / (2 * np.pi)) % 1.0).float().contiguous()
    y = F.griffinlim_batch(D.abs().contiguous(), gb, n_iter=2, init_phase=ph, layout=F.FRAME_MAJOR).view(nb, -1)
    ref = audio.view(n, -1)[:nb, :y.shape[1]]
    assert (y - ref)[:, 2048:-2048].abs().max() < 2e-3


def test_host_pipeline_matches_direct_calls(pkg, gpu):
    """pipeline.HostPipeline (chunks over several streams, pinned host buffers) == the direct batched ops."""
    import bench
    from ml_music_style_transfer_b200.pipeline import HostPipeline
    F, PR = pkg.features, pkg.pianoroll
    n, clip_len = 600, 22050  # 1 s clips -> two chunks
    audio = bench.make_audio_device(152, gpu, 9)[:n * clip_len].contiguous()
    batch = F.ClipBatch.uniform(n, clip_len, 512, device=gpu)
    T = batch.total_frames // n
    S = F.stft_batch(audio, batch, "magnitude", F.FRAME_MAJOR)
    from ml_music_style_transfer_b200 import synth
    pieces = [synth.midi_piece(i, seconds=1.0) for i in range(n)]
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(p[0]) for p in pieces], out=offs[1:])
    notes = tuple(np.concatenate([p[j] for p in pieces]) for j in range(4)) + (offs,)
    h_audio, h_S = audio.cpu().pin_memory(), S.cpu().pin_memory()
    pipe = HostPipeline(n, clip_len, n_chunks=4, n_streams=3, planes_to_host=False, device=gpu)
    assert len(pipe.chunks) == 2
    pipe.run(h_audio, h_S, notes)
    plan = F.MelPlan.get(22050, device=gpu)
    mel = F.melspectrogram_batch(audio, batch, plan, log1p=True, layout=F.BIN_MAJOR).cpu()
    assert torch.equal(pipe.h_mel, mel)
    # the pipeline rasterises every piece over the clip's duration: piece i owns rows [i * rows_per_clip, (i + 1) * rows_per_clip)
    nb = PR.NoteBatch(*notes, device=gpu, end_times=[clip_len / 22050.0] * n)
    roll, onoff, row_off, _ = PR.rasterize(nb, 250)
    assert pipe.rows_per_clip == 250 and roll.shape[0] == n * pipe.rows_per_clip
    assert torch.equal(pipe.h_roll, roll.cpu()) and torch.equal(pipe.h_onoff, onoff.cpu())
    for i in (0, 299, 300, 599):  # both sides of the chunk boundary, against the oracle
        p_, v_, s_, e_ = pieces[i]
        ref_r, ref_o = opr.binarize_and_onoff(opr.get_piano_roll(p_, v_, s_, e_, 250, end_time=1.0))
        blk = slice(i * 250, (i + 1) * 250)
        assert np.array_equal(pipe.h_roll[blk].numpy(), ref_r) and np.array_equal(pipe.h_onoff[blk].numpy(), ref_o)
    # a second run with different notes re-stages them (nothing is cached between runs)
    notes2 = tuple(np.concatenate([p[j] for p in pieces[::-1]]) for j in range(4)) + \
        (np.concatenate([[0], np.cumsum([len(p[0]) for p in pieces[::-1]])]).astype(np.int64),)
    pipe.run(h_audio, h_S, notes2)
    p_, v_, s_, e_ = pieces[n - 1]
    assert np.array_equal(pipe.h_roll[:250].numpy(), opr.binarize_and_onoff(opr.get_piano_roll(p_, v_, s_, e_, 250, end_time=1.0))[0])
    pipe.run(h_audio, h_S, notes)
    # Griffin-Lim: the random phase stream differs per chunk, so compare convergence, clip by clip
    L = 512 * (T - 1)
    for i in (0, 450):
        w = pipe.h_wave[i * L:(i + 1) * L].numpy()
        Si = S.view(n, T, 1025)[i].t().cpu().numpy()
        assert ogl.spectral_convergence(Si, w, 512) < 0.6
    h2d, d2h = pipe.bytes_per_run()
    assert h2d > h_audio.numel() * 4 and d2h > pipe.h_mel.numel() * 4


def test_device_pipeline_matches_direct_calls(pkg, gpu):
    """pipeline.DevicePipeline (device-resident inputs, chunks over several streams, results collected on the device)
    == the direct batched ops: log-mel and rolls bit for bit, planes through the consumer hook, Griffin-Lim per clip."""
    import bench
    from ml_music_style_transfer_b200 import synth
    from ml_music_style_transfer_b200.pipeline import DevicePipeline
    F, PR = pkg.features, pkg.pianoroll
    n, clip_len = 600, 22050
    audio = bench.make_audio_device(152, gpu, 19)[:n * clip_len].contiguous()
    batch = F.ClipBatch.uniform(n, clip_len, 512, device=gpu)
    T = batch.total_frames // n
    S = F.stft_batch(audio, batch, "magnitude", F.FRAME_MAJOR)
    pieces = [synth.midi_piece(1000 + i, seconds=1.0) for i in range(n)]
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(p[0]) for p in pieces], out=offs[1:])
    notes = tuple(np.concatenate([p[j] for p in pieces]) for j in range(4)) + (offs,)
    seen = {}

    def consumer(c0, c1, up_roll, up_onoff):
        if c0 <= 301 < c1:
            k = (301 - c0) * 88 * clip_len
            seen["roll"] = up_roll[k:k + 88 * clip_len].clone()
            seen["onoff"] = up_onoff[k:k + 88 * clip_len].clone()
    pipe = DevicePipeline(n, clip_len, notes, n_chunks=4, n_streams=3, collect=True, plane_consumer=consumer, device=gpu)
    assert len(pipe.chunks) == 2
    pipe.run(audio, S)
    torch.cuda.synchronize()
    plan = F.MelPlan.get(22050, device=gpu)
    assert torch.equal(pipe.mel, F.melspectrogram_batch(audio, batch, plan, log1p=True, layout=F.BIN_MAJOR))
    nb = PR.NoteBatch(*notes, device=gpu, end_times=[1.0] * n)
    roll, onoff, row_off, _ = PR.rasterize(nb, 250)
    assert torch.equal(pipe.roll, roll) and torch.equal(pipe.onoff, onoff)
    p_, v_, s_, e_ = pieces[301]
    ref_r, ref_o = opr.binarize_and_onoff(opr.get_piano_roll(p_, v_, s_, e_, 250, end_time=1.0))
    assert np.array_equal(seen["roll"].view(88, clip_len).cpu().numpy(), opr.upsample_to_audio_rate(ref_r, 250, 22050, clip_len, 21, 88, np.int8))
    assert np.array_equal(seen["onoff"].view(88, clip_len).cpu().numpy(), opr.upsample_to_audio_rate(ref_o, 250, 22050, clip_len, 21, 88, np.int8))
    L = 512 * (T - 1)
    for i in (0, 299, 300, 599):
        w = pipe.wave[i * L:(i + 1) * L].cpu().numpy()
        Si = S.view(n, T, 1025)[i].t().cpu().numpy()
        assert np.isfinite(w).all() and ogl.spectral_convergence(Si, w, 512) < 0.6
    # a second run reproduces the first bit for bit (same seed, same schedule-independent kernels)
    w1 = pipe.wave.clone()
    pipe.run(audio, S)
    torch.cuda.synchronize()
    assert torch.equal(pipe.wave, w1)
