"""CPU suite (-m "not gpu"): pins the oracle against independent implementations (golden fixtures made by
tests/golden/make_golden.py) and closed-form known answers; checks host logic and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import audio as oaudio, griffinlim as ogl, mel as omel, pianoroll as opr, preprocess as opp, stft as ostft

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


# ---- STFT / iSTFT ------------------------------------------------------------------------------
@pytest.mark.parametrize("hop", [256, 512])
@pytest.mark.parametrize("pad", ["reflect", "constant"])
def test_stft_matches_torch_golden(hop, pad):
    g = np.load(os.path.join(GOLD, "stft_torch.npz"))
    D = ostft.stft(g["y"], 2048, hop, pad_mode=pad)
    ref = g[f"stft_{hop}_{pad}"]
    assert D.shape == ref.shape == (1025, 1 + 6000 // hop)
    assert D.dtype == np.complex64 and D.flags.f_contiguous
    assert relerr(D, ref) < 2e-7


@pytest.mark.parametrize("hop", [256, 512])
def test_istft_matches_torch_golden(hop):
    g = np.load(os.path.join(GOLD, "stft_torch.npz"))
    y = ostft.istft(g[f"stft_{hop}_reflect"], hop)
    ref = g[f"istft_{hop}"]
    assert y.shape == ref.shape == (hop * (6000 // hop),)
    assert np.abs(y - ref).max() < 2e-6


def test_stft_known_answers():
    n = np.arange(8192)
    # cosine on a bin centre: |X[k0]| = N/4 * 2 = 512 (Hann sum = N/2, half per sideband), neighbours half of that
    k0 = 100
    D = ostft.stft(np.cos(2 * np.pi * k0 * n / 2048).astype(np.float32), 2048, 512, out_dtype=np.complex128)
    mid = D[:, 8]
    assert abs(abs(mid[k0]) - 512.0) < 1e-3
    assert abs(abs(mid[k0 - 1]) - 256.0) < 1e-3 and abs(abs(mid[k0 + 1]) - 256.0) < 1e-3
    assert np.abs(mid[k0 + 3:k0 + 50]).max() < 1e-3
    # DC: only bins 0 and 1
    D = ostft.stft(np.ones(8192, dtype=np.float32), 2048, 512, out_dtype=np.complex128)[:, 8]
    assert abs(D[0].real - 1024.0) < 1e-9 and abs(abs(D[1]) - 512.0) < 1e-9 and np.abs(D[2:]).max() < 1e-9
    # unit impulse at frame position n0: flat magnitude w[n0]
    x = np.zeros(8192, dtype=np.float32)
    x[4096 + 300] = 1.0
    D = ostft.stft(x, 2048, 512, out_dtype=np.complex128)[:, 8]  # frame 8 starts at padded 4096 -> sample 3072
    w = ostft.hann_window(2048)[4096 + 300 - 3072]
    assert np.allclose(np.abs(D), w, atol=1e-12)


def test_cola_round_trip():
    y = (0.1 * np.random.default_rng(3).standard_normal(20000)).astype(np.float32)
    for hop in (256, 512):
        r = ostft.istft(ostft.stft(y, 2048, hop), hop)
        assert np.abs(r - y[:len(r)])[1024:-1024].max() < 1e-6


def test_chunk_geometry():
    hp = opp.hyperparams()
    assert hp.wps == 172 and hp.spc * hp.wps == 860
    n = (hp.spc * hp.wps - 1) * hp.ws
    assert n == 219904 and ostft.frame_count(n, hp.ws) == 860 and hp.ws * hp.stride == 131072
    audio = (0.1 * np.random.default_rng(0).standard_normal(131072 + 219904)).astype(np.float32)
    out = opp.process_audio_into_chunks(audio, "s", 1, 2)
    assert out.shape == (2, 1025, 860) and out.dtype == np.float32


# ---- mel ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("sr,nnz", [(22050, 2018), (44100, 2014)])
def test_mel_filterbank_matches_torchaudio_golden(sr, nnz):
    ref = np.load(os.path.join(GOLD, "mel_torchaudio.npz"))[f"fb_{sr}"]
    W = omel.mel_filterbank(sr, 2048, 128)
    assert W.shape == (128, 1025) and W.dtype == np.float32
    assert int((W != 0).sum()) == nnz
    assert (W >= 0).all()
    assert np.abs(W - ref).max() / np.abs(ref).max() < 1e-5
    # banded: non-zeros of every row are contiguous
    for row in W:
        idx = np.nonzero(row)[0]
        assert len(idx) >= 2 and idx[-1] - idx[0] + 1 == len(idx)


def test_mel_inversion_oracle():
    """oracle.mel.nnls / mel_to_stft restate librosa.util.nnls + feature.inverse.mel_to_stft: L-BFGS-B from the clipped
    least-squares point -- feasible, objective not above the start point's, exact on a consistent 1-D problem."""
    from ml_music_style_transfer_b200 import synth
    y = synth.piano_clip(5, 0.5, 22050)
    M = omel.melspectrogram(y, 22050, 2048, 512)
    W = omel.mel_filterbank(22050, 2048, 128)
    X0 = np.clip(np.linalg.lstsq(W, M, rcond=None)[0], 0, None)
    S = omel.mel_to_stft(M, 22050, 2048, power=2.0)
    assert S.shape == (1025, M.shape[1]) and S.dtype == np.float32 and (S >= 0).all()
    r = lambda X: np.linalg.norm(W.astype(np.float64) @ X.astype(np.float64) - M)
    assert r(S.astype(np.float64) ** 2) < r(X0)
    x_true = np.abs(np.random.default_rng(0).standard_normal(1025)).astype(np.float32)
    x = omel.nnls(W, W @ x_true)                               # 1-D right-hand side -> scipy.optimize.nnls (active set)
    assert (x >= 0).all() and np.linalg.norm(W @ x - W @ x_true) < 1e-4 * np.linalg.norm(W @ x_true)


def test_c_abi_mel_filterbank_is_bit_exact_with_oracle(built_libs):
    lib = ctypes.CDLL(built_libs[0])
    for sr in (22050, 44100):
        W = np.zeros((128, 1025), dtype=np.float32)
        rc = lib.mst_mel_filterbank_f32(sr, 2048, 128, ctypes.c_double(0.0), ctypes.c_double(0.0),
                                        W.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        assert np.array_equal(W, omel.mel_filterbank(sr, 2048, 128))


# ---- piano roll --------------------------------------------------------------------------------
def test_pianoroll_known_answer():
    roll = opr.get_piano_roll([60, 60, 64, 21], [100, 90, 80, 1], [0, 0.5, 0.25, 0.999], [0.5, 1.0, 0.2501, 2.0], 172)
    assert roll.shape == (128, 344)
    pr, oo = opr.binarize_and_onoff(roll)
    on = np.nonzero(oo[:, 60])[0]
    assert list(on) == [0, 172] and list(oo[on, 60]) == [1.0, -1.0]  # legato merge: one onset, one offset
    assert pr[:, 64].sum() == 0  # zero-length note after truncation
    assert roll[60, 85] == 100 and roll[60, 86] == 90  # int(0.5*172) == 86
    assert int(0.29 * 100) == 28  # the float64 truncation case the device code must reproduce
    assert np.array_equal(oo, opr.onoff_reference_loop(pr))


def test_pianoroll_sustain_pedal_known_answer():
    """pretty_midi >= 0.2.9: inside a pedal-down span every pitch keeps the running max of its velocity sum."""
    fs = 100
    cc = [(0.2, 127), (1.5, 0), (1.8, 64), (1.9, 63), (2.5, 100)]  # last press is never released -> no effect
    roll = opr.get_piano_roll([60, 64, 67], [100, 50, 30], [0.0, 1.0, 2.6], [0.5, 1.2, 3.0], fs, cc64=cc)
    assert roll.shape == (128, 300)
    assert (roll[60, 0:150] == 100).all() and (roll[60, 150:] == 0).all()     # held from 0.5 until the release at 1.5
    assert (roll[64, 100:150] == 50).all() and roll[64, 99] == 0 and roll[64, 150] == 0
    assert (roll[67, 260:300] == 30).all() and roll[67, 259] == 0             # pedal still down at the end: untouched
    plain = opr.get_piano_roll([60, 64, 67], [100, 50, 30], [0.0, 1.0, 2.6], [0.5, 1.2, 3.0], fs)
    assert (plain[60, 50:] == 0).all()
    from ml_music_style_transfer_b200.pianoroll import pedal_spans
    assert pedal_spans(cc, fs) == [(20, 150), (180, 190)]


def test_midi_reader_control_changes(tmp_path):
    from ml_music_style_transfer_b200 import midi
    path = str(tmp_path / "p.mid")
    midi.write_midi_notes(path, [60], [90], [0.0], [0.5], cc64=[(0.25, 127), (2.0, 0)])
    p, v, s, e, cc, end_time = midi.read_midi(path)
    assert list(p) == [60] and [val for _, val in cc] == [127, 0]
    assert abs(cc[0][0] - 0.25) < 1e-9 and abs(end_time - 2.0) < 1e-9  # the late pedal release extends the roll


def test_onoff_loop_equals_difference_random():
    from ml_music_style_transfer_b200 import synth
    p, v, s, e = synth.midi_piece(0, seconds=6.0)
    pr, oo = opr.binarize_and_onoff(opr.get_piano_roll(p, v, s, e, 172))
    assert np.array_equal(oo, opr.onoff_reference_loop(pr))
    assert set(np.unique(oo)) <= {-1.0, 0.0, 1.0}


def test_upsample_definition():
    x = np.zeros((10, 128), dtype=np.int8)
    x[3, 21] = 1
    x[9, 108] = -1
    up = opr.upsample_to_audio_rate(x, 250, 22050, 1000)
    assert up.shape == (88, 1000)
    cols = (np.arange(1000) * 250) // 22050
    assert np.array_equal(up[0], (cols == 3).astype(np.int8))
    assert np.array_equal(up[87], -(cols == 9).astype(np.int8))
    assert up[:, cols >= 10].sum() == 0


def test_num_song_chunks():
    assert opr.get_num_song_chunks(5160) == (5160 - 860) // 512 - int(0.1 * ((5160 - 860) // 512))
    assert opr.get_num_song_chunks(10 ** 6) == 100


# ---- Griffin-Lim -------------------------------------------------------------------------------
@pytest.mark.parametrize("mom", [0.99, 0.0])
def test_griffinlim_matches_torchaudio_golden(mom):
    g = np.load(os.path.join(GOLD, "griffinlim_torchaudio.npz"))
    y = ogl.griffinlim(g["mag"], 8, 512, momentum=mom, init_phase=None)
    ref = g[f"y_iter8_mom{mom}"]
    assert y.shape == (512 * (g["mag"].shape[1] - 1),)
    assert np.abs(y - ref[:len(y)]).max() < 2e-4 * np.abs(ref).max()


def test_griffinlim_converges_and_inverse_map():
    from ml_music_style_transfer_b200 import synth
    y = synth.piano_clip(1, 1.0, 22050)
    S = np.abs(ostft.stft(y, 2048, 512)).astype(np.float32)
    _, hist = ogl.griffinlim(S, 16, 512, init_phase=ogl.random_phase(S.shape, 0), return_history=True)
    assert hist[-1] < hist[0] and hist[-1] < 0.5
    p = (S ** 2)
    back = ogl.logpower_to_magnitude(np.log1p(p))
    assert np.allclose(back, S, rtol=2e-3, atol=1e-4)


# ---- audio load (next row 1) -------------------------------------------------------------------
@pytest.mark.parametrize("so,sn", [(44100, 22050), (48000, 44100), (22050, 44100)])
def test_resample_matches_torchaudio_golden(so, sn):
    g = np.load(os.path.join(GOLD, "resample_torchaudio.npz"))
    y = oaudio.resample(g["x"], so, sn)
    ref = g[f"y_{so}_{sn}"]
    assert y.shape == ref.shape == (int(np.ceil(16000 * sn / so)),) and y.dtype == np.float32
    # two different realisations of the same windowed-sinc filter: agree away from the (differently padded) edges
    assert np.abs(y - ref)[2000:-2000].max() < 5e-4


def test_wav_round_trip_and_load(tmp_path):
    rng = np.random.default_rng(1)
    x = (0.5 * rng.uniform(-1, 1, (3000, 2))).astype(np.float32)
    for bits, tol in ((16, 1.0 / 32768), (24, 1.0 / (1 << 23)), (32, 0.0)):
        p = str(tmp_path / f"a{bits}.wav")
        oaudio.write_wav(p, x, 44100, bits=bits)
        back, sr = oaudio.read_wav(p)
        assert sr == 44100 and back.shape == x.shape and np.abs(back - x).max() <= tol
        from ml_music_style_transfer_b200 import audio_io
        back2, sr2 = audio_io.read_wav(p)  # the product's host decoder
        assert sr2 == sr and np.array_equal(back, back2)
    y, sr = oaudio.load(str(tmp_path / "a16.wav"), sr=22050)
    assert sr == 22050 and y.shape == (1500,) and y.dtype == np.float32
    assert oaudio.resample(x[:, 0], 44100, 44100) is not None and oaudio.resample(x[:, 0], 44100, 44100).shape == (3000,)


# ---- host logic --------------------------------------------------------------------------------
# ---- second, independent pins (tests/golden/make_golden2.py): HuggingFace audio_utils, scipy.signal, closed-form resampler ----
@pytest.mark.parametrize("sr,hop", [(22050, 512), (44100, 256)])
def test_oracle_matches_transformers_audio_utils(sr, hop):
    g = np.load(os.path.join(GOLD, "second_pins.npz"))
    y = g["y"]
    assert relerr(omel.mel_filterbank(sr, 2048, 128).astype(np.float64), g[f"hf_fb_{sr}"]) < 2e-6
    P = np.abs(ostft.stft(y, 2048, hop, out_dtype=np.complex128)) ** 2
    assert P.shape == g[f"hf_power_{hop}"].shape and relerr(P, g[f"hf_power_{hop}"]) < 1e-6
    M = omel.melspectrogram(y, sr, 2048, hop)
    assert M.shape == g[f"hf_mel_{sr}_{hop}"].shape and relerr(M.astype(np.float64), g[f"hf_mel_{sr}_{hop}"]) < 2e-6


@pytest.mark.parametrize("hop", [256, 512])
def test_oracle_matches_scipy_signal(hop):
    g = np.load(os.path.join(GOLD, "second_pins.npz"))
    D = ostft.stft(g["y"], 2048, hop)
    ref = g[f"scipy_stft_{hop}"]
    assert D.shape == ref.shape and relerr(D, ref) < 3e-7
    yi = ostft.istft(ref, hop)
    ri = g[f"scipy_istft_{hop}"]
    n = min(len(yi), len(ri))
    assert n >= hop * (ref.shape[1] - 1) - 1
    assert np.abs(yi[:n] - ri[:n])[1024:-1024].max() < 2e-6   # away from the edge frames, where the two normalise alike


@pytest.mark.parametrize("so,sn", [(44100, 22050), (48000, 44100), (22050, 44100)])
def test_resample_matches_closed_form_full_length(so, sn):
    """Full-length pin (same zero edge handling): resampy's table + linear interpolation vs the closed-form filter."""
    g = np.load(os.path.join(GOLD, "second_pins.npz"))
    y = oaudio.resample(g["rs_x"], so, sn)
    ref = g[f"rs_direct_{so}_{sn}"]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() < 2e-5 * max(1.0, np.abs(ref).max()), np.abs(y - ref).max()


def test_header_symbols_exported(built_libs):
    hdr = open(os.path.join(ROOT, "include", "mst_b200.h")).read()
    names = sorted(set(re.findall(r"\b(mst_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = ctypes.CDLL(built_libs[0])
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mst_b200.h but not exported"
    lib.mst_last_error.restype = ctypes.c_char_p
    assert lib.mst_version() >= 100


def test_torch_ops_load_and_refuse_cpu(built_libs):
    import torch
    import ml_music_style_transfer_b200 as pkg
    ops = pkg._lib.ops()
    W = ops.mel_filterbank(22050, 2048, 128, 0.0, 0.0)
    assert tuple(W.shape) == (128, 1025)
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.stft(torch.zeros(4096), 0, 3, 0)  # no CPU kernel is registered
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pkg.preprocess.process_spectrum_from_chunk(np.zeros(4096, dtype=np.float32))


def test_batch_argument_errors(built_libs):
    lib = ctypes.CDLL(built_libs[0])
    lib.mst_last_error.restype = ctypes.c_char_p
    out = ctypes.c_void_p()
    off = (ctypes.c_int64 * 1)(0)
    ln = (ctypes.c_int64 * 1)(4096)
    for n_fft in (1000, 32, 32768):   # not a power of two / outside [64, 16384] (checked before any CUDA call)
        assert lib.mst_batch_create(1, off, ln, n_fft, 256, 0, ctypes.byref(out)) == -2  # MST_ERR_UNSUPPORTED
        assert b"power of two" in lib.mst_last_error()
    ln[0] = 1000
    assert lib.mst_batch_create(1, off, ln, 2048, 256, 0, ctypes.byref(out)) == -1  # too short for reflect


def test_midi_reader_round_trip(tmp_path):
    from ml_music_style_transfer_b200 import midi
    p = [60, 60, 64, 72]
    v = [100, 90, 80, 1]
    s = [0.0, 0.5, 0.25, 1.0]
    e = [0.5, 1.0, 0.75, 2.0]
    path = str(tmp_path / "x.mid")
    midi.write_midi_notes(path, p, v, s, e)
    rp, rv, rs, re_ = midi.read_midi_notes(path)
    order = np.lexsort((rp, rs))
    assert list(rp[order]) == [60, 64, 60, 72]
    assert np.allclose(np.sort(rs), np.sort(s)) and np.allclose(np.sort(re_), np.sort(e))


def test_shard_dataset_container(tmp_path):
    """dataset.ShardManager / ShardDataset mirror io_manager.h5pyManager / train.py::Dataseth5py."""
    from ml_music_style_transfer_b200.dataset import ShardDataset, ShardManager
    rng = np.random.default_rng(0)
    m = ShardManager(str(tmp_path / "ds_train"))
    rolls, oos, specs = [], [], {"cuba": [], "upright": []}
    for n in (3, 2):  # two "songs" appended one after the other
        r = (rng.random((n, 860, 128)) < 0.05).astype(np.float64)
        o = np.sign(rng.standard_normal((n, 860, 128))).astype(np.float64) * (rng.random((n, 860, 128)) < 0.02)
        m.write_pianoroll(r, o)
        rolls.append(r); oos.append(o)
        for st in specs:
            sp = rng.random((n, 1025, 860)).astype(np.float32)
            m.write_spectrum(sp, st)
            specs[st].append(sp)
    ds = ShardDataset(str(tmp_path / "ds_train"))
    assert len(ds) == 5 and sorted(ds.styles) == ["spec_cuba", "spec_upright"]
    X, Xc, y = ds[4]
    assert X.shape == (256, 860) and Xc.shape == (1025, 860) and y.shape == (1025, 860) and X.dtype == torch.float32
    R, O = np.concatenate(rolls), np.concatenate(oos)
    assert np.array_equal(X.numpy(), np.concatenate((R[4], O[4]), axis=-1).T.astype(np.float32))
    assert any(np.array_equal(y.numpy(), np.concatenate(specs[st])[4]) for st in specs)
    with pytest.raises(ValueError):
        m.write_spectrum(np.zeros((1, 1025, 100), dtype=np.float32), "cuba")
    # modes: 'a' re-opens and appends, an explicit dtype must agree with the container, 'w' truncates like h5py 'w'
    m2 = ShardManager(str(tmp_path / "ds_train"))
    assert m2.dtype == "native" and m2.n_rows("pianoroll") == 5
    with pytest.raises(ValueError):
        ShardManager(str(tmp_path / "ds_train"), dtype="float64")
    with pytest.raises(IOError):
        ShardManager(str(tmp_path / "ds_train"), mode="r").write_pianoroll(rolls[0], oos[0])
    m3 = ShardManager(str(tmp_path / "ds_train"), dtype="float64", mode="w")
    assert m3.keys() == [] and not os.path.exists(str(tmp_path / "ds_train" / "pianoroll" / "00000.npy"))
    m3.write_pianoroll(rolls[0], oos[0])
    assert m3.n_rows("pianoroll") == 3 and m3.read("pianoroll").dtype == np.float64


def test_shard_ranges_partition():
    from ml_music_style_transfer_b200 import sharding
    for n in (0, 1, 7, 16384, 4080):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints ONE JSON line with the contract's keys (runs the oracle on the host cores)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "audio-s/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


# ---- pins at other FFT sizes / window lengths (tests/golden/make_golden3.py: torch.stft, torch.istft, torchaudio) ----------
@pytest.mark.parametrize("n_fft,hop,win", [(1024, 256, 1024), (4096, 1024, 4096), (512, 100, 512), (1024, 256, 800),
                                           (2048, 512, 1200)])
def test_oracle_other_nfft_matches_torch_golden(n_fft, hop, win):
    g = np.load(os.path.join(GOLD, "other_nfft_pins.npz"))
    tag = f"{n_fft}_{hop}_{win}"
    D = ostft.stft(g["y"], n_fft, hop, win_length=win)
    ref = g[f"stft_{tag}"]
    assert D.shape == ref.shape == (n_fft // 2 + 1, 1 + 5000 // hop)
    assert relerr(D, ref) < 3e-7
    if f"istft_{tag}" in g.files:
        y = ostft.istft(ref, hop, win_length=win)
        r = g[f"istft_{tag}"]
        assert y.shape == r.shape and np.abs(y - r).max() < 3e-6
    if f"gl6_{tag}" in g.files:
        w = ogl.griffinlim(np.abs(ref).astype(np.float32), 6, hop, win_length=win, init_phase=None)
        r = g[f"gl6_{tag}"]
        assert np.abs(w - r[:len(w)]).max() < 2e-4 * np.abs(r).max()


@pytest.mark.parametrize("sr,n_fft,n_mels", [(22050, 1024, 80), (44100, 4096, 128)])
def test_oracle_mel_filterbank_other_nfft(sr, n_fft, n_mels, built_libs):
    ref = np.load(os.path.join(GOLD, "other_nfft_pins.npz"))[f"fb_{sr}_{n_fft}_{n_mels}"]
    W = omel.mel_filterbank(sr, n_fft, n_mels)
    assert W.shape == ref.shape and np.abs(W - ref).max() / np.abs(ref).max() < 1e-5
    lib = ctypes.CDLL(built_libs[0])
    Wc = np.zeros_like(W)
    assert lib.mst_mel_filterbank_f32(sr, n_fft, n_mels, ctypes.c_double(0.0), ctypes.c_double(0.0),
                                      Wc.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(Wc, W)


def test_window_and_fft_size_validation():
    """features._check_window: host-side argument checking shared by stft / melspectrogram / griffinlim (no GPU needed)."""
    from ml_music_style_transfer_b200 import features as F
    assert F._check_window("hann", None, 2048) == 2048 and F._check_window("hann", 1200, 2048) == 1200
    for n_fft in (64, 512, 1024, 4096, 16384):
        assert F._check_window("hann", None, n_fft) == n_fft
    for n_fft in (1000, 32, 32768, 3000):
        with pytest.raises(NotImplementedError):
            F._check_window("hann", None, n_fft)
    with pytest.raises(NotImplementedError):
        F._check_window("hamming", None, 2048)
    for wl in (0, 2049):
        with pytest.raises(ValueError):
            F._check_window("hann", wl, 2048)


@pytest.mark.parametrize("dtype,bits", [(np.int16, 16), (np.int32, 32), (np.uint8, 8), (np.float32, 32), (np.float64, 64)])
@pytest.mark.parametrize("channels", [1, 2])
def test_wav_decoders_match_scipy_wavfile(tmp_path, dtype, bits, channels):
    """Independent pin of both RIFF decoders (oracle.audio.read_wav and the product's audio_io.read_wav): files written
    by scipy.io.wavfile, samples compared with scipy's own read scaled the way soundfile/librosa.load scale them
    (int16 / 2**15, int32 / 2**31, (uint8 - 128) / 128, floats as they are)."""
    import scipy.io.wavfile as wavfile
    from ml_music_style_transfer_b200 import audio_io
    rng = np.random.default_rng(bits + channels)
    n = 777
    if np.issubdtype(dtype, np.floating):
        data = (0.5 * rng.standard_normal((n, channels))).astype(dtype)
        want = data.astype(np.float32)
    else:
        info = np.iinfo(dtype)
        data = rng.integers(info.min, info.max, size=(n, channels), endpoint=True).astype(dtype)
        if dtype == np.uint8:
            want = (data.astype(np.float32) - 128.0) / 128.0
        else:
            want = (data.astype(np.float64) / float(2 ** (bits - 1))).astype(np.float32)
    path = str(tmp_path / "s.wav")
    wavfile.write(path, 48000, data if channels > 1 else data[:, 0])
    sr_s, back_s = wavfile.read(path)
    assert sr_s == 48000 and np.array_equal(back_s.reshape(n, channels), data)
    for reader in (oaudio.read_wav, audio_io.read_wav):
        x, sr = reader(path)
        assert sr == 48000 and x.shape == (n, channels) and x.dtype == np.float32
        assert np.array_equal(x, want), reader.__module__
