"""n_fft other than 2048: the general path (csrc/generic_fft.cu) behind the same entry points, against the oracle.
The reference only uses n_fft = 2048 (preprocessing/preprocess.py:25); librosa's signatures take any size."""
import numpy as np
import pytest
import torch

from oracle import griffinlim as ogl, mel as omel, stft as ostft

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def pkg(gpu):
    import ml_music_style_transfer_b200 as p
    return p


def clip(seed, n):
    from ml_music_style_transfer_b200 import synth
    return synth.piano_clip(seed, n / 22050.0, 22050)[:n]


@pytest.mark.parametrize("n_fft,hop", [(64, 16), (256, 64), (512, 100), (1024, 256), (4096, 1024), (4096, 512), (8192, 2048),
                                       (16384, 4096)])
@pytest.mark.parametrize("pad", ["reflect", "constant"])
def test_stft_complex_other_sizes(pkg, n_fft, hop, pad):
    y = clip(n_fft + hop, 3 * n_fft + 1234)
    ref = ostft.stft(y, n_fft, hop, pad_mode=pad)
    got = pkg.features.stft(y, n_fft=n_fft, hop_length=hop, pad_mode=pad)
    assert got.shape == ref.shape == (n_fft // 2 + 1, 1 + len(y) // hop) and got.dtype == np.complex64
    assert got.flags.f_contiguous
    assert rel_l2(got, ref) < TOL, rel_l2(got, ref)
    assert np.abs(got - ref).max() <= TOL * np.abs(ref).max()


def test_stft_known_answers_n_fft_1024(pkg):
    n = np.arange(8192)
    D = pkg.features.stft(np.cos(2 * np.pi * 50 * n / 1024).astype(np.float32), n_fft=1024, hop_length=256)[:, 16]
    assert abs(abs(D[50]) - 256.0) < 0.05 and abs(abs(D[49]) - 128.0) < 0.05 and np.abs(D[54:200]).max() < 0.05
    D = pkg.features.stft(np.ones(8192, dtype=np.float32), n_fft=1024, hop_length=256)[:, 16]
    assert abs(D[0].real - 512.0) < 0.01 and abs(abs(D[1]) - 256.0) < 0.01 and np.abs(D[2:]).max() < 0.01


@pytest.mark.parametrize("n_fft", [1024, 4096])
@pytest.mark.parametrize("out", ["magnitude", "power", "log1p_power"])
def test_epilogues_and_layouts(pkg, gpu, n_fft, out):
    F = pkg.features
    hop = n_fft // 4
    lens = [3 * n_fft + 17, 5 * n_fft, 2 * n_fft + 1]
    ys = [clip(60 + i, L) for i, L in enumerate(lens)]
    refs = []
    for y in ys:
        m = np.abs(ostft.stft(y, n_fft, hop, out_dtype=np.complex128))
        refs.append({"magnitude": m, "power": m ** 2, "log1p_power": np.log1p(m ** 2)}[out])
    one = F.spectrogram(ys[0], hop, out=out, n_fft=n_fft)
    assert one.shape == refs[0].shape and rel_l2(one, refs[0]) < TOL
    audio = torch.from_numpy(np.concatenate(ys)).to(gpu)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    with F.ClipBatch.from_clips(offs, lens, hop, device=gpu, n_fft=n_fft) as b:
        assert b.n_bins == n_fft // 2 + 1
        fm = F.stft_batch(audio, b, out, F.FRAME_MAJOR).cpu().numpy()
        bm = F.stft_batch(audio, b, out, F.BIN_MAJOR).cpu().numpy()
    o = 0
    for r in refs:
        K, T = r.shape
        assert rel_l2(fm[o:o + K * T].reshape(T, K).T, r) < TOL
        assert rel_l2(bm[o:o + K * T].reshape(K, T), r) < TOL
        o += K * T
    assert o == fm.shape[0] == bm.shape[0]


@pytest.mark.parametrize("n_fft,hop,n_mels,sr", [(1024, 256, 80, 22050), (4096, 1024, 128, 44100), (512, 128, 40, 16000)])
def test_melspectrogram_other_sizes(pkg, n_fft, hop, n_mels, sr):
    y = clip(5, 30000)
    ref = omel.melspectrogram(y, sr, n_fft, hop, n_mels).astype(np.float64)
    got = pkg.features.melspectrogram(y=y, sr=sr, n_fft=n_fft, hop_length=hop, n_mels=n_mels)
    assert got.shape == ref.shape == (n_mels, 1 + len(y) // hop)
    assert rel_l2(got, ref) < TOL and np.abs(got - ref).max() <= TOL * np.abs(ref).max()
    got_log = pkg.features.logmel(y, sr=sr, n_fft=n_fft, hop_length=hop, n_mels=n_mels)
    assert rel_l2(got_log, np.log1p(ref)) < TOL


@pytest.mark.parametrize("n_fft,hop", [(1024, 256), (4096, 1024), (512, 100), (1024, 512)])
def test_griffinlim_other_sizes(pkg, n_fft, hop):
    """Waveform parity while float32 rounding has not been amplified (0-3 iterations), spectral convergence within 1e-3
    after 32; a hop that does not divide n_fft takes the same code (general overlap-add)."""
    y = clip(36, 12 * n_fft + 321)
    S = np.abs(ostft.stft(y, n_fft, hop)).astype(np.float32)
    u = ogl.random_phase(S.shape, 2)
    for n_iter in (0, 1, 3):
        w_ref = ogl.griffinlim(S, n_iter, hop, init_phase=u)
        w_got = pkg.features.griffinlim(S, n_iter=n_iter, hop_length=hop, init_phase=u)
        assert w_got.shape == w_ref.shape == (hop * (S.shape[1] - 1),)
        assert rel_l2(w_got, w_ref.astype(np.float64)) < 5e-4, (n_fft, hop, n_iter, rel_l2(w_got, w_ref.astype(np.float64)))
    w_ref = ogl.griffinlim(S, 32, hop, init_phase=u)
    w_got = pkg.features.griffinlim(S, n_iter=32, hop_length=hop, init_phase=u)
    sc_ref = ogl.spectral_convergence(S, w_ref, hop, n_fft=n_fft)
    sc_got = ogl.spectral_convergence(S, w_got, hop, n_fft=n_fft)
    assert abs(sc_ref - sc_got) <= 1e-3, (sc_ref, sc_got)
    # the fused convergence metric of the general path agrees with the oracle's
    assert abs(pkg.features.spectral_convergence(S, w_got, hop) - sc_got) < 1e-4 * max(sc_got, 1e-3)
    # momentum 0 (classic Griffin-Lim) and the device RNG / unit-phase starts run and converge
    for kw in (dict(momentum=0.0, init_phase=u), dict(random_state=7), dict(init=None)):
        w = pkg.features.griffinlim(S, n_iter=16, hop_length=hop, **kw)
        assert np.isfinite(w).all() and ogl.spectral_convergence(S, w, hop, n_fft=n_fft) < 0.6


def test_griffinlim_ragged_batch_n_fft_1024(pkg, gpu):
    """Ragged batch, bin-major log1p-power input and win_length < n_fft on the general path."""
    F = pkg.features
    n_fft, hop, win = 1024, 256, 800
    frames = [40, 97, 23]
    S_list, u_list = [], []
    for i, T in enumerate(frames):
        yy = clip(40 + i, hop * (T - 1))
        S_list.append(np.abs(ostft.stft(yy, n_fft, hop, win_length=win)).astype(np.float32))
        u_list.append(ogl.random_phase(S_list[-1].shape, 10 + i).astype(np.float32))
        assert S_list[-1].shape == (513, T)
    logp = np.concatenate([np.log1p(S.astype(np.float64) ** 2).astype(np.float32).ravel() for S in S_list])
    ph = np.concatenate([u.ravel() for u in u_list])
    with F.ClipBatch.from_frames(frames, hop, device=gpu, n_fft=n_fft, win_length=win) as gb:
        out = F.griffinlim_batch(torch.from_numpy(logp).to(gpu), gb, n_iter=2, init_phase=torch.from_numpy(ph).to(gpu),
                                 layout=F.BIN_MAJOR, is_log1p_power=True).cpu().numpy()
    o = 0
    for S, u, T in zip(S_list, u_list, frames):
        L = hop * (T - 1)
        mag = ogl.logpower_to_magnitude(np.log1p(S.astype(np.float64) ** 2).astype(np.float32))
        ref = ogl.griffinlim(mag, 2, hop, win_length=win, init_phase=u)
        assert rel_l2(out[o:o + L], ref.astype(np.float64)) < 5e-4
        o += L
    assert o == out.shape[0]


def test_unsupported_sizes_raise(pkg):
    y = clip(1, 9000)
    for n_fft in (1000, 32, 32768):
        with pytest.raises(NotImplementedError):
            pkg.features.stft(y, n_fft=n_fft, hop_length=128)
    with pytest.raises(NotImplementedError):
        pkg.features.mel_to_stft(np.ones((128, 10), dtype=np.float32), n_fft=1024)
