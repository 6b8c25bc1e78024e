"""Host-side operators of the hot path: thin Python over ``torch.ops.mst_b200`` (which wraps the C ABI).

Names and defaults follow the third-party entry points the reference calls
(``librosa.stft`` preprocess.py:48, ``librosa.feature.melspectrogram`` tests/plot_spec.py:20,
``librosa.filters.mel``, ``librosa.griffinlim`` model/inference.py:110) so that the drop-in
modules ``preprocess`` / ``inference`` read like the reference.  NumPy in -> NumPy out (host<->device
copies included, as the reference's users expect); CUDA tensor in -> CUDA tensor out, no host round trip.
"""
import numpy as np
import torch

from . import _lib

N_FFT = 2048
N_BINS = N_FFT // 2 + 1

OUT_COMPLEX, OUT_MAGNITUDE, OUT_POWER, OUT_LOG1P_POWER = 0, 1, 2, 3
FRAME_MAJOR, BIN_MAJOR = 0, 1
PAD_MODES = {"reflect": 0, "constant": 1}
DTYPE_CODES = {torch.int8: 0, torch.float32: 1, torch.float64: 2}


def _pad_code(pad_mode):
    try:
        return PAD_MODES[pad_mode]
    except KeyError:
        raise ValueError(f"pad_mode={pad_mode!r} unsupported (reflect | constant)") from None


def _check_window(window, win_length, n_fft):
    """-> win_length to use.  window='hann' (any win_length <= n_fft, centre-padded like librosa); n_fft = 2048 runs the
    tuned warp-per-frame kernels, every other power of two in [64, 16384] the general path (csrc/generic_fft.cu)."""
    if window != "hann":
        raise NotImplementedError("only window='hann' is implemented (the reference uses no other)")
    n_fft = int(n_fft)
    if n_fft != N_FFT and not (64 <= n_fft <= 16384 and n_fft & (n_fft - 1) == 0):
        raise NotImplementedError(f"n_fft={n_fft}: n_fft must be a power of two in [64, 16384]")
    if win_length is None:
        return n_fft
    win_length = int(win_length)
    if not 1 <= win_length <= n_fft:
        raise ValueError(f"win_length={win_length} must be in [1, n_fft]")
    return win_length


class ClipBatch:
    """Ragged set of clips inside one audio buffer (``mst_batch_t``).  Chunks may overlap."""

    def __init__(self, handle, n_clips, hop, device, n_fft=N_FFT):
        self.handle, self.n_clips, self.hop, self.device = handle, n_clips, hop, device
        self.n_fft, self.n_bins = int(n_fft), int(n_fft) // 2 + 1
        o = _lib.ops()
        self.total_frames = int(o.batch_total_frames(handle))
        self.total_samples = int(o.batch_total_samples(handle))

    @classmethod
    def from_clips(cls, offsets, lengths, hop, pad_mode="reflect", device=None, n_fft=N_FFT, win_length=None):
        device = _lib.require_cuda(device)
        offsets = torch.as_tensor(np.asarray(offsets, dtype=np.int64))
        lengths = torch.as_tensor(np.asarray(lengths, dtype=np.int64))
        try:
            h = _lib.ops().batch_create(offsets, lengths, n_fft, int(hop), _pad_code(pad_mode), device.index,
                                        n_fft if win_length is None else int(win_length))
        except RuntimeError as e:
            if "too short" in str(e) or "empty clip" in str(e):
                # np.pad(mode='reflect') / librosa raise for inputs shorter than the padding: keep the exception type
                raise ValueError(str(e)) from None
            raise
        return cls(h, int(offsets.numel()), int(hop), device, n_fft)

    @classmethod
    def uniform(cls, n_clips, clip_length, hop, clip_stride=None, pad_mode="reflect", device=None, win_length=None,
                n_fft=N_FFT):
        stride = clip_length if clip_stride is None else clip_stride
        offsets = np.arange(n_clips, dtype=np.int64) * stride
        return cls.from_clips(offsets, np.full(n_clips, clip_length, dtype=np.int64), hop, pad_mode, device,
                              n_fft=n_fft, win_length=win_length)

    @classmethod
    def from_frames(cls, frames_per_clip, hop, pad_mode="reflect", device=None, n_fft=N_FFT, win_length=None):
        device = _lib.require_cuda(device)
        frames = torch.as_tensor(np.asarray(frames_per_clip, dtype=np.int64))
        try:
            h = _lib.ops().batch_create_from_frames(frames, n_fft, int(hop), _pad_code(pad_mode), device.index,
                                                    n_fft if win_length is None else int(win_length))
        except RuntimeError as e:
            if "too short" in str(e):
                raise ValueError(str(e)) from None
            raise
        return cls(h, int(frames.numel()), int(hop), device, n_fft)

    def clip_frames(self, c):
        return int(_lib.ops().batch_clip_frames(self.handle, c))

    def close(self):
        if getattr(self, "handle", 0):
            _lib.ops().batch_destroy(self.handle)
            self.handle = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MelPlan:
    """Device-resident filterbank (``mst_mel_plan_t``) built from librosa.filters.mel semantics."""

    _cache = {}

    def __init__(self, weights, device):
        self.weights = weights  # CPU float32 (n_mels, 1 + n_fft/2)
        self.n_mels = int(weights.shape[0])
        self.device = device
        self.handle = _lib.ops().mel_plan_create(weights.contiguous(), device.index)

    @classmethod
    def get(cls, sr, n_fft=N_FFT, n_mels=128, fmin=0.0, fmax=None, device=None):
        device = _lib.require_cuda(device)
        key = (int(sr), int(n_fft), int(n_mels), float(fmin), None if fmax is None else float(fmax), device.index)
        if key not in cls._cache:
            cls._cache[key] = cls(mel_filterbank(sr, n_fft, n_mels, fmin, fmax), device)
        return cls._cache[key]


class MelInversePlan:
    """Device-resident data of the mel inversion (``mst_mel_inverse_plan_t``): pinv of the filterbank, its sparse
    column lists and the gradient step, built once per (sr, n_fft, n_mels, fmin, fmax, device)."""

    _cache = {}

    def __init__(self, weights, device):
        self.n_mels = int(weights.shape[0])
        self.device = device
        self.handle = _lib.ops().mel_inverse_plan_create(weights.contiguous(), device.index)

    @classmethod
    def get(cls, sr, n_fft=N_FFT, n_mels=128, fmin=0.0, fmax=None, device=None):
        device = _lib.require_cuda(device)
        key = (int(sr), int(n_fft), int(n_mels), float(fmin), None if fmax is None else float(fmax), device.index)
        if key not in cls._cache:
            cls._cache[key] = cls(mel_filterbank(sr, n_fft, n_mels, fmin, fmax), device)
        return cls._cache[key]


def mel_filterbank(sr, n_fft=N_FFT, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney') -> CPU float32 (n_mels, 1+n_fft/2)."""
    return _lib.ops().mel_filterbank(int(sr), int(n_fft), int(n_mels), float(fmin), 0.0 if fmax is None else float(fmax))


def to_numpy(t):
    """Device tensor -> NumPy.  Large results go through page-locked memory from torch's caching host allocator (one DMA,
    no pageable staging): the returned array aliases that pinned block, which returns to the cache when the array dies."""
    t = t.contiguous()
    if t.numel() * t.element_size() < (1 << 20):
        return t.cpu().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def _to_device_audio(y, device=None):
    """Returns (float32 CUDA 1-D tensor, was_numpy)."""
    if isinstance(y, torch.Tensor):
        if not y.is_cuda:
            raise TypeError("torch inputs must be CUDA tensors (there is no CPU path); pass NumPy for host data")
        return y.to(torch.float32).contiguous().view(-1), False
    device = _lib.require_cuda(device)
    a = np.ascontiguousarray(y, dtype=np.float32).reshape(-1)
    return torch.from_numpy(a).to(device, non_blocking=False), True


def stft_batch(audio, batch, out="log1p_power", layout=FRAME_MAJOR):
    """Raw batched op.  ``audio``: CUDA float32 buffer; returns the flat output tensor of mst_stft_f32."""
    mode = {"complex": OUT_COMPLEX, "magnitude": OUT_MAGNITUDE, "power": OUT_POWER, "log1p_power": OUT_LOG1P_POWER}[out]
    return _lib.ops().stft(audio, batch.handle, mode, layout)


def melspectrogram_batch(audio, batch, plan, log1p=False, layout=FRAME_MAJOR):
    return _lib.ops().stft_mel(audio, batch.handle, plan.handle, plan.n_mels, bool(log1p), layout)


def stft(y, n_fft=N_FFT, hop_length=None, win_length=None, window="hann", center=True, pad_mode="reflect"):
    """librosa.stft drop-in for one clip: complex64 (1 + n_fft/2, T), Fortran-ordered like librosa's result."""
    win_length = _check_window(window, win_length, n_fft)
    if not center:
        raise NotImplementedError("center=False is not used by the reference")
    hop = win_length // 4 if hop_length is None else int(hop_length)
    a, was_np = _to_device_audio(y)
    with ClipBatch.uniform(1, a.numel(), hop, pad_mode=pad_mode, device=a.device, win_length=win_length, n_fft=n_fft) as b:
        out = stft_batch(a, b, "complex")  # [T][K] memory; its transpose view == librosa's Fortran-ordered (K, T)
    return to_numpy(out).T if was_np else out.t()


def spectrogram(y, hop_length, out="log1p_power", pad_mode="reflect", n_fft=N_FFT):
    """(1 + n_fft/2, T) float32 epilogue of the STFT for one clip (Fortran-ordered view)."""
    _check_window("hann", None, n_fft)
    a, was_np = _to_device_audio(y)
    with ClipBatch.uniform(1, a.numel(), int(hop_length), pad_mode=pad_mode, device=a.device, n_fft=n_fft) as b:
        o = stft_batch(a, b, out).view(b.total_frames, b.n_bins)
    return to_numpy(o).T if was_np else o.t()


def melspectrogram(y=None, sr=22050, n_fft=N_FFT, hop_length=512, n_mels=128, fmin=0.0, fmax=None, pad_mode="reflect",
                   log1p=False):
    """librosa.feature.melspectrogram(y=, sr=, n_fft=, hop_length=) drop-in: float32 (n_mels, T)."""
    _check_window("hann", None, n_fft)
    a, was_np = _to_device_audio(y)
    plan = MelPlan.get(sr, n_fft, n_mels, fmin, fmax, a.device)
    with ClipBatch.uniform(1, a.numel(), int(hop_length), pad_mode=pad_mode, device=a.device, n_fft=n_fft) as b:
        o = melspectrogram_batch(a, b, plan, log1p=log1p, layout=BIN_MAJOR).view(n_mels, b.total_frames)
    return to_numpy(o) if was_np else o


def logmel(y, sr=22050, n_fft=N_FFT, hop_length=512, n_mels=128, **kw):
    """log1p(mel_basis @ |stft|^2): the reference's log1p convention (preprocess.py:49) on the mel projection."""
    return melspectrogram(y=y, sr=sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels, log1p=True, **kw)


def griffinlim_batch(S, batch, n_iter=32, momentum=0.99, init_phase=None, init="random", seed=0, layout=BIN_MAJOR,
                     is_log1p_power=False):
    """Raw batched op: S is a flat / shaped CUDA float32 tensor holding every clip's (1025 x T_c) block in `layout`."""
    init_mode = {"random": 0, None: 1}[init]
    return _lib.ops().griffinlim(S.contiguous().view(-1), layout, bool(is_log1p_power), batch.handle, int(n_iter),
                                 float(momentum), None if init_phase is None else init_phase.contiguous().view(-1),
                                 init_mode, int(seed))


def griffinlim(S, n_iter=32, hop_length=None, win_length=None, window="hann", momentum=0.99, init="random",
               random_state=None, init_phase=None, pad_mode="reflect"):
    """librosa.griffinlim drop-in for one (1 + n_fft/2, T) magnitude spectrogram -> float32 waveform of hop*(T-1) samples
    (n_fft is inferred from the bin count, as librosa does).

    ``random_state=int`` reproduces librosa's ``RandomState(seed).rand(*S.shape)`` initial phase exactly (drawn on the
    host); ``init_phase`` supplies the uniform [0,1) field directly; with ``random_state=None`` the phase comes from the
    device's counter-based RNG, seeded from NumPy's global generator (the generator librosa itself would draw from, so
    ``np.random.seed`` makes runs repeatable here too).
    """
    was_np = not isinstance(S, torch.Tensor)
    device = _lib.require_cuda(None if was_np else S.device)
    if was_np:
        S = torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32)).to(device)
    if S.dim() != 2 or S.shape[0] < 2:
        raise ValueError(f"S must be (1 + n_fft/2, T); got {tuple(S.shape)}")
    n_fft = 2 * (S.shape[0] - 1)
    n_bins = int(S.shape[0])
    win_length = _check_window(window, win_length, n_fft)
    hop = win_length // 4 if hop_length is None else int(hop_length)
    T = int(S.shape[1])
    if init_phase is None and init == "random" and isinstance(random_state, (int, np.integer)):
        init_phase = np.random.RandomState(int(random_state)).rand(n_bins, T)
    if init_phase is not None and not isinstance(init_phase, torch.Tensor):
        init_phase = torch.from_numpy(np.ascontiguousarray(init_phase, dtype=np.float32)).to(device)
    # a (K,T) tensor whose memory is [T][K] (librosa's Fortran order) is consumed without a transpose
    if S.stride() == (1, n_bins) and (init_phase is None or init_phase.stride() == (1, n_bins)):
        layout, S_flat = FRAME_MAJOR, S.t()
        ph = None if init_phase is None else init_phase.t()
    else:
        layout, S_flat = BIN_MAJOR, S.contiguous()
        ph = None if init_phase is None else init_phase.to(torch.float32).contiguous()
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    with ClipBatch.from_frames([T], hop, pad_mode=pad_mode, device=device, n_fft=n_fft, win_length=win_length) as b:
        y = griffinlim_batch(S_flat.to(torch.float32), b, n_iter, momentum, ph, init, seed, layout)
    return to_numpy(y) if was_np else y


def mel_to_stft_batch(M, batch, plan, power=2.0, layout=BIN_MAJOR, max_iter=200, tol=1e-6):
    """Raw batched op: ``M`` holds every clip's (n_mels x T_c) block in ``layout``; ``batch`` from ClipBatch.from_frames.
    Returns frame-major magnitudes (total_frames, 1025) -- the form griffinlim_batch consumes in place."""
    return _lib.ops().mel_to_stft(M.contiguous().view(-1), layout, batch.handle, plan.handle, plan.n_mels, float(power),
                                  int(max_iter), float(tol))


def mel_to_stft(M, sr=22050, n_fft=N_FFT, power=2.0, fmin=0.0, fmax=None, max_iter=200, tol=1e-6):
    """librosa.feature.inverse.mel_to_stft drop-in: (n_mels, T) mel power spectrogram -> (1025, T) magnitudes
    ``nnls(mel_basis, M) ** (1 / power)``.  Same start point as librosa.util.nnls (clipped least squares); the NNLS
    problem of every frame is then solved to ``tol`` (relative residual) instead of stopping where L-BFGS-B's scaled
    projected-gradient test stops, so the residual ||mel_basis @ S**power - M|| is <= librosa's."""
    if int(n_fft) != N_FFT:
        raise NotImplementedError("mel inversion is built for n_fft=2048 (the size the reference uses)")
    was_np = not isinstance(M, torch.Tensor)
    device = _lib.require_cuda(None if was_np else M.device)
    if was_np:
        M = torch.from_numpy(np.ascontiguousarray(M, dtype=np.float32)).to(device)
    if M.dim() != 2:
        raise ValueError(f"M must be (n_mels, T); got {tuple(M.shape)}")
    n_mels, T = int(M.shape[0]), int(M.shape[1])
    plan = MelInversePlan.get(sr, n_fft, n_mels, fmin, fmax, device)
    with ClipBatch.from_frames([T], n_fft // 4, pad_mode="constant", device=device) as b:   # hop is irrelevant here
        S = mel_to_stft_batch(M.to(torch.float32), b, plan, power, BIN_MAJOR, max_iter, tol)
    return to_numpy(S).T if was_np else S.t()


def mel_to_audio(M, sr=22050, n_fft=N_FFT, hop_length=512, win_length=None, window="hann", center=True, pad_mode="reflect",
                 power=2.0, n_iter=32, momentum=0.99, init="random", random_state=None, init_phase=None, fmin=0.0, fmax=None):
    """librosa.feature.inverse.mel_to_audio drop-in (tests/test_griffinlim.py:24, commented in the reference):
    ``griffinlim(mel_to_stft(M, sr, n_fft, power), n_iter, hop_length, ...)``; the magnitudes never leave the device."""
    if not center:
        raise NotImplementedError("center=False is not used by the reference")
    was_np = not isinstance(M, torch.Tensor)
    device = _lib.require_cuda(None if was_np else M.device)
    Md = torch.from_numpy(np.ascontiguousarray(M, dtype=np.float32)).to(device) if was_np else M
    S = mel_to_stft(Md, sr, n_fft, power, fmin, fmax)        # (1025, T) view of frame-major memory
    y = griffinlim(S, n_iter=n_iter, hop_length=hop_length, win_length=win_length, window=window, momentum=momentum,
                   init=init, random_state=random_state,
                   init_phase=None if init_phase is None else (init_phase if isinstance(init_phase, torch.Tensor) else
                                                               torch.from_numpy(np.ascontiguousarray(init_phase, dtype=np.float32)).to(device).t().contiguous().t()),
                   pad_mode=pad_mode)
    return to_numpy(y) if was_np else y


def spectral_convergence_batch(y, batch, S, layout=BIN_MAJOR):
    """Per-clip || |STFT(y_c)| - S_c ||_F / || S_c ||_F, fused into the STFT kernel (no spectrogram is materialised).
    ``batch`` describes the clips of the waveform buffer ``y``; ``S`` holds one (1025 x T_c) block per clip in ``layout``."""
    return _lib.ops().spectral_convergence(y.contiguous().view(-1), batch.handle, S.contiguous().view(-1), layout)


def spectral_convergence(S, y, hop_length, pad_mode="reflect"):
    """|| |STFT(y)| - S ||_F / ||S||_F for one clip (normalised form of the loss printed at model/inference.py:149-150)."""
    a, _ = _to_device_audio(y)
    if not isinstance(S, torch.Tensor):
        S = torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32)).to(a.device)
    n_fft = 2 * (int(S.shape[0]) - 1)
    _check_window("hann", None, n_fft)
    with ClipBatch.uniform(1, a.numel(), int(hop_length), pad_mode=pad_mode, device=a.device, n_fft=n_fft) as b:
        if S.shape != (b.n_bins, b.total_frames):
            raise ValueError(f"S must be ({b.n_bins}, {b.total_frames}) for this waveform; got {tuple(S.shape)}")
        sc = spectral_convergence_batch(a, b, S.to(torch.float32), BIN_MAJOR)
    return float(sc[0])
