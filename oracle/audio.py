"""librosa-0.8 ``load`` (WAV decode -> mono -> resampy 'kaiser_best' resample) restated in NumPy
(oracle; test infrastructure).  SURVEY section 8f "next" #1: the step immediately before P1.

Reference call sites: preprocessing/preprocess.py:106 (``librosa.load(audio_file[0], sr=hp.sr)``),
model/inference.py:54, tests/test_griffinlim.py:16.

Upstream semantics followed (neither librosa nor resampy is present in this image):
  * soundfile decode to float32: int16 / 2**15, int24 / 2**23, int32 / 2**31, uint8 (x-128)/128, float as is;
  * ``librosa.to_mono``: mean over channels;
  * ``librosa.resample(res_type='kaiser_best')``: resampy 0.2.2 ``resample_f`` -- band-limited sinc interpolation with
    the half-window ``rolloff*sinc(rolloff*t)*kaiser(beta)`` sampled 512 times per zero crossing over 64 zero crossings
    (beta = 14.769656459379492, rolloff = 0.9475937167399596: the published 'kaiser_best' parameters), linear
    interpolation between table entries, left and right wings accumulated tap by tap; the window is scaled by the ratio
    when down-sampling.  Output length floor(n*ratio), then ``fix_length`` to ceil(n*ratio) (zero padding).
PARITY UNPINNED by the reference; cross-checked in tests against torchaudio's ``sinc_interp_kaiser`` resampler with the
same three parameters (the configuration torchaudio documents as matching librosa's kaiser_best).
"""
import struct

import numpy as np

KAISER_BEST = dict(num_zeros=64, precision=9, beta=14.769656459379492, rolloff=0.9475937167399596)


def kaiser_best_window():
    """-> (interp_win float64 [num_zeros*2**precision + 1], num_table)."""
    nz, prec, beta, rolloff = (KAISER_BEST[k] for k in ("num_zeros", "precision", "beta", "rolloff"))
    num_bits = 2 ** prec
    n = num_bits * nz
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, nz, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits


def resample(y, sr_orig, sr_new):
    """librosa.resample(y, sr_orig, sr_new, res_type='kaiser_best') for 1-D float32 input."""
    y = np.asarray(y, dtype=np.float32)
    if sr_orig == sr_new:
        return y
    ratio = float(sr_new) / sr_orig
    n_out = int(y.shape[0] * ratio)
    n_fix = int(np.ceil(y.shape[0] * ratio))
    interp_win, num_table = kaiser_best_window()
    if ratio < 1:
        interp_win = interp_win * ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    scale = min(1.0, ratio)
    time_increment = 1.0 / ratio
    index_step = int(scale * num_table)
    nwin = interp_win.shape[0]
    n_orig = y.shape[0]
    t = np.arange(n_out, dtype=np.float64)
    time_register = t * time_increment
    n = time_register.astype(np.int64)
    out = np.zeros(n_out, dtype=np.float64)
    x = y.astype(np.float64)
    # left wing
    frac = scale * (time_register - n)
    index_frac = frac * num_table
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    i_max = np.minimum(n + 1, (nwin - offset) // index_step)
    for i in range(int(i_max.max())):
        m = i < i_max
        idx = offset[m] + i * index_step
        out[m] += (interp_win[idx] + eta[m] * interp_delta[idx]) * x[n[m] - i]
    # right wing
    frac = scale - frac
    index_frac = frac * num_table
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    k_max = np.minimum(n_orig - n - 1, (nwin - offset) // index_step)
    for k in range(int(k_max.max())):
        m = k < k_max
        idx = offset[m] + k * index_step
        out[m] += (interp_win[idx] + eta[m] * interp_delta[idx]) * x[n[m] + k + 1]
    res = np.zeros(n_fix, dtype=np.float32)
    res[:n_out] = out.astype(np.float32)
    return res


def read_wav(path):
    """-> (float32 array (n, channels), sample_rate).  PCM 8/16/24/32-bit and IEEE float 32/64 RIFF files."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt, raw = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError("missing fmt or data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(raw[:len(raw) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v >= 1 << 23, v - (1 << 24), v)
            x = v.astype(np.float32) / float(1 << 23)
        elif bits == 32:
            x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / float(1 << 31)).astype(np.float32)
        else:
            raise ValueError(f"unsupported PCM width {bits}")
    elif tag == 3:
        x = np.frombuffer(raw, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"unsupported WAV format tag {tag}")
    n = len(x) // ch
    return x[:n * ch].reshape(n, ch), sr


def load(path, sr=22050):
    """librosa.load(path, sr=sr): float32 mono at `sr`."""
    x, sr_native = read_wav(path)
    y = x.mean(axis=1, dtype=np.float32) if x.shape[1] > 1 else x[:, 0]
    if sr is not None and sr != sr_native:
        y = resample(y, sr_native, sr)
        sr_native = sr
    return np.ascontiguousarray(y, dtype=np.float32), sr_native


def write_wav(path, x, sr, bits=16):
    """Test helper: PCM16 / PCM24 / float32 writer.  x: (n,) or (n, channels) float in [-1, 1)."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[:, None]
    ch = x.shape[1]
    if bits == 16:
        raw = np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2").tobytes()
        tag, width = 1, 2
    elif bits == 24:
        v = np.clip(np.round(x * float(1 << 23)), -(1 << 23), (1 << 23) - 1).astype(np.int32).ravel()
        v = np.where(v < 0, v + (1 << 24), v)
        raw = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], axis=1).astype(np.uint8).tobytes()
        tag, width = 1, 3
    else:
        raw = x.astype("<f4").tobytes()
        tag, width, bits = 3, 4, 32
    fmt = struct.pack("<HHIIHH", tag, ch, sr, sr * ch * width, ch * width, bits)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(raw)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<I", len(fmt)) + fmt)
        f.write(b"data" + struct.pack("<I", len(raw)) + raw)
