"""``librosa.load(path, sr=)`` for WAV files: host RIFF decode, device mono-mix + 'kaiser_best' resampling.
SURVEY section 8f "next" #1 -- the step immediately before P1 (preprocessing/preprocess.py:99-115, model/inference.py:54).
"""
import struct

import numpy as np
import torch

from . import _lib


def read_wav(path):
    """-> (float32 ndarray (n, channels), native sample rate).  PCM 8/16/24/32-bit and IEEE float RIFF/WAVE files,
    scaled like soundfile's float32 read (int16 / 2**15, int24 / 2**23, int32 / 2**31, uint8 (x-128)/128)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, raw = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: real tag is the sub-format
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1 and bits == 8:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif tag == 1 and bits == 16:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 24:
        b = np.frombuffer(raw[:len(raw) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float32) / float(1 << 23)
    elif tag == 1 and bits == 32:
        x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / float(1 << 31)).astype(np.float32)
    elif tag == 3 and bits in (32, 64):
        x = np.frombuffer(raw, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (tag {tag}, {bits} bits)")
    n = len(x) // ch
    return x[:n * ch].reshape(n, ch), sr


def resample(y, orig_sr, target_sr):
    """librosa.resample(y, orig_sr, target_sr, res_type='kaiser_best') for a mono signal (NumPy or CUDA tensor)."""
    was_np = not isinstance(y, torch.Tensor)
    device = _lib.require_cuda(None if was_np else y.device)
    t = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(device) if was_np else y.to(torch.float32)
    out = _lib.ops().resample(t.contiguous().view(-1), int(orig_sr), int(target_sr))
    return out.cpu().numpy() if was_np else out


def load(path, sr=22050, as_numpy=True, device=None):
    """librosa.load(path, sr=sr) -> (y float32 mono, sr).  sr=None keeps the native rate."""
    x, sr_native = read_wav(path)
    device = _lib.require_cuda(device)
    t = torch.from_numpy(x).to(device)
    y = _lib.ops().mono_mix(t.contiguous()) if t.shape[1] > 1 else t[:, 0].contiguous()   # librosa.to_mono
    if sr is not None and sr != sr_native:
        y = _lib.ops().resample(y.contiguous(), int(sr_native), int(sr))
        sr_native = sr
    return (y.cpu().numpy() if as_numpy else y), sr_native
