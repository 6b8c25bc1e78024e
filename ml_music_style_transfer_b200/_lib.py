"""Loader for the native libraries.  There is deliberately no fallback: if the CUDA extension is
missing or no GPU is present, the product path raises."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_CUDA = os.path.join(_HERE, "libmst_b200.so")
LIB_OPS = os.path.join(_HERE, "libmst_torch_ops.so")

_ops = None
_cdll = None


class NativeLibraryMissing(RuntimeError):
    pass


def ops():
    """``torch.ops.mst_b200`` after loading libmst_torch_ops.so (which links libmst_b200.so)."""
    global _ops
    if _ops is None:
        if not (os.path.exists(LIB_CUDA) and os.path.exists(LIB_OPS)):
            raise NativeLibraryMissing(
                "native libraries not built: run `python -m ml_music_style_transfer_b200.build` "
                "(needs nvcc with sm_100a support); there is no CPU fallback")
        torch.ops.load_library(LIB_OPS)
        _ops = torch.ops.mst_b200
    return _ops


def cdll():
    """The raw C ABI (include/mst_b200.h) through ctypes -- used by tests and external bindings."""
    global _cdll
    if _cdll is None:
        if not os.path.exists(LIB_CUDA):
            raise NativeLibraryMissing("libmst_b200.so not built; there is no CPU fallback")
        _cdll = ctypes.CDLL(LIB_CUDA)
        _cdll.mst_last_error.restype = ctypes.c_char_p
        _cdll.mst_launch_count.restype = ctypes.c_int64
    return _cdll


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("ml_music_style_transfer_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"device {device} is not a CUDA device; there is no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def launch_count():
    return int(ops().launch_count())
