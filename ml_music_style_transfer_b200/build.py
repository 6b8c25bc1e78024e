"""In-tree build of the two native libraries (no JIT cache, so the .so files travel with the repo):

  libmst_b200.so       nvcc, sm_100a only: every CUDA kernel + the C ABI of include/mst_b200.h
  libmst_torch_ops.so  g++: torch.library custom ops (namespace ``mst_b200``) forwarding to the C ABI

Run as ``python -m ml_music_style_transfer_b200.build`` or through ``__graft_entry__.build()``.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CUDA_SOURCES = ["core.cu", "stft.cu", "mel_gemm.cu", "mel_inverse.cu", "griffinlim.cu", "generic_fft.cu", "pianoroll.cu", "resample.cu"]
HEADERS = ["mst_common.cuh", "fft_warp.cuh", "mel_plan.cuh", os.path.join("..", "..", "include", "mst_b200.h")]
LIB_CUDA = os.path.join(HERE, "libmst_b200.so")
LIB_OPS = os.path.join(HERE, "libmst_torch_ops.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, stamp_value):
    stamp = target + ".stamp"
    if not (os.path.exists(target) and os.path.exists(stamp)):
        return True
    with open(stamp) as f:
        return f.read().strip() != stamp_value


def _write_stamp(target, stamp_value):
    with open(target + ".stamp", "w") as f:
        f.write(stamp_value)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    stamp = _digest(deps, " ".join(flags))
    if not force and not _stale(LIB_CUDA, stamp):
        return LIB_CUDA
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(CSRC, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_CUDA] + objs + ["-cudart", "static"]
    subprocess.check_call(cmd)
    _write_stamp(LIB_CUDA, stamp)
    return LIB_CUDA


def build_ops(force=False):
    import torch
    from torch.utils import cpp_extension
    src = os.path.join(CSRC, "torch_ops.cpp")
    stamp = _digest([src, os.path.join(HERE, "..", "include", "mst_b200.h")], torch.__version__)
    if not force and not _stale(LIB_OPS, stamp):
        return LIB_OPS
    inc = []
    for p in cpp_extension.include_paths("cuda") if "device_type" in cpp_extension.include_paths.__code__.co_varnames \
            else cpp_extension.include_paths(True):
        inc += ["-I", p]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_API_INCLUDE_EXTENSION_H",
           src, "-o", LIB_OPS] + inc + ["-I", "/usr/local/cuda/include",
           "-L", torch_lib, "-L", HERE, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-lmst_b200",
           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + torch_lib, "-Wl,--no-as-needed"]
    subprocess.check_call(cmd)
    _write_stamp(LIB_OPS, stamp)
    return LIB_OPS


def build_tools(force=False):
    """tools/ubench: standalone microbenchmarks (write bandwidth ceiling, FFMA2 issue rate) quoted in DESIGN.md."""
    src = os.path.join(HERE, "..", "tools", "ubench.cu")
    out = os.path.join(HERE, "..", "tools", "ubench")
    if not os.path.exists(src):
        return None
    stamp = _digest([src])
    if not force and not _stale(out, stamp):
        return out
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                           "-o", out, src])
    _write_stamp(out, stamp)
    return out


def build_all(force=False, verbose=False):
    build_cuda(force=force, verbose=verbose)
    build_ops(force=force)
    build_tools(force=force)
    return LIB_CUDA, LIB_OPS


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB_CUDA)
    print(LIB_OPS)
