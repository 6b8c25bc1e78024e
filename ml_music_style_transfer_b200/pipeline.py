"""Host-buffer pipeline over the whole hot path (the call a user with data in host memory makes).

``HostPipeline`` cuts a batch of equal-length clips into chunks that rotate over a few CUDA streams, so that the
pinned-host -> device copies of the next chunks and the device -> host copies of the previous ones overlap the kernels of
the current one.  Per chunk it runs, in the reference's order of use (preprocess.py:163-200, inference.py:74-110):

    audio           -> STFT + log-mel                                   -> host (n, n_mels, T)
    MIDI notes      -> piano roll + on/off at `roll_fs`                  -> host (rows, 128) x 2
                    -> audio-rate planes (n_keys, N) int8 x 2            -> stay on the device (model conditioning),
                                                                           or host when planes_to_host=True
    spectrogram S   -> n_iter Griffin-Lim                                -> host waveforms

Everything is asynchronous on the pipeline's streams; ``run`` returns after a device synchronise.
"""
import os

import numpy as np
import torch

from . import _lib, features as F, pianoroll as PR


class HostPipeline:
    def __init__(self, n_clips, clip_len, sr=22050, hop=512, n_mels=128, roll_fs=250, pitch_lo=21, n_keys=88, gl_iters=32,
                 n_chunks=8, n_streams=None, planes_to_host=False, device=None, plan=None):
        self.device = _lib.require_cuda(device)
        self.n, self.clip_len, self.sr, self.hop = int(n_clips), int(clip_len), int(sr), int(hop)
        self.n_mels, self.roll_fs, self.pitch_lo, self.n_keys, self.gl_iters = n_mels, roll_fs, pitch_lo, n_keys, gl_iters
        self.frames = 1 + self.clip_len // self.hop
        self.wave_len = self.hop * (self.frames - 1)
        self.planes_to_host = planes_to_host
        self.plan = plan if plan is not None else F.MelPlan.get(sr, F.N_FFT, n_mels, device=self.device)
        if n_streams is None:
            n_streams = int(os.environ.get("MST_E2E_STREAMS", "4"))
        self.n_streams = max(1, int(n_streams))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_streams)]
        n_chunks = max(1, min(int(n_chunks), self.n // 256)) if self.n >= 512 else 1
        self.bounds = [(self.n * i) // n_chunks for i in range(n_chunks + 1)]
        self.chunks = []
        for ci in range(n_chunks):
            a0, a1 = self.bounds[ci], self.bounds[ci + 1]
            self.chunks.append(dict(a0=a0, a1=a1, m=a1 - a0,
                                    batch=F.ClipBatch.uniform(a1 - a0, self.clip_len, self.hop, device=self.device),
                                    gl_batch=F.ClipBatch.from_frames([self.frames] * (a1 - a0), self.hop, device=self.device)))
        self.seconds = self.clip_len / float(self.sr)
        # every piece is rasterised over the clip's duration (end_time = clip seconds), so each owns exactly
        # rows_per_clip rows and piece i of the batch sits at h_roll[i * rows_per_clip : (i + 1) * rows_per_clip]
        self.rows_per_clip = int(self.roll_fs * self.seconds)
        self.plane_sub = max(1, min(self.n, int(256 * 4.0 / max(self.seconds, 1e-3))))  # pieces per audio-rate launch
        # pinned host outputs
        self.h_mel = torch.empty(self.n * n_mels * self.frames, dtype=torch.float32).pin_memory()
        self.h_wave = torch.empty(self.n * self.wave_len, dtype=torch.float32).pin_memory()
        self.h_roll = torch.empty((self.n * self.rows_per_clip, 128), dtype=torch.uint8).pin_memory()
        self.h_onoff = torch.empty((self.n * self.rows_per_clip, 128), dtype=torch.int8).pin_memory()
        # audio-rate planes are 15.5 MB per 4 s clip: the host side is a staging ring with ONE slot per stream (copies on
        # a stream are ordered, so a slot is never written by two streams); a consumer drains a slot before the stream's
        # next sub-batch lands.  This measures the transfer; it does not retain 250 GB of planes.
        self.h_planes = ([torch.empty(2 * self.plane_sub * n_keys * self.clip_len, dtype=torch.int8).pin_memory()
                          for _ in range(self.n_streams)] if planes_to_host else None)
        self.d2h_roll_bytes = 0
        self.h2d_note_bytes = 0
        self._note_cap = 0

    def _stage_notes(self, notes):
        """notes = (pitch, velocity, start, end, note_offsets) host arrays for the n pieces -> pinned staging buffers.
        Runs on EVERY call (it is part of the host -> device path): four memcpys into page-locked memory plus the
        per-chunk offset arrays."""
        pitch, vel, start, end, off = notes
        off = np.asarray(off, dtype=np.int64)
        total = int(off[self.n])
        if total > self._note_cap:
            cap = int(total * 1.25) + 1024
            self._pins = [torch.empty(cap, dtype=dt).pin_memory() for dt in (torch.int32, torch.int32, torch.float64, torch.float64)]
            self._pin_off = torch.empty(self.n + len(self.chunks) + 1, dtype=torch.int64).pin_memory()
            self._note_cap = cap
        for t, a in zip(self._pins, (pitch, vel, start, end)):
            np.copyto(t.numpy()[:total], np.asarray(a)[:total], casting="same_kind")
        o = 0
        for ch in self.chunks:
            a0, a1 = ch["a0"], ch["a1"]
            n0, n1 = int(off[a0]), int(off[a1])
            ch["notes"] = [t[n0:n1] for t in self._pins]
            dst = self._pin_off[o:o + (a1 - a0 + 1)]
            np.subtract(off[a0:a1 + 1], off[a0], out=dst.numpy())
            ch["noff"] = dst
            o += a1 - a0 + 1
        self.h2d_note_bytes = total * (4 + 4 + 8 + 8) + (self.n + len(self.chunks)) * 8

    def run(self, h_audio, h_S, notes, seed=7):
        """h_audio: pinned float32 [n * clip_len]; h_S: pinned float32 frame-major magnitudes [n * frames * 1025];
        notes: host SoA arrays.  Results land in self.h_mel / h_wave / h_roll / h_onoff (and the h_planes ring)."""
        self._stage_notes(notes)
        dev, K = self.device, F.N_BINS
        main = torch.cuda.current_stream()
        for s_ in self.streams:
            s_.wait_stream(main)
        self.d2h_roll_bytes = 0
        for ci, ch in enumerate(self.chunks):
            a0, a1, m = ch["a0"], ch["a1"], ch["m"]
            si = ci % self.n_streams
            with torch.cuda.stream(self.streams[si]):
                a = h_audio[a0 * self.clip_len:a1 * self.clip_len].to(dev, non_blocking=True)
                mel = F.melspectrogram_batch(a, ch["batch"], self.plan, log1p=True, layout=F.BIN_MAJOR)
                self.h_mel[a0 * self.n_mels * self.frames:a1 * self.n_mels * self.frames].copy_(mel, non_blocking=True)
                nb = PR.NoteBatch.from_host_tensors(*ch["notes"], ch["noff"], np.full(m, self.seconds), device=dev)
                roll, onoff, row_off, _ = PR.rasterize(nb, self.roll_fs)
                rows = m * self.rows_per_clip
                assert roll.shape[0] == rows
                r0 = a0 * self.rows_per_clip
                self.h_roll[r0:r0 + rows].copy_(roll, non_blocking=True)
                self.h_onoff[r0:r0 + rows].copy_(onoff, non_blocking=True)
                self.d2h_roll_bytes += 2 * rows * 128
                for s in range(0, m, self.plane_sub):
                    e = min(m, s + self.plane_sub)
                    ro = row_off[s:e + 1]
                    ua, ub, _ = PR.upsample_pair(roll, onoff, ro, self.clip_len, self.roll_fs, self.sr, self.pitch_lo,
                                                 self.n_keys, torch.int8)
                    if self.planes_to_host:
                        k = (e - s) * self.n_keys * self.clip_len
                        self.h_planes[si][:k].copy_(ua, non_blocking=True)
                        self.h_planes[si][k:2 * k].copy_(ub, non_blocking=True)
                Sd = h_S[a0 * self.frames * K:a1 * self.frames * K].to(dev, non_blocking=True)
                y = F.griffinlim_batch(Sd, ch["gl_batch"], n_iter=self.gl_iters, momentum=0.99, init="random", seed=seed,
                                       layout=F.FRAME_MAJOR)
                self.h_wave[a0 * self.wave_len:a1 * self.wave_len].copy_(y, non_blocking=True)
        for s_ in self.streams:
            main.wait_stream(s_)
        torch.cuda.synchronize(dev)
        return self

    def bytes_per_run(self):
        h2d = self.n * self.clip_len * 4 + self.n * self.frames * F.N_BINS * 4 + self.h2d_note_bytes
        d2h = self.h_mel.numel() * 4 + self.h_wave.numel() * 4 + self.d2h_roll_bytes
        if self.planes_to_host:
            d2h += 2 * self.n * self.n_keys * self.clip_len
        return int(h2d), int(d2h)


class DevicePipeline:
    """The same chunked, multi-stream schedule for inputs that already live in HBM and results that stay there.

    The three stages have different bottlenecks (log-mel: FP32 pipe; audio-rate rolls: HBM writes; Griffin-Lim: shared
    memory / latency), so running chunk k's stages while chunk k-1's Griffin-Lim is still in flight -- chunks rotate over
    `n_streams` CUDA streams -- fills units a single stream leaves idle.  ``run`` returns after joining the streams on
    the caller's stream (no host synchronisation).  With ``collect=True`` the log-mel, waveforms and frame-rate rolls of
    every chunk are gathered into full-size device tensors (``self.mel / wave / roll / onoff``); the audio-rate planes
    (15.5 MB per 4 s clip) are handed to ``plane_consumer(chunk_start, chunk_stop, up_roll, up_onoff)`` chunk by chunk.
    """

    def __init__(self, n_clips, clip_len, notes, sr=22050, hop=512, n_mels=128, roll_fs=250, pitch_lo=21, n_keys=88,
                 gl_iters=32, n_chunks=32, n_streams=4, collect=False, plane_consumer=None, device=None, plan=None):
        self.device = _lib.require_cuda(device)
        self.n, self.clip_len, self.sr, self.hop = int(n_clips), int(clip_len), int(sr), int(hop)
        self.n_mels, self.roll_fs, self.pitch_lo, self.n_keys, self.gl_iters = n_mels, roll_fs, pitch_lo, n_keys, gl_iters
        self.frames = 1 + self.clip_len // self.hop
        self.wave_len = self.hop * (self.frames - 1)
        self.seconds = self.clip_len / float(self.sr)
        self.rows_per_clip = int(self.roll_fs * self.seconds)
        self.plan = plan if plan is not None else F.MelPlan.get(sr, F.N_FFT, n_mels, device=self.device)
        self.n_streams = max(1, int(n_streams))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_streams)]
        n_chunks = max(1, min(int(n_chunks), self.n // 256)) if self.n >= 512 else 1
        bounds = [(self.n * i) // n_chunks for i in range(n_chunks + 1)]
        self.plane_sub = max(1, min(self.n, int(256 * 4.0 / max(self.seconds, 1e-3))))
        pitch, vel, start, end, off = notes
        off = np.asarray(off, dtype=np.int64)
        self.chunks = []
        for ci in range(n_chunks):
            a0, a1 = bounds[ci], bounds[ci + 1]
            n0, n1 = int(off[a0]), int(off[a1])
            nb = PR.NoteBatch(pitch[n0:n1], vel[n0:n1], start[n0:n1], end[n0:n1], off[a0:a1 + 1] - off[a0],
                              device=self.device, end_times=[self.seconds] * (a1 - a0))
            self.chunks.append(dict(a0=a0, a1=a1, m=a1 - a0, notes=nb,
                                    batch=F.ClipBatch.uniform(a1 - a0, self.clip_len, self.hop, device=self.device),
                                    gl_batch=F.ClipBatch.from_frames([self.frames] * (a1 - a0), self.hop, device=self.device)))
        self.collect, self.plane_consumer = bool(collect), plane_consumer
        if self.collect:
            dev = self.device
            self.mel = torch.empty(self.n * n_mels * self.frames, dtype=torch.float32, device=dev)
            self.wave = torch.empty(self.n * self.wave_len, dtype=torch.float32, device=dev)
            self.roll = torch.empty((self.n * self.rows_per_clip, 128), dtype=torch.uint8, device=dev)
            self.onoff = torch.empty((self.n * self.rows_per_clip, 128), dtype=torch.int8, device=dev)

    def run(self, audio, S, seed=7):
        """audio: CUDA float32 [n * clip_len]; S: CUDA float32 frame-major magnitudes [n * frames * 1025]."""
        K = F.N_BINS
        main = torch.cuda.current_stream(self.device)
        for s_ in self.streams:
            s_.wait_stream(main)
        for ci, ch in enumerate(self.chunks):
            a0, a1, m = ch["a0"], ch["a1"], ch["m"]
            with torch.cuda.stream(self.streams[ci % self.n_streams]):
                mel = F.melspectrogram_batch(audio[a0 * self.clip_len:a1 * self.clip_len], ch["batch"], self.plan, log1p=True,
                                             layout=F.BIN_MAJOR)
                roll, onoff, row_off, _ = PR.rasterize(ch["notes"], self.roll_fs)
                for s in range(0, m, self.plane_sub):
                    e = min(m, s + self.plane_sub)
                    ua, ub, _ = PR.upsample_pair(roll, onoff, row_off[s:e + 1], self.clip_len, self.roll_fs, self.sr,
                                                 self.pitch_lo, self.n_keys, torch.int8)
                    if self.plane_consumer is not None:
                        self.plane_consumer(a0 + s, a0 + e, ua, ub)
                y = F.griffinlim_batch(S[a0 * self.frames * K:a1 * self.frames * K], ch["gl_batch"], n_iter=self.gl_iters,
                                       momentum=0.99, init="random", seed=seed, layout=F.FRAME_MAJOR)
                if self.collect:
                    self.mel[a0 * self.n_mels * self.frames:a1 * self.n_mels * self.frames].copy_(mel)
                    self.wave[a0 * self.wave_len:a1 * self.wave_len].copy_(y)
                    r0 = a0 * self.rows_per_clip
                    self.roll[r0:r0 + m * self.rows_per_clip].copy_(roll)
                    self.onoff[r0:r0 + m * self.rows_per_clip].copy_(onoff)
                # the chunk's temporaries were allocated on this side stream: the caching allocator may hand them to the
                # next chunk of the SAME stream only, which is ordered after this one
        for s_ in self.streams:
            main.wait_stream(s_)
        return self
