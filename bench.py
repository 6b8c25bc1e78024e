#!/usr/bin/env python
"""Benchmark of the preprocessing / inversion hot path (BASELINE.json metric: audio-seconds per second).

One "step" = one pass of the whole hot path over one batch of synthetic clips (config C4 of BASELINE.json:
16384 clips x 4 s @ 22.05 kHz, n_fft 2048, hop 512, 128 mels), per GPU:

    stage A  STFT + log-mel of every clip                         (P1 + P2)
    stage B  MIDI notes -> 88-key onset piano roll @ 250 Hz -> audio-rate int8 planes (roll + onoff)   (P3)
    stage C  32-iteration Griffin-Lim of every clip's magnitude spectrogram                             (P4)

value = audio seconds in the batch / (time of A + B + C): the throughput of the full path; the per-stage figures and
their rooflines are in "stages".  Inputs are resident in HBM for `value`; `e2e` repeats the step through the NumPy-facing
public API with pinned HOST buffers (H2D of the audio / spectrograms / notes and D2H of every result inside the timed
region).  N > 1: one process per GPU (torchrun), every rank runs its own shard of the same size (weak scaling), no
data-path collective; time is the max over ranks.

--impl reference times the CPU restatement of the reference's path (oracle/, NumPy float64 pocketfft -- librosa and
pretty_midi are not installable in this image) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, N_FFT, HOP, N_MELS, K = 22050, 2048, 512, 128, 1025
CLIP_SECONDS = 4.0
CLIP_LEN = int(SR * CLIP_SECONDS)          # 88200
T_FRAMES = 1 + CLIP_LEN // HOP             # 173
GL_ITERS = 32
ROLL_FS, PITCH_LO, N_KEYS = 250, 21, 88
METRIC = "audio-seconds/sec, STFT+log-mel + piano-roll + 32-iter Griffin-Lim (full hot path)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------------------------
def make_notes(n_pieces, seed0):
    """Host note arrays for n_pieces 4-second pieces (12 notes/s Poisson; SURVEY 8d C3 distribution)."""
    from ml_music_style_transfer_b200 import synth
    pool = [synth.midi_piece(seed0 + i, seconds=CLIP_SECONDS) for i in range(min(n_pieces, 256))]
    pieces = [pool[i % len(pool)] for i in range(n_pieces)]
    offs = np.zeros(n_pieces + 1, dtype=np.int64)
    np.cumsum([len(p[0]) for p in pieces], out=offs[1:])
    cat = lambda j, dt: np.concatenate([p[j] for p in pieces]).astype(dt)
    return cat(0, np.int32), cat(1, np.int32), cat(2, np.float64), cat(3, np.float64), offs


def make_audio_device(n_clips, device, seed0):
    """Piano-like pool (host, seeded) tiled over the batch with per-clip gain + device noise floor."""
    import torch
    from ml_music_style_transfer_b200 import synth
    pool = np.stack([synth.piano_clip(seed0 + i, CLIP_SECONDS, SR)[:CLIP_LEN] for i in range(64)])
    pool_d = torch.from_numpy(pool).to(device)
    g = torch.Generator(device=device).manual_seed(1234 + seed0)
    audio = torch.empty(n_clips * CLIP_LEN, dtype=torch.float32, device=device)
    view = audio.view(n_clips, CLIP_LEN)
    for s in range(0, n_clips, 1024):
        e = min(n_clips, s + 1024)
        idx = torch.arange(s, e, device=device) % 64
        gain = 0.5 + 0.5 * torch.rand(e - s, 1, generator=g, device=device)
        view[s:e] = pool_d[idx] * gain + 1e-3 * torch.randn(e - s, CLIP_LEN, generator=g, device=device)
    return audio


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ml_music_style_transfer_b200 as pkg
    from ml_music_style_transfer_b200 import features as F, pianoroll as PR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pkg._lib.ops()
    hbm_peak, peak_src = peaks()

    n_clips = args.clips                      # per GPU (weak scaling)
    audio_seconds = n_clips * CLIP_SECONDS
    audio = make_audio_device(n_clips, device, seed0=1000 * rank)
    batch = F.ClipBatch.uniform(n_clips, CLIP_LEN, HOP, device=device)
    gl_sub = min(n_clips, args.gl_sub)        # clips per Griffin-Lim call (bounds the HBM workspace: 20 B/bin + accumulators)
    gl_batches = [(s0, min(n_clips, s0 + gl_sub)) for s0 in range(0, n_clips, gl_sub)]
    gl_batches = [(a0, a1, F.ClipBatch.from_frames([T_FRAMES] * (a1 - a0), HOP, device=device)) for a0, a1 in gl_batches]
    plan = F.MelPlan.get(SR, N_FFT, N_MELS, device=device)
    notes_h = make_notes(n_clips, seed0=99 + 1000 * rank)
    roll_sub = max(1, min(n_clips, int(1024 * 4.0 / CLIP_SECONDS)))   # pieces per piano-roll launch (ring of output buffers)
    notes = PR.NoteBatch(*notes_h, device=device)
    # GL input: the magnitude spectrogram the model would emit (made once, untimed)
    S = F.stft_batch(audio, batch, "magnitude", F.FRAME_MAJOR)
    torch.cuda.synchronize()

    def stage_a():
        return F.melspectrogram_batch(audio, batch, plan, log1p=True, layout=F.BIN_MAJOR)

    def stage_b():
        roll, onoff, row_off, _ = PR.rasterize(notes, ROLL_FS)
        outs = None
        for s in range(0, n_clips, roll_sub):
            e = min(n_clips, s + roll_sub)
            ro = row_off[s:e + 1]
            a, b, _ = PR.upsample_pair(roll, onoff, ro, CLIP_LEN, ROLL_FS, SR, PITCH_LO, N_KEYS, torch.int8)
            outs = (a, b)
        return outs

    def stage_c(n_iter=GL_ITERS):
        y = None
        for a0, a1, gb in gl_batches:
            y = F.griffinlim_batch(S[a0 * T_FRAMES * K:a1 * T_FRAMES * K], gb, n_iter=n_iter, momentum=0.99,
                                   init_phase=None, init="random", seed=7, layout=F.FRAME_MAJOR)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    # The timed step runs the three stages through the package's DevicePipeline: the batch is cut into chunks that rotate
    # over 4 CUDA streams, so the write-bound roll writer and the FP32-bound log-mel of one chunk run under the
    # latency-bound Griffin-Lim of another (same kernels, same work, every launch counted).  --serial times the three
    # stages back to back on one stream instead; `stages` below are always measured that way, each stage alone.
    dpipe = None
    if not args.serial:
        from ml_music_style_transfer_b200.pipeline import DevicePipeline
        dpipe = DevicePipeline(n_clips, CLIP_LEN, notes_h, sr=SR, hop=HOP, n_mels=N_MELS, roll_fs=ROLL_FS, pitch_lo=PITCH_LO,
                               n_keys=N_KEYS, gl_iters=GL_ITERS, n_chunks=args.step_chunks, n_streams=args.step_streams, device=device,
                               plan=plan)

    def step():
        if dpipe is not None:
            dpipe.run(audio, S)
        else:
            stage_a(); stage_b(); stage_c()

    # warm-up (W >= 3 steps)
    for _ in range(max(3, args.warmup)):
        step()
    for _ in range(2):
        stage_a(); stage_b(); stage_c()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = pkg._lib.launch_count()
    ta, tb, tc = [], [], []
    barrier()
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    wall_s = time.perf_counter() - t_wall0
    launches = pkg._lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-stage figures: each stage alone on one stream (not part of `value`)
    for _ in range(max(3, min(args.steps, 5))):
        ta += timed(stage_a, 1); tb += timed(stage_b, 1); tc += timed(stage_c, 1)

    # reference point for the write-only stage: what the driver's own device memset achieves on this GPU (measured live;
    # tools/ubench.cu has the full study: memset 7.4 TB/s, one-shot st.global.v4 grid 7.6, grid-stride 6.2-6.9, bulk
    # (TMA) stores 6.4 -- torch's fill_ kernel, which round 1 took for the write ceiling, only reaches 3.9)
    fill_buf = torch.empty(2 * 1024 ** 3, dtype=torch.int8, device=device)
    try:
        from cuda.bindings import runtime as cudart   # cuda-python: the driver's memset, not a torch kernel
    except Exception:  # noqa: BLE001
        cudart = None

    def _memset():
        if cudart is not None:
            cudart.cudaMemsetAsync(fill_buf.data_ptr(), 1, fill_buf.numel(), torch.cuda.current_stream().cuda_stream)
        else:
            fill_buf.fill_(1)
    _memset()
    fill_ms = float(np.median(timed(_memset, 5)))
    write_peak = fill_buf.numel() / (fill_ms * 1e-3) / 1e9
    del fill_buf

    # kernel-level roofline of the dominant kernel (Griffin-Lim iteration): (t[32 iters] - t[0 iters]) / 32
    med = lambda x: float(np.median(x))
    t0 = float(np.median(timed(lambda: stage_c(0), 3)))
    t32 = float(np.median(tc))
    iter_ms = (t32 - t0) / GL_ITERS
    L = HOP * (T_FRAMES - 1)
    gl_iter_bytes = n_clips * (36 * K * T_FRAMES + 8 * L)              # SURVEY 8d, per launch
    gl_total_bytes = n_clips * (GL_ITERS * (36 * K * T_FRAMES + 8 * L) + 12 * K * T_FRAMES + 4 * L)
    a_bytes = n_clips * (4 * CLIP_LEN + 4 * N_MELS * T_FRAMES)
    b_bytes = n_clips * 2 * N_KEYS * CLIP_LEN
    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, med(ta), med(tb), med(tc), iter_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, ma, mb, mc, iter_ms = [float(x) for x in t.tolist()]
    else:
        ma, mb, mc = med(ta), med(tb), med(tc)
    ms_per_step = total_ms / args.steps
    total_audio = world * audio_seconds
    if world > 1:
        ta_ = torch.tensor([audio_seconds], device=device, dtype=torch.float64)
        dist.all_reduce(ta_, op=dist.ReduceOp.SUM)
        total_audio = float(ta_.item())
    value = total_audio / (ms_per_step * 1e-3)

    # ---- optional final gather of the features over NCCL (north_star: "NCCL over NVLink used only for the optional
    # final gather"); measured separately, never part of `value` -------------------------------------------------------
    gather = None
    if world > 1 and not args.no_gather:
        from ml_music_style_transfer_b200 import sharding
        feats = stage_a().view(n_clips, N_MELS * T_FRAMES)
        sharding.gather_features(feats, n_clips * world)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = sharding.gather_features(feats, n_clips * world)
        g1.record()
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], device=device, dtype=torch.float64)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        ok = bool(torch.equal(full[rank * n_clips:(rank + 1) * n_clips], feats))
        gather = {"ms": float(gms.item()), "bytes_received_per_rank": int(full.numel() * 4), "own_shard_intact": ok,
                  "algbw_gbs": full.numel() * 4 / (float(gms.item()) * 1e-3) / 1e9}
        del full, feats

    # ---- strong scaling (N > 1): the SAME 16384-clip C4 batch split over the ranks (the weak line above gives every rank
    # its own 16384 clips); no data-path collective either way, so this measures launch / tail overheads at small shards ----
    strong = None
    if world > 1 and not args.no_strong and args.workload == "c4":
        from ml_music_style_transfer_b200 import sharding
        s0, s1 = sharding.shard_range(n_clips, rank, world)
        m = s1 - s0
        batch_s = F.ClipBatch.uniform(m, CLIP_LEN, HOP, device=device)
        gl_s = F.ClipBatch.from_frames([T_FRAMES] * m, HOP, device=device)
        offs = notes_h[4]
        notes_s = PR.NoteBatch(*[a[:offs[m]] for a in notes_h[:4]], offs[:m + 1], device=device)

        def step_s():
            F.melspectrogram_batch(audio[:m * CLIP_LEN], batch_s, plan, log1p=True, layout=F.BIN_MAJOR)
            roll, onoff, row_off, _ = PR.rasterize(notes_s, ROLL_FS)
            for q0 in range(0, m, roll_sub):
                q1 = min(m, q0 + roll_sub)
                PR.upsample_pair(roll, onoff, row_off[q0:q1 + 1], CLIP_LEN, ROLL_FS, SR, PITCH_LO, N_KEYS, torch.int8)
            return F.griffinlim_batch(S[:m * T_FRAMES * K], gl_s, n_iter=GL_ITERS, momentum=0.99, init="random", seed=7,
                                      layout=F.FRAME_MAJOR)
        if not args.serial:
            sp = DevicePipeline(m, CLIP_LEN, tuple(a[:offs[m]] for a in notes_h[:4]) + (offs[:m + 1],), sr=SR, hop=HOP,
                                n_mels=N_MELS, roll_fs=ROLL_FS, pitch_lo=PITCH_LO, n_keys=N_KEYS, gl_iters=GL_ITERS,
                                n_chunks=max(1, args.step_chunks // world), n_streams=args.step_streams, device=device, plan=plan)
            step_s = lambda: sp.run(audio[:m * CLIP_LEN], S[:m * T_FRAMES * K])   # noqa: E731
        for _ in range(3):
            step_s()
        barrier()
        k_s = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k_s):
            step_s()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / k_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "total_clips": n_clips, "clips_per_gpu": m, "steps": k_s, "ms_per_step": float(t.item()),
                  "value": n_clips * CLIP_SECONDS / (float(t.item()) * 1e-3), "unit": "audio-s/s"}
        del batch_s, gl_s, notes_s

    # ---- configs[0]/[1] of BASELINE.json: ONE 30 s clip (latency view; tiny against a B200, reported for completeness) ----
    single = None
    if rank == 0 and not args.no_single:
        n30 = 30 * SR
        a30 = make_audio_device(max(1, -(-n30 // CLIP_LEN)), device, 77)[:n30].contiguous()
        b30 = F.ClipBatch.uniform(1, n30, HOP, device=device)
        b30r = F.ClipBatch.uniform(1, n30, 256, device=device)   # the repo's own hop (preprocess.py:40)
        T30 = b30.total_frames
        g30 = F.ClipBatch.from_frames([T30], HOP, device=device)
        S30 = F.stft_batch(a30, b30, "magnitude", F.FRAME_MAJOR)
        t_lm = med(timed(lambda: F.melspectrogram_batch(a30, b30, plan, log1p=True, layout=F.BIN_MAJOR), 20))
        t_lp = med(timed(lambda: F.stft_batch(a30, b30r, "log1p_power", F.FRAME_MAJOR), 20))
        t_gl = med(timed(lambda: F.griffinlim_batch(S30, g30, n_iter=GL_ITERS, seed=3, layout=F.FRAME_MAJOR), 10))
        single = {"logmel_hop512_ms": t_lm, "log1p_power_hop256_ms": t_lp, "griffinlim32_hop512_ms": t_gl,
                  "griffinlim32_audio_s_per_s": 30.0 / (t_gl * 1e-3), "frames": T30}

    # ---- "next" rows and the reference's own geometry: measured, not part of `value` -------------------------
    extras = None
    if rank == 0 and not args.no_single:
        from ml_music_style_transfer_b200 import _lib as L
        x44 = torch.randn(64 * 8 * 44100, device=device) * 0.1           # 64 clips x 8 s at 44.1 kHz, one buffer
        t_rs = med(timed(lambda: L.ops().resample(x44, 44100, 22050), 5))
        taps = 2 * 64 * 2                                                 # both wings, 64 zero crossings, ratio 1/2
        t_rs48 = med(timed(lambda: L.ops().resample(x44, 48000, 44100), 5))  # rational ratio 160/147: per-phase weight tables
        n_ch, step_, clen = 256, 131072, 219904                           # preprocess.py:66-67 chunk geometry
        a_repo = torch.randn((n_ch - 1) * step_ + clen, device=device) * 0.1
        b_repo = F.ClipBatch.uniform(n_ch, clen, 256, clip_stride=step_, device=device)
        t_repo = med(timed(lambda: F.stft_batch(a_repo, b_repo, "log1p_power", F.BIN_MAJOR), 5))
        repo_bytes = n_ch * 4 * clen + 4 * K * b_repo.total_frames
        extras = {
            "resample_44k1_to_22k05": {"ms": t_rs, "audio_s_per_s": 64 * 8 / (t_rs * 1e-3),
                                       "gflops": 2.0 * taps * 2 * (x44.numel() // 2) / (t_rs * 1e-3) / 1e9},
            "resample_48k_to_44k1": {"ms": t_rs48, "audio_s_per_s": x44.numel() / 48000 / (t_rs48 * 1e-3)},
            "reference_geometry_log1p_power": {"chunks": n_ch, "ms": t_repo, "audio_s_per_s": n_ch * clen / 44100 / (t_repo * 1e-3),
                                               "hbm_frac": repo_bytes / (t_repo * 1e-3) / 1e9 / hbm_peak,
                                               "note": "44.1 kHz, hop 256, 219904-sample chunks every 131072, bin-major stack"}}
        del x44, a_repo

    # ---- e2e: NumPy-facing API with pinned host buffers ------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, pkg, F, PR, device, audio, S, notes_h, plan, world, barrier)
        try:   # the 15 MB-per-clip variant is informative only (and 62 GB of D2H per 4096 clips): fewer clips, never fatal
            e2e_all = run_e2e(args, pkg, F, PR, device, audio, S, notes_h, plan, world, barrier, planes_to_host=True,
                              n=min(args.e2e_clips, 2048))
        except Exception as exc:  # noqa: BLE001
            e2e_all = {"error": repr(exc)}

    if rank == 0:
        props = torch.cuda.get_device_properties(device)
        fp32_peak = props.multi_processor_count * 128 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6
        stages = {
            # this stage is bound by fp32 issue, not HBM -- report both rooflines (fp32 peak = SMs x 128 lanes x 2
            # x max SM clock; algorithmic flops = T x (2.5 F log2 F + 3 K) for the FFT + magnitude, GEMM flops not counted)
            "stft_logmel": {"ms": ma, "audio_s_per_s": total_audio / (ma * 1e-3),
                            "hbm_frac": a_bytes / (ma * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": a_bytes,
                            "fp32_frac": n_clips * T_FRAMES * 59395.0 / (ma * 1e-3) / fp32_peak, "fp32_peak_tflops": fp32_peak / 1e12},
            "pianoroll_upsample": {"ms": mb, "audio_s_per_s": total_audio / (mb * 1e-3),
                                   "hbm_frac": b_bytes / (mb * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": b_bytes,
                                   "write_only_peak_gbs": write_peak, "write_only_frac": b_bytes / (mb * 1e-3) / 1e9 / write_peak},
            "griffinlim32": {"ms": mc, "audio_s_per_s": total_audio / (mc * 1e-3),
                             "hbm_frac": gl_total_bytes / (mc * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": gl_total_bytes,
                             "fp32_frac": n_clips * T_FRAMES * 133140.0 * GL_ITERS / (mc * 1e-3) / fp32_peak},
        }
        achieved = gl_iter_bytes / (iter_ms * 1e-3) / 1e9
        traffic, traffic_src = args.traffic, "--traffic"
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if traffic is None and os.path.exists(tpath):
            with open(tpath) as f:  # ncu: dram__bytes_read.sum + dram__bytes_write.sum of one GL iteration launch
                tj = json.load(f)
            # captured on 173-frame clips: scale per FRAME when this run's geometry differs (the kernel's traffic is per frame)
            per_frame = tj["gl_iteration_dram_bytes_per_clip"] / float(tj.get("frames_per_clip", 173))
            traffic = per_frame * T_FRAMES * n_clips
            same = tj.get("clips") == n_clips and tj.get("frames_per_clip", 173) == T_FRAMES
            traffic_src = tj.get("source", "profiles/traffic.json") + (
                "" if same else f" -- scaled per frame from {tj.get('clips')} x {tj.get('frames_per_clip', 173)} to "
                                f"{n_clips} x {T_FRAMES} frames")
        line = {
            "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c5" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload.upper()}: {n_clips} clips/GPU x {CLIP_SECONDS:g} s @ 22.05 kHz, n_fft 2048, "
                                   f"hop 512, 128 mels; 88-key roll @ 250 Hz -> audio rate (int8, roll+onoff); "
                                   f"Griffin-Lim {GL_ITERS} it",
                       "clips_per_gpu": n_clips,
                       "l2_policy": f"inputs larger than L2 ({n_clips * CLIP_LEN * 4 / 1e9:.1f} GB audio, "
                                    f"{n_clips * T_FRAMES * K * 4 / 1e9:.1f} GB spectrogram per GPU)",
                       "parallelism": f"clips sharded over {world} GPU(s), no data-path collective",
                       "step_schedule": ("serial: stages back to back on one stream" if dpipe is None else
                                         f"pipeline.DevicePipeline: {len(dpipe.chunks)} chunks over {dpipe.n_streams} streams "
                                         "(stages of different chunks overlap; `stages` are each stage alone)")},
            "stages_serial_ms": ma + mb + mc,
            "stages": stages,
            "roofline": {"bound": "hbm", "kernel": "gl_kernel<false> (one Griffin-Lim iteration)", "achieved": achieved,
                         "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "ms_per_launch": iter_ms, "algorithmic_bytes_per_launch": gl_iter_bytes, "traffic": traffic,
                         # the kernel never stores the phase, so its real DRAM traffic is below the algorithmic figure:
                         # dram_frac = measured DRAM bytes / time / peak is the honest utilisation of the HBM pipe
                         "dram_frac": (traffic / (iter_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                         "traffic_source": traffic_src},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall_s,
        }
        if not args.no_single:
            line["single_clip_30s"] = single
            line["extras"] = extras
        if gather is not None:
            line["final_gather_logmel"] = gather
        if strong is not None:
            line["strong_scaling"] = strong
        if e2e is not None:
            line["e2e"] = e2e
            line["e2e_planes_to_host"] = e2e_all
        if not args.no_cpu_baseline and world == 1:   # reported baseline: rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_baseline(cores=1, budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def host_memory_available():
    """Bytes of host memory this process may still use: MemAvailable, bounded by the cgroup limit when there is one."""
    avail = 64 << 30
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    avail = int(line.split()[1]) * 1024
    except OSError:
        pass
    for path, cur in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                      ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            with open(path) as f:
                lim = f.read().strip()
            if lim != "max" and int(lim) < (1 << 60):
                with open(cur) as f:
                    used = int(f.read().strip())
                avail = min(avail, int(lim) - used)
        except (OSError, ValueError):
            pass
    return max(avail, 1 << 30)


def host_copy_ceiling(device, world, barrier, n_streams=4, mib=256, reps=3):
    """What the host side of this box gives ONE rank while ALL ranks copy at once, with the pipeline's own pattern:
    `n_streams` streams, each alternating a pinned H2D copy and a pinned D2H copy of `mib` MiB.  Returns the combined
    (both directions) GB/s of the slowest rank.  At N = 8 the eight GPUs share the host's memory system / PCIe root, so
    this -- not the kernels, not NVLink -- bounds `e2e`."""
    import torch
    n = int(mib) << 20
    bufs = [(torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
             torch.empty(n, dtype=torch.uint8, device=device), torch.ones(n, dtype=torch.uint8, device=device),
             torch.cuda.Stream(device=device)) for _ in range(n_streams)]
    main = torch.cuda.current_stream()

    def go(k):
        for h_in, h_out, d_in, d_out, st in bufs:
            st.wait_stream(main)
            with torch.cuda.stream(st):
                for _ in range(k):
                    d_in.copy_(h_in, non_blocking=True)
                    h_out.copy_(d_out, non_blocking=True)
        for *_, st in bufs:
            main.wait_stream(st)
    go(1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    go(reps)
    e1.record()
    barrier()
    gbs = 2.0 * reps * n_streams * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([gbs], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)   # the slowest rank's share
        gbs = float(t.item())
    return gbs


def run_e2e(args, pkg, F, PR, device, audio_d, S_d, notes_h, plan, world, barrier, planes_to_host=False, n=None):
    """Same step through host buffers via the package's public ``pipeline.HostPipeline``: pinned host -> device copies
    and device -> host reads are inside the timing, and so is the staging of the MIDI note arrays into pinned memory.

    Host inputs: audio, magnitude spectrograms (the model output Griffin-Lim inverts), MIDI note arrays -- FRESH per
    repetition: every timed run slides a window over a host buffer that holds `reps` extra clips, so no two runs copy the
    same bytes to the same place.  Host outputs: log-mel, reconstructed waveforms, frame-rate piano roll + on/off (what the
    reference's load_midi returns).  The audio-rate int8 planes are the model's conditioning input and stay on the device
    by default (15.5 MB per 4 s clip, 45x the audio itself); `planes_to_host=True` also streams them out (reported
    separately, on fewer clips).  Chunks rotate over 4 CUDA streams.
    """
    import torch
    from ml_music_style_transfer_b200.pipeline import HostPipeline
    n = min(args.clips, args.e2e_clips) if n is None else min(args.clips, n)
    # page-locked host memory this pass needs per clip (inputs + outputs); never pin more than half of what the host has
    # free, shared by the ranks of this node (a box driven out of memory would take the whole run down)
    per_clip = 4 * CLIP_LEN + 4 * T_FRAMES * K + 4 * N_MELS * T_FRAMES + 4 * HOP * (T_FRAMES - 1) + 2 * 128 * int(ROLL_FS * CLIP_SECONDS)
    n_cap = int(0.5 * host_memory_available() / max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))) / per_clip)
    capped = n_cap < n
    if capped:
        n = max(256, (n_cap // 256) * 256)
    reps = max(1, min(args.steps, 3))
    extra = min(reps, n)
    h_audio = torch.empty((n + extra) * CLIP_LEN, dtype=torch.float32).pin_memory()
    h_audio[:n * CLIP_LEN].copy_(audio_d[:n * CLIP_LEN])
    h_audio[n * CLIP_LEN:].copy_(audio_d[:extra * CLIP_LEN])
    fs = T_FRAMES * K
    h_S = torch.empty((n + extra) * fs, dtype=torch.float32).pin_memory()
    h_S[:n * fs].copy_(S_d[:n * fs])
    h_S[n * fs:].copy_(S_d[:extra * fs])
    offs = np.asarray(notes_h[4], dtype=np.int64)

    def notes_window(k):   # pieces k .. k+n-1 (wrapping), re-based offsets
        if k == 0:
            return tuple(a[:offs[n]] for a in notes_h[:4]) + (offs[:n + 1],)
        lo, hi = int(offs[k]), int(offs[n])
        parts = [np.concatenate([a[lo:hi], a[:offs[k]]]) for a in notes_h[:4]]
        sizes = np.concatenate([np.diff(offs[k:n + 1]), np.diff(offs[:k + 1])])
        return tuple(parts) + (np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64),)
    windows = [notes_window(k % n) for k in range(reps + 1)]   # built outside the timing: they are the caller's inputs

    pipe = HostPipeline(n, CLIP_LEN, sr=SR, hop=HOP, n_mels=N_MELS, roll_fs=ROLL_FS, pitch_lo=PITCH_LO, n_keys=N_KEYS,
                        gl_iters=GL_ITERS, n_chunks=args.e2e_chunks, planes_to_host=planes_to_host, device=device, plan=plan)
    pipe.run(h_audio[:n * CLIP_LEN], h_S[:n * fs], windows[0])
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for r in range(reps):
        k = (r + 1) % (extra + 1)
        pipe.run(h_audio[k * CLIP_LEN:(k + n) * CLIP_LEN], h_S[k * fs:(k + n) * fs], windows[r + 1])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / reps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    h2d, d2h = pipe.bytes_per_run()
    out = {"value": world * n * CLIP_SECONDS / (ms * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "clips_per_gpu": n, "ms_per_step": ms, "pipeline_chunks": len(pipe.chunks),
           "pipeline_streams": pipe.n_streams, "api": "ml_music_style_transfer_b200.pipeline.HostPipeline.run",
           "capped_by_host_memory": capped,
           "inputs": "fresh per repetition (sliding window over the host buffers; note arrays re-staged every run)",
           "outputs_to_host": "log-mel, waveforms, frame-rate roll+onoff" + (", audio-rate planes" if planes_to_host else
                              " (audio-rate planes stay on the device for the model)")}
    del pipe, h_audio, h_S
    if not planes_to_host:
        bw = host_copy_ceiling(device, world, barrier)
        out["host_copy"] = {"gbs_per_rank_both_directions": bw, "ranks_copying": world,
                            "gbs_all_ranks": bw * world,
                            "e2e_gbs_per_rank": (h2d + d2h) / (ms * 1e-3) / 1e9,
                            "copy_only_ms_per_step": 1e3 * (h2d + d2h) / (bw * 1e9),
                            "note": "pinned copies alone, the pipeline's pattern (4 streams alternating H2D / D2H) on every rank "
                                    "at once, slowest rank; copy_only_ms is the time this box's host side needs just to move "
                                    "the step's bytes -- e2e cannot be faster than that at this N"}
    return out


# ------------------------------------------------------------------------------------------------------------------
# CPU arms (oracle port of the reference's path)
# ------------------------------------------------------------------------------------------------------------------
def _cpu_clip_job(seed):
    """Full hot path for ONE 4 s clip on one core; returns per-stage seconds."""
    from ml_music_style_transfer_b200 import synth
    from oracle import griffinlim as ogl, mel as omel, pianoroll as opr, stft as ostft
    y = synth.piano_clip(seed, CLIP_SECONDS, SR)[:CLIP_LEN]
    p, v, s, e = synth.midi_piece(seed, seconds=CLIP_SECONDS)
    t0 = time.perf_counter()
    omel.logmel(y, SR, N_FFT, HOP, N_MELS)
    t1 = time.perf_counter()
    roll, onoff = opr.binarize_and_onoff(opr.get_piano_roll(p, v, s, e, ROLL_FS))
    opr.upsample_to_audio_rate(roll, ROLL_FS, SR, CLIP_LEN, PITCH_LO, N_KEYS)
    opr.upsample_to_audio_rate(onoff, ROLL_FS, SR, CLIP_LEN, PITCH_LO, N_KEYS)
    t2 = time.perf_counter()
    S = np.abs(ostft.stft(y, N_FFT, HOP)).astype(np.float32)
    t3 = time.perf_counter()
    ogl.griffinlim(S, GL_ITERS, HOP, init_phase=ogl.random_phase(S.shape, seed))
    t4 = time.perf_counter()
    return t1 - t0, t2 - t1, t4 - t3


def cpu_baseline(cores, budget_s):
    """Oracle timed on a bounded sample of the same workload (whole 4 s clips through all three stages)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _cpu_clip_job(0)  # warm caches / imports
    t0 = time.perf_counter()
    n, acc = 0, np.zeros(3)
    while time.perf_counter() - t0 < budget_s or n < 2:
        acc += np.array(_cpu_clip_job(100 + n)); n += 1
    wall = time.perf_counter() - t0
    return {"value": n * CLIP_SECONDS / float(acc.sum()), "unit": "audio-s/s", "cores": cores, "kind": "port",
            "sample": f"{n} clips x 4 s (log-mel + piano-roll + GL-32 each), {wall:.1f} s of CPU work, single thread",
            "stage_seconds_per_clip": {"stft_logmel": acc[0] / n, "pianoroll_upsample": acc[1] / n, "griffinlim32": acc[2] / n}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = "1"
    per_step = max(cores, 8)  # clips per step: one per core
    with mp.get_context("fork").Pool(cores) as pool:
        for w in range(max(1, min(args.warmup, 1))):
            pool.map(_cpu_clip_job, range(cores))
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_clip_job, range(1000 + k * per_step, 1000 + (k + 1) * per_step))
        wall = time.perf_counter() - t0
    value = args.steps * per_step * CLIP_SECONDS / wall
    sample = f"{per_step} clips x 4 s per step (log-mel + piano-roll + GL-32 each), {cores} worker processes"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4 clip shape: 4 s @ 22.05 kHz, n_fft 2048, hop 512, 128 mels; 88-key roll @ 250 Hz -> "
                                   f"audio rate; Griffin-Lim {GL_ITERS} it (bounded sample)", "parallelism": f"{cores} host processes"},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4: 16384 x 4 s clips per GPU (weak scaling); c5: MusicNet-piano-scale corpus, 4080 x 30 s clips "
                         "in total, sharded over the ranks (strong scaling)")
    ap.add_argument("--clips", type=int, default=None, help="clips per GPU (default: 16384 for c4, 4080/world for c5)")
    ap.add_argument("--gl-sub", type=int, default=16384, help="clips per Griffin-Lim call (workspace bound)")
    ap.add_argument("--gather", action="store_true", help="(default when N > 1) time the optional NCCL all-gather of the log-mel features")
    ap.add_argument("--no-gather", action="store_true", help="skip the optional final NCCL gather at N > 1")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling pass (16384 clips in total) at N > 1")
    ap.add_argument("--no-single", action="store_true", help="skip the single 30 s clip latency section")
    ap.add_argument("--e2e-clips", type=int, default=None, help="clips per GPU for the host-buffer end-to-end pass (default: the same as --clips)")
    ap.add_argument("--e2e-chunks", type=int, default=32, help="pipeline depth of the end-to-end pass (chunks rotating over MST_E2E_STREAMS streams, default 4)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--serial", action="store_true", help="time the three stages back to back on one stream instead of the "
                                                        "chunk-pipelined DevicePipeline schedule")
    ap.add_argument("--step-chunks", type=int, default=32, help="chunks of the pipelined step (rotating over --step-streams streams)")
    ap.add_argument("--step-streams", type=int, default=4, help="CUDA streams the chunks of the pipelined step rotate over")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traffic", type=float, default=None, help="ncu dram bytes per GL-iteration launch (from profiles/)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "c5":
        global CLIP_SECONDS, CLIP_LEN, T_FRAMES
        CLIP_SECONDS, CLIP_LEN = 30.0, 30 * SR
        T_FRAMES = 1 + CLIP_LEN // HOP
        if args.clips is None:
            rank = int(os.environ.get("RANK", "0"))
            base, extra = divmod(4080, world)
            args.clips = base + (1 if rank < extra else 0)
        args.gl_sub = min(args.gl_sub, 1024)
        args.e2e_clips = min(args.e2e_clips or 256, 256)
    elif args.clips is None:
        args.clips = 16384
    if args.e2e_clips is None:
        args.e2e_clips = args.clips
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
