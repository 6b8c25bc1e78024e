// STFT / mel / Griffin-Lim for every power-of-two n_fft OTHER than 2048 (64 ... 16384).
//
// The reference calls librosa with n_fft = 2048 only (preprocessing/preprocess.py:25,48, model/inference.py:105-110) and
// the warp-per-frame kernels of stft.cu / griffinlim.cu are built around that size.  librosa's signatures take any
// n_fft, so the same C entry points (mst_stft_f32, mst_stft_mel_f32, mst_spectral_convergence_f32, mst_griffinlim_f32)
// route batches created with another n_fft here.  This is the general path, not the tuned one: one CTA per frame, the
// real FFT as an n_fft/2-point complex Stockham radix-2 FFT in shared memory (even/odd packing + split butterflies),
// twiddles from a per-(device, n_fft) table computed in double.  Griffin-Lim keeps librosa's order of operations
// literally (istft: irfft -> window -> overlap-add in frame order in float32 -> window-sum-square normalisation; stft
// of the trimmed, reflect-padded signal; momentum update), one kernel per step, all state in the caller's workspace.
#include <math.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>
#include "mel_plan.cuh"
#include "mst_common.cuh"

namespace mst {

constexpr int kGenThreads = 256;
constexpr int kGenMaxSmem = 2 * (16384 / 2) * (int)sizeof(float2);  // two ping-pong buffers of the largest FFT (128 KB)

struct GenGeom {
  int n_fft, M, logM, K, hop, pad_mode;  // M = n_fft / 2 (complex FFT length), K = M + 1 bins
};

// ---- twiddle tables: tw[i] = exp(-2*pi*i*i / n_fft), i < n_fft / 2 ---------------------------------------------------------
static std::mutex g_gen_mutex;
struct GenTable { int dev, n_fft; float2* d_tw; };
static std::vector<GenTable> g_gen_tables;

static int gen_twiddles(int n_fft, const float2** out) {
  int dev = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_gen_mutex);
  for (const GenTable& t : g_gen_tables)
    if (t.dev == dev && t.n_fft == n_fft) { *out = t.d_tw; return MST_OK; }
  std::vector<float2> tw((size_t)n_fft / 2);
  const double two_pi = 6.283185307179586476925286766559;
  for (int i = 0; i < n_fft / 2; ++i) {
    const double a = -two_pi * (double)i / (double)n_fft;
    tw[(size_t)i] = make_float2((float)cos(a), (float)sin(a));
  }
  float2* d = nullptr;
  MST_CUDA_OK(cudaMalloc(&d, sizeof(float2) * tw.size()));
  MST_CUDA_OK(cudaMemcpy(d, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
  g_gen_tables.push_back({dev, n_fft, d});
  *out = d;
  return MST_OK;
}

bool generic_n_fft_ok(int n_fft) { return n_fft >= 64 && n_fft <= 16384 && (n_fft & (n_fft - 1)) == 0; }

static GenGeom geom_of(const mst_batch* b) {
  GenGeom g;
  g.n_fft = b->n_fft; g.M = b->n_fft / 2; g.K = g.M + 1; g.hop = b->hop; g.pad_mode = b->pad_mode;
  g.logM = 0;
  while ((1 << g.logM) < g.M) ++g.logM;
  return g;
}

// ---- device helpers -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// M-point complex FFT (SIGN = -1 forward, +1 inverse, unnormalised) of `in`, ping-ponging with `out`; the caller has
// synchronised after filling `in`.  Returns the buffer that holds the result (natural order); ends with a CTA barrier.
template <int SIGN>
__device__ float2* block_fft(float2* in, float2* out, int M, int logM, const float2* __restrict__ tw) {
  const int half = M >> 1;
  for (int s = 0; s < logM; ++s) {
    const int Ns = 1 << s;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
      const int k = j & (Ns - 1);
      float2 w = __ldg(tw + ((size_t)k << (logM - s)));  // exp(-2*pi*i*k / (2*Ns))
      if (SIGN > 0) w.y = -w.y;
      const float2 a0 = in[j];
      const float2 a1 = cmulf(in[j + half], w);
      const int j0 = ((j - k) << 1) + k;
      out[j0] = make_float2(a0.x + a1.x, a0.y + a1.y);
      out[j0 + Ns] = make_float2(a0.x - a1.x, a0.y - a1.y);
    }
    __syncthreads();
    float2* t = in; in = out; out = t;
  }
  return in;
}

// bin k (0..M) of the n_fft-point real FFT from the M-point FFT Z of z[m] = x[2m] + i*x[2m+1]
__device__ __forceinline__ float2 rfft_bin(const float2* Z, int k, int M, const float2* __restrict__ tw) {
  const float2 zk = Z[k == M ? 0 : k];
  const float2 zm = Z[k == 0 ? 0 : M - k];
  const float2 W = k < M ? __ldg(tw + k) : make_float2(-1.0f, 0.0f);
  const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
  const float2 D = make_float2(zk.x - zm.x, zk.y + zm.y);
  const float2 WD = cmulf(W, D);
  return make_float2(fmaf(0.5f, WD.y, E.x), fmaf(-0.5f, WD.x, E.y));
}

// packed spectrum Z[k] = E[k] + i*O[k] (k < M) of the real signal whose rfft is X[0..M] (imaginary parts of DC and Nyquist
// are ignored, as numpy.fft.irfft does)
__device__ __forceinline__ float2 irfft_pack(float2 xk, float2 xm, int k, const float2* __restrict__ tw) {
  // E = (X[k] + conj(X[M-k])) / 2, O = (X[k] - conj(X[M-k])) / 2 * conj(W^k)
  const float2 E = make_float2(0.5f * (xk.x + xm.x), 0.5f * (xk.y - xm.y));
  const float2 D = make_float2(0.5f * (xk.x - xm.x), 0.5f * (xk.y + xm.y));
  float2 W = __ldg(tw + k);
  W.y = -W.y;
  const float2 O = cmulf(D, W);
  return make_float2(E.x - O.y, E.y + O.x);
}

__device__ __forceinline__ bool frame_of_block(const ClipDesc* __restrict__ clips, const int32_t* __restrict__ tile_clip,
                                               int* c_out, ClipDesc* cd_out, int* t_out) {
  const int tile = blockIdx.x / kWarpsPerCta, sub = blockIdx.x % kWarpsPerCta;
  const int c = __ldg(tile_clip + tile);
  const ClipDesc cd = clips[c];
  const int t = (tile - cd.tile_offset) * kWarpsPerCta + sub;
  *c_out = c; *cd_out = cd; *t_out = t;
  return t < cd.frames;
}

// windowed, centre-padded frame t of signal x (length L) packed into A[m] = (x[2m], x[2m+1]); `scale` (may be NULL) is a
// per-sample factor indexed like x
__device__ __forceinline__ void load_frame(float2* A, const float* __restrict__ x, const float* __restrict__ scale, int64_t L,
                                           int t, const GenGeom& g, const float* __restrict__ window) {
  const int64_t base = (int64_t)t * g.hop - g.M;
  for (int j = threadIdx.x; j < g.n_fft; j += blockDim.x) {
    int64_t n = base + j;
    bool ok = true;
    if (n < 0) {
      if (g.pad_mode == MST_PAD_REFLECT) n = -n; else ok = false;
    } else if (n >= L) {
      if (g.pad_mode == MST_PAD_REFLECT) n = 2 * (L - 1) - n; else ok = false;
    }
    float v = 0.0f;
    if (ok) {
      v = x[n] * __ldg(window + j);
      if (scale) v *= __ldg(scale + n);
    }
    reinterpret_cast<float*>(A)[j] = v;
  }
}

constexpr int kGenConv = 100, kGenMel = 101;

struct GenMel {
  const float* W;        // [n_mels][K] dense filterbank
  const int32_t* k_lo;   // [n_mels] first non-zero bin
  const int32_t* k_hi;   // [n_mels] one past the last non-zero bin
  int n_mels;
  int apply_log1p;
};

struct GenConv {
  const float* target;
  double* num;
  double* den;
};

// ---- STFT and its epilogues ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kGenThreads)
gen_stft_kernel(const float* __restrict__ audio, const ClipDesc* __restrict__ clips, const int32_t* __restrict__ tile_clip,
                GenGeom g, const float2* __restrict__ tw, const float* __restrict__ window, int layout,
                void* __restrict__ out_v, GenMel mel, GenConv conv) {
  extern __shared__ __align__(16) float2 gen_smem[];
  float2* A = gen_smem;
  float2* B = gen_smem + g.M;
  int c, t;
  ClipDesc cd;
  if (!frame_of_block(clips, tile_clip, &c, &cd, &t)) return;
  load_frame(A, audio + cd.sample_offset, nullptr, cd.length, t, g, window);
  __syncthreads();
  const float2* Z = block_fft<-1>(A, B, g.M, g.logM, tw);
  const int64_t frame = cd.frame_offset + t;

  if (MODE == MST_OUT_COMPLEX) {
    float2* out = reinterpret_cast<float2*>(out_v) + frame * g.K;
    for (int k = threadIdx.x; k <= g.M; k += blockDim.x) out[k] = rfft_bin(Z, k, g.M, tw);
    return;
  }
  if (MODE == kGenConv) {
    double num = 0.0, den = 0.0;
    for (int k = threadIdx.x; k <= g.M; k += blockDim.x) {
      const float2 X = rfft_bin(Z, k, g.M, tw);
      const float mag = sqrtf(fmaf(X.x, X.x, X.y * X.y));
      const int64_t idx = layout == MST_LAYOUT_FRAME_MAJOR ? frame * g.K + k
                                                            : cd.frame_offset * g.K + (int64_t)k * cd.frames + t;
      const float sref = __ldg(conv.target + idx);
      const float d = mag - sref;
      num += (double)d * (double)d;
      den += (double)sref * (double)sref;
    }
    __shared__ double red[2][kGenThreads / 32];
    for (int off = 16; off; off >>= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, off);
      den += __shfl_xor_sync(0xffffffffu, den, off);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = num; red[1][threadIdx.x >> 5] = den; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kGenThreads / 32; ++w) { num += red[0][w]; den += red[1][w]; }
      atomicAdd(conv.num + c, num);
      atomicAdd(conv.den + c, den);
    }
    return;
  }
  if (MODE == kGenMel) {
    // power spectrum into the idle half of the ping-pong buffer's neighbour (float view of the buffer Z does not live in)
    float* P = reinterpret_cast<float*>(Z == A ? B : A);  // K floats <= M float2 for every M >= 2
    for (int k = threadIdx.x; k <= g.M; k += blockDim.x) {
      const float2 X = rfft_bin(Z, k, g.M, tw);
      P[k] = fmaf(X.x, X.x, X.y * X.y);
    }
    __syncthreads();
    float* out = reinterpret_cast<float*>(out_v);
    for (int m = threadIdx.x; m < mel.n_mels; m += blockDim.x) {
      const float* w = mel.W + (size_t)m * g.K;
      float acc = 0.0f;
      for (int k = __ldg(mel.k_lo + m); k < __ldg(mel.k_hi + m); ++k) acc = fmaf(__ldg(w + k), P[k], acc);
      if (mel.apply_log1p) acc = log1pf(acc);
      const int64_t idx = layout == MST_LAYOUT_FRAME_MAJOR ? frame * mel.n_mels + m
                                                            : cd.frame_offset * mel.n_mels + (int64_t)m * cd.frames + t;
      out[idx] = acc;
    }
    return;
  }
  float* out = reinterpret_cast<float*>(out_v);
  for (int k = threadIdx.x; k <= g.M; k += blockDim.x) {
    const float2 X = rfft_bin(Z, k, g.M, tw);
    const float p = fmaf(X.x, X.x, X.y * X.y);
    const float v = MODE == MST_OUT_MAGNITUDE ? sqrtf(p) : (MODE == MST_OUT_POWER ? p : log1pf(p));
    const int64_t idx = layout == MST_LAYOUT_FRAME_MAJOR ? frame * g.K + k
                                                          : cd.frame_offset * g.K + (int64_t)k * cd.frames + t;
    out[idx] = v;
  }
}

template <int MODE>
static int launch_gen_stft(const float* d_audio, const mst_batch* b, int layout, void* d_out, const GenMel& mel,
                           const GenConv& conv, cudaStream_t s) {
  const GenGeom g = geom_of(b);
  const float2* tw = nullptr;
  int rc = gen_twiddles(g.n_fft, &tw);
  if (rc) return rc;
  const size_t smem = sizeof(float2) * 2 * (size_t)g.M;
  if (smem > 48 * 1024) {  // n_fft >= 8192: opt in once per device and instantiation (to the largest size served)
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    MST_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(MST_ERR_INVALID, "device index %d out of range", dev);
    if (!attr_set[dev].load(std::memory_order_acquire)) {
      MST_CUDA_OK(cudaFuncSetAttribute(gen_stft_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGenMaxSmem));
      attr_set[dev].store(true, std::memory_order_release);
    }
  }
  const int64_t blocks = (int64_t)b->total_tiles * kWarpsPerCta;
  if (blocks > 0x7fffffff) return fail(MST_ERR_INVALID, "too many frames for one launch");
  gen_stft_kernel<MODE><<<(unsigned)blocks, kGenThreads, smem, s>>>(d_audio, b->d_clips, b->d_tile_clip, g, tw, b->d_window,
                                                                    layout, d_out, mel, conv);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

int generic_stft(const float* d_audio, const mst_batch* b, int out_mode, int layout, void* d_out, cudaStream_t s) {
  const GenMel nomel{};
  const GenConv noconv{};
  switch (out_mode) {
    case MST_OUT_COMPLEX:
      if (layout != MST_LAYOUT_FRAME_MAJOR)
        return fail(MST_ERR_UNSUPPORTED, "complex STFT output is frame-major only (librosa's native Fortran order)");
      return launch_gen_stft<MST_OUT_COMPLEX>(d_audio, b, layout, d_out, nomel, noconv, s);
    case MST_OUT_MAGNITUDE: return launch_gen_stft<MST_OUT_MAGNITUDE>(d_audio, b, layout, d_out, nomel, noconv, s);
    case MST_OUT_POWER: return launch_gen_stft<MST_OUT_POWER>(d_audio, b, layout, d_out, nomel, noconv, s);
    case MST_OUT_LOG1P_POWER: return launch_gen_stft<MST_OUT_LOG1P_POWER>(d_audio, b, layout, d_out, nomel, noconv, s);
    default: return fail(MST_ERR_INVALID, "bad out_mode %d", out_mode);
  }
}

int generic_spectral_convergence(const float* d_y, const mst_batch* b, const float* d_S, int s_layout, double* d_num,
                                 double* d_den, cudaStream_t s) {
  GenConv conv{d_S, d_num, d_den};
  return launch_gen_stft<kGenConv>(d_y, b, s_layout, nullptr, GenMel{}, conv, s);
}

int generic_stft_mel(const float* d_audio, const mst_batch* b, const mst_mel_plan* plan, int apply_log1p, int layout,
                     float* d_out, cudaStream_t s) {
  if (!plan->d_dense_w || plan->n_bins != b->n_fft / 2 + 1)
    return fail(MST_ERR_INVALID, "mel plan has %d bins, the batch (n_fft=%d) needs %d", plan->n_bins, b->n_fft,
                b->n_fft / 2 + 1);
  GenMel mel{plan->d_dense_w, plan->d_k_lo, plan->d_k_hi, plan->n_mels, apply_log1p};
  return launch_gen_stft<kGenMel>(d_audio, b, layout, d_out, mel, GenConv{}, s);
}

// ---- Griffin-Lim --------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gen_uniform_hash(unsigned long long seed, unsigned long long idx) {
  // same counter-based generator as griffinlim.cu (the random initial phase is unspecified in librosa)
  unsigned x = (unsigned)idx * 0x9E3779B1u ^ ((unsigned)(idx >> 32) + 0x7F4A7C15u) * 0x85EBCA77u ^ (unsigned)seed;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  x += (unsigned)(seed >> 32) * 0x27D4EB2Fu + 0x165667B1u;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

// S (any layout, magnitudes or log1p-power) -> frame-major magnitudes; proj = S * exp(2*pi*i*u) (or S for init_mode 1)
__global__ void __launch_bounds__(kGenThreads)
gen_gl_init_kernel(const float* __restrict__ S_in, int layout, int is_log1p_power, const ClipDesc* __restrict__ clips,
                   const int32_t* __restrict__ tile_clip, int K, const float* __restrict__ init_phase, int init_mode,
                   unsigned long long seed, float* __restrict__ S_out, float2* __restrict__ proj) {
  int c, t;
  ClipDesc cd;
  if (!frame_of_block(clips, tile_clip, &c, &cd, &t)) return;
  const int64_t frame = cd.frame_offset + t;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int64_t idx = layout == MST_LAYOUT_FRAME_MAJOR ? frame * K + k : cd.frame_offset * K + (int64_t)k * cd.frames + t;
    float s = S_in[idx];
    if (is_log1p_power) s = sqrtf(expm1f(fminf(fmaxf(s, 0.0f), 20.0f)));  // inference.py:109
    S_out[frame * K + k] = s;
    float sn = 0.0f, cs = 1.0f;
    if (init_mode == 0) {
      const float u = init_phase ? __ldg(init_phase + idx) : gen_uniform_hash(seed, (unsigned long long)(frame * K + k));
      const float turns = u - rintf(u);
      __sincosf(6.283185307179586f * turns, &sn, &cs);
    }
    proj[frame * K + k] = make_float2(s * cs, s * sn);
  }
}

// one frame of istft: window * irfft(proj[frame]) -> fbuf[frame][n_fft]
__global__ void __launch_bounds__(kGenThreads)
gen_gl_synth_kernel(const float2* __restrict__ proj, const ClipDesc* __restrict__ clips, const int32_t* __restrict__ tile_clip,
                    GenGeom g, const float2* __restrict__ tw, const float* __restrict__ window, float* __restrict__ fbuf) {
  extern __shared__ __align__(16) float2 gen_smem[];
  float2* A = gen_smem;
  float2* B = gen_smem + g.M;
  int c, t;
  ClipDesc cd;
  if (!frame_of_block(clips, tile_clip, &c, &cd, &t)) return;
  const int64_t frame = cd.frame_offset + t;
  const float2* X = proj + frame * g.K;
  for (int k = threadIdx.x; k < g.M; k += blockDim.x) {
    float2 xk = X[k], xm = X[g.M - k];
    if (k == 0) { xk.y = 0.0f; xm.y = 0.0f; }
    A[k] = irfft_pack(xk, xm, k, tw);
  }
  __syncthreads();
  const float2* z = block_fft<+1>(A, B, g.M, g.logM, tw);
  const float inv = 1.0f / (float)g.M;
  float* out = fbuf + frame * g.n_fft;
  for (int m = threadIdx.x; m < g.M; m += blockDim.x) {
    const float2 v = z[m];
    out[2 * m] = v.x * inv * __ldg(window + 2 * m);
    out[2 * m + 1] = v.y * inv * __ldg(window + 2 * m + 1);
  }
}

// librosa.istft's overlap-add (float32, frames in increasing order) and window-sum-square normalisation.  FINAL: write
// the centre-trimmed waveform y[0, hop*(T-1)); else the normalised signal over the whole accumulator span.
template <bool FINAL>
__global__ void gen_gl_ola_kernel(const float* __restrict__ fbuf, const ClipDesc* __restrict__ clips, int n_fft, int hop,
                                  const float* __restrict__ inv_wss, const int64_t* __restrict__ wss_off,
                                  float* __restrict__ dst) {
  const int c = blockIdx.y;
  const ClipDesc cd = clips[c];
  const int64_t span = n_fft + (int64_t)hop * (cd.frames - 1);
  const int64_t p0 = FINAL ? n_fft / 2 : 0, p1 = FINAL ? n_fft / 2 + cd.length : span;
  const float* env = inv_wss + wss_off[c];
  const float* frames = fbuf + cd.frame_offset * n_fft;
  for (int64_t p = p0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t f_lo = p - (n_fft - 1);
    f_lo = f_lo > 0 ? (f_lo + hop - 1) / hop : 0;
    const int64_t f_hi = min((int64_t)cd.frames - 1, p / hop);
    float sum = 0.0f;
    for (int64_t f = f_lo; f <= f_hi; ++f) sum += frames[f * n_fft + (p - f * hop)];
    const float v = sum * __ldg(env + p);
    if (FINAL) dst[cd.sample_offset + (p - p0)] = v;
    else dst[cd.acc_offset + p] = v;
  }
}

// rebuilt = stft(y_{j-1}) for one frame; angles = rebuilt - alpha * tprev; proj = S * angles / (|angles| + 1e-16)
__global__ void __launch_bounds__(kGenThreads)
gen_gl_analysis_kernel(const float* __restrict__ ynorm, const float* __restrict__ S, float2* __restrict__ tprev,
                       float2* __restrict__ proj, const ClipDesc* __restrict__ clips, const int32_t* __restrict__ tile_clip,
                       GenGeom g, const float2* __restrict__ tw, const float* __restrict__ window, float alpha, int last_iter) {
  extern __shared__ __align__(16) float2 gen_smem[];
  float2* A = gen_smem;
  float2* B = gen_smem + g.M;
  int c, t;
  ClipDesc cd;
  if (!frame_of_block(clips, tile_clip, &c, &cd, &t)) return;
  // sample n of the trimmed signal lives at accumulator position n + n_fft/2
  load_frame(A, ynorm + cd.acc_offset + g.M, nullptr, cd.length, t, g, window);
  __syncthreads();
  const float2* Z = block_fft<-1>(A, B, g.M, g.logM, tw);
  const int64_t frame = cd.frame_offset + t;
  for (int k = threadIdx.x; k <= g.M; k += blockDim.x) {
    const float2 r = rfft_bin(Z, k, g.M, tw);
    const int64_t i = frame * g.K + k;
    const float2 tp = tprev[i];
    if (!last_iter) tprev[i] = r;
    const float ax = fmaf(-alpha, tp.x, r.x), ay = fmaf(-alpha, tp.y, r.y);
    const float sc = __ldg(S + i) * rsqrtf(fmaf(ax, ax, fmaf(ay, ay, 1e-32f)));
    proj[i] = make_float2(ax * sc, ay * sc);
  }
}

static size_t gen_align(size_t x) { return (x + 255) / 256 * 256; }

size_t generic_gl_workspace_bytes(const mst_batch* b) {
  const size_t spec = (size_t)b->total_frames * (size_t)(b->n_fft / 2 + 1);
  return 2 * gen_align(spec * sizeof(float2)) + gen_align(spec * sizeof(float)) +
         gen_align((size_t)b->total_frames * (size_t)b->n_fft * sizeof(float)) + gen_align((size_t)b->total_acc * sizeof(float)) + 256;
}

int generic_griffinlim(const float* d_S, int s_layout, int s_is_log1p_power, const mst_batch* b, int n_iter, float momentum,
                       const float* d_init_phase, int init_mode, uint64_t seed, float* d_y_out, void* d_workspace,
                       cudaStream_t s) {
  const GenGeom g = geom_of(b);
  const float2* tw = nullptr;
  int rc = gen_twiddles(g.n_fft, &tw);
  if (rc) return rc;
  const size_t spec = (size_t)b->total_frames * (size_t)g.K;
  char* ws = reinterpret_cast<char*>(d_workspace);
  float2* proj = reinterpret_cast<float2*>(ws); ws += gen_align(spec * sizeof(float2));
  float2* tprev = reinterpret_cast<float2*>(ws); ws += gen_align(spec * sizeof(float2));
  float* S_t = reinterpret_cast<float*>(ws); ws += gen_align(spec * sizeof(float));
  float* fbuf = reinterpret_cast<float*>(ws); ws += gen_align((size_t)b->total_frames * (size_t)g.n_fft * sizeof(float));
  float* ynorm = reinterpret_cast<float*>(ws);

  const size_t smem = sizeof(float2) * 2 * (size_t)g.M;
  if (smem > 48 * 1024) {
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    MST_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(MST_ERR_INVALID, "device index %d out of range", dev);
    if (!attr_set[dev].load(std::memory_order_acquire)) {
      MST_CUDA_OK(cudaFuncSetAttribute(gen_gl_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGenMaxSmem));
      MST_CUDA_OK(cudaFuncSetAttribute(gen_gl_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGenMaxSmem));
      attr_set[dev].store(true, std::memory_order_release);
    }
  }
  const int64_t blocks64 = (int64_t)b->total_tiles * kWarpsPerCta;
  if (blocks64 > 0x7fffffff) return fail(MST_ERR_INVALID, "too many frames for one launch");
  const unsigned blocks = (unsigned)blocks64;
  int64_t max_span = 0;
  for (int c = 0; c < b->n_clips; ++c)
    max_span = std::max<int64_t>(max_span, g.n_fft + (int64_t)g.hop * (b->h_clips[c].frames - 1));

  MST_CUDA_OK(cudaMemsetAsync(tprev, 0, spec * sizeof(float2), s));
  gen_gl_init_kernel<<<blocks, kGenThreads, 0, s>>>(d_S, s_layout, s_is_log1p_power, b->d_clips, b->d_tile_clip, g.K,
                                                    d_init_phase, init_mode, (unsigned long long)seed, S_t, proj);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  const float alpha = momentum / (1.0f + momentum);
  for (int j = 0; j <= n_iter; ++j) {
    if (j > 0) {
      gen_gl_analysis_kernel<<<blocks, kGenThreads, smem, s>>>(ynorm, S_t, tprev, proj, b->d_clips, b->d_tile_clip, g, tw,
                                                               b->d_window, alpha, j == n_iter);
      MST_CUDA_OK(cudaGetLastError());
      count_launch();
    }
    gen_gl_synth_kernel<<<blocks, kGenThreads, smem, s>>>(proj, b->d_clips, b->d_tile_clip, g, tw, b->d_window, fbuf);
    MST_CUDA_OK(cudaGetLastError());
    count_launch();
    for (int c0 = 0; c0 < b->n_clips; c0 += 65535) {
      const int nc = std::min(65535, b->n_clips - c0);
      dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(256, (max_span + 255) / 256)), (unsigned)nc);
      if (j == n_iter)
        gen_gl_ola_kernel<true><<<grid, 256, 0, s>>>(fbuf, b->d_clips + c0, g.n_fft, g.hop, b->d_inv_wss, b->d_wss_offset + c0, d_y_out);
      else
        gen_gl_ola_kernel<false><<<grid, 256, 0, s>>>(fbuf, b->d_clips + c0, g.n_fft, g.hop, b->d_inv_wss, b->d_wss_offset + c0, ynorm);
      MST_CUDA_OK(cudaGetLastError());
      count_launch();
    }
  }
  return MST_OK;
}

}  // namespace mst
