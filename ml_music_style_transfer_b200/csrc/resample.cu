// "Next" row 1: the resampling half of librosa.load(path, sr=hp.sr)
// (reference preprocessing/preprocess.py:106, model/inference.py:54, tests/test_griffinlim.py:16).
// Band-limited sinc interpolation exactly as resampy 0.2.2 'kaiser_best' evaluates it (librosa 0.8's default
// res_type): half-window rolloff*sinc(rolloff*t)*kaiser(beta) tabulated 512 times per zero crossing over 64 zero
// crossings, linear interpolation between table entries, left wing then right wing.  Three kernels: exact halving of the
// rate -> resample_half_kernel (one symmetric FIR); rational ratios p/q with q <= 2048 -> resample_phase_kernel (tap weights
// precomputed per phase); anything else -> resample_kernel (one thread per output interpolating the 256 KB L2-resident
// (value, delta) table).  Index arithmetic is done in double so that table offsets match
// the Python evaluation bit for bit.
#include <math.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <vector>
#include "mst_common.cuh"

namespace mst {

constexpr int kRsZeros = 64, kRsPrecision = 9, kRsTable = 1 << kRsPrecision;  // 512 entries per zero crossing
constexpr int kRsWin = kRsZeros * kRsTable + 1;                               // 32769
constexpr double kRsBeta = 14.769656459379492, kRsRolloff = 0.9475937167399596;

__global__ void resample_kernel(const float* __restrict__ x, int64_t n_in, float* __restrict__ y, int64_t n_out,
                                int64_t n_fix, const float2* __restrict__ table, double time_increment, double scale,
                                int index_step) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_fix; t += (int64_t)gridDim.x * blockDim.x) {
    if (t >= n_out) {  // librosa util.fix_length: zero padding up to ceil(n * ratio)
      y[t] = 0.0f;
      continue;
    }
    const double time_register = __dmul_rn((double)t, time_increment);
    const int64_t n = (int64_t)time_register;
    float acc = 0.0f;
    // left wing
    double frac = __dmul_rn(scale, time_register - (double)n);
    double index_frac = __dmul_rn(frac, (double)kRsTable);
    int offset = (int)index_frac;
    float eta = (float)(index_frac - (double)offset);
    int64_t lim = (kRsWin - offset) / index_step;
    const int64_t i_max = n + 1 < lim ? n + 1 : lim;
    for (int64_t i = 0; i < i_max; ++i) {
      const float2 w = __ldg(table + offset + i * index_step);
      acc = fmaf(fmaf(eta, w.y, w.x), __ldg(x + n - i), acc);
    }
    // right wing
    frac = scale - frac;
    index_frac = __dmul_rn(frac, (double)kRsTable);
    offset = (int)index_frac;
    eta = (float)(index_frac - (double)offset);
    lim = (kRsWin - offset) / index_step;
    const int64_t k_max = n_in - n - 1 < lim ? n_in - n - 1 : lim;
    for (int64_t k = 0; k < k_max; ++k) {
      const float2 w = __ldg(table + offset + k * index_step);
      acc = fmaf(fmaf(eta, w.y, w.x), __ldg(x + n + k + 1), acc);
    }
    y[t] = acc;
  }
}

// Rational ratios sr_in / sr_out = p / q with a small q (48 -> 44.1 kHz: 160 / 147; 44.1 -> 16 kHz: 441 / 160; any integer
// up-sampling factor ...): output t sits at input position t*p/q, so only q distinct table phases exist.  The interpolated
// tap weights of every phase (left wing then right wing, resampy's order) are evaluated once on the host with the same
// float32 fma as resample_kernel and stored as one 16-byte-aligned row per phase; a thread then runs
// acc = fma(w[j], x[..], acc) over its row with 128-bit weight loads -- no per-tap table lookup, no interpolation and no
// index arithmetic in the loop.  Summation order is the reference's.
struct PhaseGeom {
  int64_t p, q;
  int row;        // floats per weight row (multiple of 4): [0, max_left) left wing, [max_left, max_left + max_right) right wing
  int max_left;   // multiple of 4
};

__global__ void __launch_bounds__(256)
resample_phase_kernel(const float* __restrict__ x, int64_t n_in, float* __restrict__ y, int64_t n_out, int64_t n_fix,
                      const float* __restrict__ W, const int2* __restrict__ lims, PhaseGeom g) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_fix; t += (int64_t)gridDim.x * blockDim.x) {
    if (t >= n_out) {  // librosa util.fix_length: zero padding up to ceil(n * ratio)
      y[t] = 0.0f;
      continue;
    }
    const int64_t tp = t * g.p;
    const int64_t n = tp / g.q;
    const int phase = (int)(tp - n * g.q);
    const int2 lim = __ldg(lims + phase);
    const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)phase * g.row);
    float acc = 0.0f;
    // left wing: x[n - i], i = 0 .. min(n + 1, lim.x) - 1
    const int i_max = (int)(n + 1 < lim.x ? n + 1 : lim.x);
    const float* xl = x + n;
    int i = 0;
    for (; i + 4 <= i_max; i += 4) {
      const float4 w = __ldg(w4 + (i >> 2));
      acc = fmaf(w.x, __ldg(xl - i), acc);
      acc = fmaf(w.y, __ldg(xl - i - 1), acc);
      acc = fmaf(w.z, __ldg(xl - i - 2), acc);
      acc = fmaf(w.w, __ldg(xl - i - 3), acc);
    }
    for (; i < i_max; ++i) acc = fmaf(__ldg(W + (size_t)phase * g.row + i), __ldg(xl - i), acc);
    // right wing: x[n + 1 + k], k = 0 .. min(n_in - n - 1, lim.y) - 1
    const int64_t avail = n_in - n - 1;
    const int k_max = (int)(avail < lim.y ? (avail > 0 ? avail : 0) : lim.y);
    const float* xr = x + n + 1;
    const float4* r4 = w4 + (g.max_left >> 2);
    int k = 0;
    for (; k + 4 <= k_max; k += 4) {
      const float4 w = __ldg(r4 + (k >> 2));
      acc = fmaf(w.x, __ldg(xr + k), acc);
      acc = fmaf(w.y, __ldg(xr + k + 1), acc);
      acc = fmaf(w.z, __ldg(xr + k + 2), acc);
      acc = fmaf(w.w, __ldg(xr + k + 3), acc);
    }
    for (; k < k_max; ++k) acc = fmaf(__ldg(W + (size_t)phase * g.row + g.max_left + k), __ldg(xr + k), acc);
    y[t] = acc;
  }
}

// Decimation by exactly 2 (44.1 -> 22.05 kHz, 48 -> 24 kHz ...): time_increment is the integer 2, so every output has
// the SAME table phase (offset 0 / eta 0 on the left wing, offset 256 / eta 0 on the right) and resampy's two wings
// collapse into one symmetric 255-tap FIR  y[t] = sum_{d=-127}^{127} c_|d| x[2t + d],  c_k = table[256 k]  (the wing
// limits (32769 - offset) / 256 give 128 left and 127 right taps; beyond the signal the wings stop, i.e. zero padding).
// A CTA of 128 threads stages its input span in shared memory split into even and odd samples (so that consecutive
// lanes read consecutive words), keeps the 128 coefficients there too (broadcast reads), and every thread accumulates
// 4 outputs at once: 1 coefficient load + 4 sample loads + 4 FMAs per tap instead of 2 global loads + index arithmetic
// per tap and output.
constexpr int kHalfThreads = 128, kHalfPerThread = 4, kHalfOut = kHalfThreads * kHalfPerThread;  // 512 outputs per CTA
constexpr int kHalfSpan = kHalfOut + 128;                                                       // even / odd samples staged

__global__ void __launch_bounds__(kHalfThreads)
resample_half_kernel(const float* __restrict__ x, int64_t n_in, float* __restrict__ y, int64_t n_out, int64_t n_fix,
                     const float2* __restrict__ table) {
  __shared__ float s_c[128];
  __shared__ float s_e[kHalfSpan], s_o[kHalfSpan];
  const int tid = threadIdx.x;
  s_c[tid] = __ldg(table + tid * 256).x;
  for (int64_t T0 = (int64_t)blockIdx.x * kHalfOut; T0 < n_fix; T0 += (int64_t)gridDim.x * kHalfOut) {
    __syncthreads();
    for (int i = tid; i < kHalfSpan; i += kHalfThreads) {   // s_e[i] = x[2 (T0 - 64 + i)], s_o[i] = x[2 (T0 - 64 + i) + 1]
      const int64_t m = 2 * (T0 - 64 + i);
      s_e[i] = (m >= 0 && m < n_in) ? __ldg(x + m) : 0.0f;
      s_o[i] = (m + 1 >= 0 && m + 1 < n_in) ? __ldg(x + m + 1) : 0.0f;
    }
    __syncthreads();
    float acc[kHalfPerThread];
#pragma unroll
    for (int q = 0; q < kHalfPerThread; ++q) acc[q] = 0.0f;
    // even taps d = 2e, e = -63..63: x[2t + 2e] = s_e[(t - T0) + 64 + e]
#pragma unroll 4
    for (int e = -63; e <= 63; ++e) {
      const float c = s_c[e < 0 ? -2 * e : 2 * e];
#pragma unroll
      for (int q = 0; q < kHalfPerThread; ++q) acc[q] = fmaf(c, s_e[tid + kHalfThreads * q + 64 + e], acc[q]);
    }
    // odd taps d = 2e + 1, e = -64..63: x[2t + 2e + 1] = s_o[(t - T0) + 64 + e]
#pragma unroll 4
    for (int e = -64; e <= 63; ++e) {
      const int d = 2 * e + 1;
      const float c = s_c[d < 0 ? -d : d];
#pragma unroll
      for (int q = 0; q < kHalfPerThread; ++q) acc[q] = fmaf(c, s_o[tid + kHalfThreads * q + 64 + e], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kHalfPerThread; ++q) {
      const int64_t t = T0 + tid + kHalfThreads * q;
      if (t < n_fix) y[t] = t < n_out ? acc[q] : 0.0f;   // fix_length: zero padding up to ceil(n / 2)
    }
  }
}

// librosa.to_mono = np.mean(y, axis=0) on the decoded (channels, n) array: float32 sum over channels in channel order,
// divided by the channel count.  Input here is the interleaved frame layout a WAV file stores.
__global__ void mono_mix_kernel(const float* __restrict__ interleaved, int64_t n_frames, int channels, float* __restrict__ y) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_frames; t += (int64_t)gridDim.x * blockDim.x) {
    const float* f = interleaved + t * channels;
    float acc = f[0];
    for (int c = 1; c < channels; ++c) acc += f[c];
    y[t] = __fdiv_rn(acc, (float)channels);
  }
}

static double bessel_i0(double x) {
  // power series, converges fast for x <= ~15 (beta = 14.77)
  double sum = 1.0, term = 1.0;
  const double q = 0.25 * x * x;
  for (int k = 1; k < 500; ++k) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}

static std::mutex g_rs_mutex;
static std::map<std::tuple<int, int, int>, float2*> g_rs_tables;  // (device, sr_in, sr_out)

struct PhaseTable {
  bool usable = false;
  PhaseGeom geom{};
  float* d_w = nullptr;
  int2* d_lims = nullptr;
};
static std::map<std::tuple<int, int, int>, PhaseTable> g_rs_phase;  // (device, sr_in, sr_out)
constexpr int64_t kRsMaxPhases = 2048;
constexpr size_t kRsMaxPhaseFloats = (size_t)1 << 21;  // 8 MB of weights at most (L2 resident)

static int64_t gcd64(int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; }

// Per-phase tap weights from the host copy of the (value, delta) table.
static int build_phase_table(const std::vector<float2>& tab, int sr_in, int sr_out, PhaseTable* pt) {
  const int64_t g = gcd64(sr_in, sr_out);
  const int64_t p = sr_in / g, q = sr_out / g;
  const double ratio = (double)sr_out / (double)sr_in;
  const double scale = ratio < 1.0 ? ratio : 1.0;
  const int index_step = (int)(scale * (double)kRsTable);
  if (q > kRsMaxPhases || index_step < 1) return MST_OK;  // not usable: the per-output kernel handles it
  const int max_lim = (kRsWin + index_step - 1) / index_step + 1;
  const int max_side = (max_lim + 3) & ~3;
  const int row = 2 * max_side;
  if ((size_t)q * (size_t)row > kRsMaxPhaseFloats) return MST_OK;
  std::vector<float> w((size_t)q * (size_t)row, 0.0f);
  std::vector<int2> lims((size_t)q);
  for (int64_t ph = 0; ph < q; ++ph) {
    const double pos = (double)ph / (double)q;  // fractional input position of this phase
    float* wl = w.data() + (size_t)ph * row;
    float* wr = wl + max_side;
    double frac = scale * pos;
    double index_frac = frac * (double)kRsTable;
    int offset = (int)index_frac;
    float eta = (float)(index_frac - (double)offset);
    int lim = (kRsWin - offset) / index_step;
    for (int i = 0; i < lim; ++i) {
      const float2 t = tab[(size_t)offset + (size_t)i * index_step];
      wl[i] = fmaf(eta, t.y, t.x);
    }
    lims[(size_t)ph].x = lim;
    frac = scale - frac;
    index_frac = frac * (double)kRsTable;
    offset = (int)index_frac;
    eta = (float)(index_frac - (double)offset);
    lim = (kRsWin - offset) / index_step;
    for (int k = 0; k < lim; ++k) {
      const float2 t = tab[(size_t)offset + (size_t)k * index_step];
      wr[k] = fmaf(eta, t.y, t.x);
    }
    lims[(size_t)ph].y = lim;
    if (lims[(size_t)ph].x > max_side || lims[(size_t)ph].y > max_side) return fail(MST_ERR_INVALID, "resampler phase table overflow");
  }
  MST_CUDA_OK(cudaMalloc(&pt->d_w, sizeof(float) * w.size()));
  MST_CUDA_OK(cudaMalloc(&pt->d_lims, sizeof(int2) * lims.size()));
  MST_CUDA_OK(cudaMemcpy(pt->d_w, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
  MST_CUDA_OK(cudaMemcpy(pt->d_lims, lims.data(), sizeof(int2) * lims.size(), cudaMemcpyHostToDevice));
  pt->geom.p = p; pt->geom.q = q; pt->geom.row = row; pt->geom.max_left = max_side;
  pt->usable = true;
  return MST_OK;
}

static int get_rs_table(int sr_in, int sr_out, const float2** out, const PhaseTable** phase_out) {
  int dev = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_rs_mutex);
  auto key = std::make_tuple(dev, sr_in, sr_out);
  auto it = g_rs_tables.find(key);
  if (it == g_rs_tables.end()) {
    const double ratio = (double)sr_out / (double)sr_in;
    const int n = kRsWin - 1;
    const double pi = 3.14159265358979323846, i0b = bessel_i0(kRsBeta);
    std::vector<double> win((size_t)kRsWin);
    for (int i = 0; i <= n; ++i) {
      const double t = (i == n) ? (double)kRsZeros : (double)i * ((double)kRsZeros / (double)n);  // np.linspace
      const double a = kRsRolloff * t;
      const double sinc = a == 0.0 ? 1.0 : sin(pi * a) / (pi * a);
      const double r = (double)i / (double)n;
      const double arg = 1.0 - r * r;
      const double taper = bessel_i0(kRsBeta * sqrt(arg > 0.0 ? arg : 0.0)) / i0b;
      win[i] = taper * kRsRolloff * sinc;
      if (ratio < 1.0) win[i] *= ratio;
    }
    std::vector<float2> tab((size_t)kRsWin);
    for (int i = 0; i <= n; ++i) tab[i] = make_float2((float)win[i], i < n ? (float)(win[i + 1] - win[i]) : 0.0f);
    float2* d = nullptr;
    MST_CUDA_OK(cudaMalloc(&d, sizeof(float2) * tab.size()));
    MST_CUDA_OK(cudaMemcpy(d, tab.data(), sizeof(float2) * tab.size(), cudaMemcpyHostToDevice));
    it = g_rs_tables.emplace(key, d).first;
    PhaseTable pt;
    if (sr_in != 2 * sr_out) {  // exact halving has its own kernel
      const int rc = build_phase_table(tab, sr_in, sr_out, &pt);
      if (rc) return rc;
    }
    g_rs_phase.emplace(key, pt);
  }
  *out = it->second;
  *phase_out = &g_rs_phase[key];
  return MST_OK;
}

}  // namespace mst

using namespace mst;

extern "C" {

int64_t mst_resample_length(int64_t n_in, int sr_in, int sr_out) {
  if (n_in < 0 || sr_in <= 0 || sr_out <= 0) return -1;
  if (sr_in == sr_out) return n_in;
  return (int64_t)ceil((double)n_in * ((double)sr_out / (double)sr_in));
}

int mst_mono_mix_f32(const float* d_interleaved, int64_t n_frames, int channels, float* d_out, mst_stream_t stream) {
  if (!d_interleaved || !d_out) return fail(MST_ERR_INVALID, "mst_mono_mix_f32: null argument");
  if (n_frames < 0 || channels < 1) return fail(MST_ERR_INVALID, "mst_mono_mix_f32: bad sizes");
  if (n_frames == 0) return MST_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t blocks = (n_frames + 255) / 256;
  mono_mix_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, s>>>(d_interleaved, n_frames, channels, d_out);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

int mst_resample_f32(const float* d_in, int64_t n_in, int sr_in, int sr_out, float* d_out, mst_stream_t stream) {
  if (!d_in || !d_out) return fail(MST_ERR_INVALID, "mst_resample_f32: null argument");
  if (n_in <= 0 || sr_in <= 0 || sr_out <= 0) return fail(MST_ERR_INVALID, "mst_resample_f32: bad sizes");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (sr_in == sr_out) {
    MST_CUDA_OK(cudaMemcpyAsync(d_out, d_in, sizeof(float) * (size_t)n_in, cudaMemcpyDeviceToDevice, s));
    return MST_OK;
  }
  const double ratio = (double)sr_out / (double)sr_in;
  const int64_t n_out = (int64_t)((double)n_in * ratio);
  const int64_t n_fix = mst_resample_length(n_in, sr_in, sr_out);
  const double scale = ratio < 1.0 ? ratio : 1.0;
  const int index_step = (int)(scale * (double)kRsTable);
  if (index_step < 1) return fail(MST_ERR_UNSUPPORTED, "down-sampling ratio %g too small", ratio);
  const float2* table = nullptr;
  const PhaseTable* phases = nullptr;
  int rc = get_rs_table(sr_in, sr_out, &table, &phases);
  if (rc) return rc;
  if (sr_in == 2 * sr_out) {  // one table phase for every output: the symmetric-FIR decimator
    const int64_t ctas = (n_fix + kHalfOut - 1) / kHalfOut;
    resample_half_kernel<<<(unsigned)(ctas < 148 * 16 ? ctas : 148 * 16), kHalfThreads, 0, s>>>(d_in, n_in, d_out, n_out, n_fix, table);
    MST_CUDA_OK(cudaGetLastError());
    count_launch();
    return MST_OK;
  }
  const int threads = 256;
  const int64_t blocks = (n_fix + threads - 1) / threads;
  // MST_RS_NO_PHASES=1 (environment; tests and A/B runs): per-output table interpolation for every ratio
  const char* no_phases = getenv("MST_RS_NO_PHASES");
  if (phases->usable && !(no_phases && no_phases[0] == '1') && (double)n_fix * (double)phases->geom.p < 9.0e18) {
    resample_phase_kernel<<<(unsigned)(blocks < 148 * 64 ? blocks : 148 * 64), threads, 0, s>>>(
        d_in, n_in, d_out, n_out, n_fix, phases->d_w, phases->d_lims, phases->geom);
    MST_CUDA_OK(cudaGetLastError());
    count_launch();
    return MST_OK;
  }
  resample_kernel<<<(unsigned)(blocks < 148 * 64 ? blocks : 148 * 64), threads, 0, s>>>(d_in, n_in, d_out, n_out, n_fix, table,
                                                                                 1.0 / ratio, scale, index_step);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

}  // extern "C"
