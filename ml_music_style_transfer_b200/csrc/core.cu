// Host-side plumbing of the C ABI: error strings, per-device constant tables, batch descriptors,
// the Slaney mel filterbank (host, float64 maths).  No kernels of the hot path live here.
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "mst_common.cuh"

namespace mst {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- constant tables --------------------------------------------------------------------------
static std::mutex g_table_mutex;
static Tables g_tables[64];
static bool g_tables_ready[64] = {false};

int get_tables(Tables* out) {
  int dev = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(MST_ERR_INVALID, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_table_mutex);
  if (!g_tables_ready[dev]) {
    std::vector<float2> tw1(1024), twp(kTwpCount, make_float2(0.f, 0.f));
    std::vector<float> win(kNfft), wsyn(kNfft);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k1 = 0; k1 < 32; ++k1)
      for (int n2 = 0; n2 < 32; ++n2) {
        const double a = -two_pi * (double)(k1 * n2) / 1024.0;
        tw1[k1 * 32 + n2] = make_float2((float)cos(a), (float)sin(a));
      }
    for (int k = 0; k <= 512; ++k) {
      const double a = -two_pi * (double)k / 2048.0;  // w = (cos a, sin a); -0.5i*w = (0.5 sin a, -0.5 cos a)
      twp[k] = make_float2((float)(0.5 * sin(a)), (float)(-0.5 * cos(a)));
    }
    for (int n = 0; n < kNfft; ++n) {
      win[n] = (float)(0.5 - 0.5 * cos(two_pi * (double)n / (double)kNfft));
      wsyn[n] = win[n] * (1.0f / 1024.0f);
    }
    const size_t bytes = 8192 + sizeof(float2) * kTwpCount + 8192 + 8192;
    char* d = nullptr;
    MST_CUDA_OK(cudaMalloc(&d, bytes));
    MST_CUDA_OK(cudaMemcpy(d, tw1.data(), 8192, cudaMemcpyHostToDevice));
    MST_CUDA_OK(cudaMemcpy(d + 8192, twp.data(), sizeof(float2) * kTwpCount, cudaMemcpyHostToDevice));
    MST_CUDA_OK(cudaMemcpy(d + 8192 + sizeof(float2) * kTwpCount, win.data(), 8192, cudaMemcpyHostToDevice));
    MST_CUDA_OK(cudaMemcpy(d + 16384 + sizeof(float2) * kTwpCount, wsyn.data(), 8192, cudaMemcpyHostToDevice));
    g_tables[dev].tw1024 = reinterpret_cast<const float2*>(d);
    g_tables[dev].twp = reinterpret_cast<const float2*>(d + 8192);
    g_tables[dev].window = reinterpret_cast<const float*>(d + 8192 + sizeof(float2) * kTwpCount);
    g_tables[dev].wsyn = reinterpret_cast<const float*>(d + 16384 + sizeof(float2) * kTwpCount);
    g_tables_ready[dev] = true;
  }
  *out = g_tables[dev];
  return MST_OK;
}

}  // namespace mst

using namespace mst;

extern "C" {

const char* mst_last_error(void) { return g_last_error.c_str(); }
int mst_version(void) { return 100; }
int64_t mst_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- batch descriptors --------------------------------------------------------------------
static int batch_finish(mst_batch* b) {
  // prefix sums, tile table, device copies
  int64_t frame_off = 0, acc_off = 0, samples = 0;
  int64_t tile_off = 0;
  for (int c = 0; c < b->n_clips; ++c) {
    ClipDesc& d = b->h_clips[c];
    d.frames = (int32_t)(1 + d.length / b->hop);
    d.frame_offset = frame_off;
    d.acc_offset = acc_off;
    d.tile_offset = (int32_t)tile_off;
    frame_off += d.frames;
    samples += d.length;
    int64_t acc_len = b->n_fft + (int64_t)b->hop * (d.frames - 1);
    acc_off += (acc_len + 3) & ~(int64_t)3;  // keep every clip's accumulator 16-byte aligned
    tile_off += (d.frames + kWarpsPerCta - 1) / kWarpsPerCta;
    if (tile_off > 0x7fffffff) return fail(MST_ERR_INVALID, "too many frame tiles in one batch");
  }
  b->total_frames = frame_off;
  b->audio_extent = 0;
  for (int c = 0; c < b->n_clips; ++c)
    b->audio_extent = std::max(b->audio_extent, b->h_clips[c].sample_offset + b->h_clips[c].length);
  b->total_samples = samples;
  b->total_acc = acc_off;
  b->total_tiles = (int)tile_off;
  b->uniform_frames = b->h_clips[0].frames;
  for (int c = 1; c < b->n_clips; ++c)
    if (b->h_clips[c].frames != b->uniform_frames) { b->uniform_frames = 0; break; }
  MST_CUDA_OK(cudaGetDevice(&b->device));
  MST_CUDA_OK(cudaMalloc(&b->d_clips, sizeof(ClipDesc) * (size_t)b->n_clips));
  MST_CUDA_OK(cudaMemcpy(b->d_clips, b->h_clips, sizeof(ClipDesc) * (size_t)b->n_clips, cudaMemcpyHostToDevice));
  std::vector<int32_t> tile_clip((size_t)b->total_tiles);
  for (int c = 0; c < b->n_clips; ++c) {
    const int nt = (b->h_clips[c].frames + kWarpsPerCta - 1) / kWarpsPerCta;
    for (int t = 0; t < nt; ++t) tile_clip[(size_t)b->h_clips[c].tile_offset + t] = c;
  }
  MST_CUDA_OK(cudaMalloc(&b->d_tile_clip, sizeof(int32_t) * (size_t)std::max(1, b->total_tiles)));
  MST_CUDA_OK(cudaMemcpy(b->d_tile_clip, tile_clip.data(), sizeof(int32_t) * tile_clip.size(), cudaMemcpyHostToDevice));
  return MST_OK;
}

// librosa: get_window('hann', win_length, fftbins=True) centre-padded to n_fft (util.pad_center), in double
static std::vector<double> padded_hann(int n_fft, int win_length) {
  std::vector<double> w((size_t)n_fft, 0.0);
  const double two_pi = 6.283185307179586476925286766559;
  const int lpad = (n_fft - win_length) / 2;
  for (int n = 0; n < win_length; ++n) w[(size_t)lpad + n] = 0.5 - 0.5 * cos(two_pi * (double)n / (double)win_length);
  return w;
}

// Non-default window length or FFT size: upload the padded analysis / synthesis windows of this batch.
static int batch_upload_window(mst_batch* b) {
  if (b->win_length == kNfft && b->n_fft == kNfft) return MST_OK;  // kernels use the per-device default tables
  const int n_fft = b->n_fft;
  const std::vector<double> w = padded_hann(n_fft, b->win_length);
  std::vector<float> wf((size_t)n_fft), ws((size_t)n_fft);
  for (int n = 0; n < n_fft; ++n) {
    wf[n] = (float)w[n];
    ws[n] = wf[n] * (1.0f / (float)(n_fft / 2));
  }
  MST_CUDA_OK(cudaMalloc(&b->d_window, sizeof(float) * n_fft));
  MST_CUDA_OK(cudaMalloc(&b->d_wsyn, sizeof(float) * n_fft));
  MST_CUDA_OK(cudaMemcpy(b->d_window, wf.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
  MST_CUDA_OK(cudaMemcpy(b->d_wsyn, ws.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
  return MST_OK;
}

static int batch_check(int n_clips, int n_fft, int hop, int win_length, int pad_mode) {
  if (n_clips <= 0) return fail(MST_ERR_INVALID, "n_clips must be positive (got %d)", n_clips);
  if (n_fft != kNfft && !generic_n_fft_ok(n_fft))
    return fail(MST_ERR_UNSUPPORTED, "n_fft=%d unsupported: n_fft must be a power of two in [64, 16384]", n_fft);
  if (win_length < 1 || win_length > n_fft) return fail(MST_ERR_INVALID, "win_length=%d must be in [1, n_fft]", win_length);
  if (hop <= 0 || hop > n_fft) return fail(MST_ERR_INVALID, "hop=%d must be in [1, n_fft]", hop);
  if (pad_mode != MST_PAD_REFLECT && pad_mode != MST_PAD_CONSTANT) return fail(MST_ERR_INVALID, "bad pad_mode %d", pad_mode);
  return MST_OK;
}

int mst_batch_create(int n_clips, const int64_t* h_clip_offsets, const int64_t* h_clip_lengths, int n_fft, int hop,
                     int pad_mode, mst_batch_t** out) {
  return mst_batch_create_ex(n_clips, h_clip_offsets, h_clip_lengths, n_fft, hop, n_fft, pad_mode, out);
}

int mst_batch_create_ex(int n_clips, const int64_t* h_clip_offsets, const int64_t* h_clip_lengths, int n_fft, int hop,
                        int win_length, int pad_mode, mst_batch_t** out) {
  if (!out || !h_clip_offsets || !h_clip_lengths) return fail(MST_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = batch_check(n_clips, n_fft, hop, win_length, pad_mode);
  if (rc) return rc;
  for (int c = 0; c < n_clips; ++c) {
    if (h_clip_offsets[c] < 0) return fail(MST_ERR_INVALID, "clip %d: negative offset", c);
    // np.pad(mode='reflect') needs len > pad width; librosa raises for shorter inputs.
    if (pad_mode == MST_PAD_REFLECT && h_clip_lengths[c] <= n_fft / 2)
      return fail(MST_ERR_INVALID, "clip %d: length %lld too short for reflect padding of %d", c,
                  (long long)h_clip_lengths[c], n_fft / 2);
    if (h_clip_lengths[c] <= 0) return fail(MST_ERR_INVALID, "clip %d: empty clip", c);
  }
  mst_batch* b = new mst_batch();
  b->n_clips = n_clips; b->n_fft = n_fft; b->hop = hop; b->pad_mode = pad_mode; b->win_length = win_length;
  b->h_clips = new ClipDesc[n_clips];
  for (int c = 0; c < n_clips; ++c) {
    b->h_clips[c].sample_offset = h_clip_offsets[c];
    b->h_clips[c].length = h_clip_lengths[c];
  }
  rc = batch_finish(b);
  if (!rc) rc = batch_upload_window(b);
  if (rc) { mst_batch_destroy(b); return rc; }
  *out = b;
  return MST_OK;
}

int mst_batch_create_from_frames(int n_clips, const int64_t* h_frames, int n_fft, int hop, int pad_mode,
                                 mst_batch_t** out) {
  return mst_batch_create_from_frames_ex(n_clips, h_frames, n_fft, hop, n_fft, pad_mode, out);
}

int mst_batch_create_from_frames_ex(int n_clips, const int64_t* h_frames, int n_fft, int hop, int win_length, int pad_mode,
                                    mst_batch_t** out) {
  if (!out || !h_frames) return fail(MST_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = batch_check(n_clips, n_fft, hop, win_length, pad_mode);
  if (rc) return rc;
  mst_batch* b = new mst_batch();
  b->n_clips = n_clips; b->n_fft = n_fft; b->hop = hop; b->pad_mode = pad_mode; b->from_frames = true;
  b->win_length = win_length;
  b->h_clips = new ClipDesc[n_clips];
  int64_t off = 0;
  for (int c = 0; c < n_clips; ++c) {
    const int64_t T = h_frames[c];
    const int64_t len = (int64_t)hop * (T - 1);
    if (T < 1 || (pad_mode == MST_PAD_REFLECT && len <= n_fft / 2)) {
      delete[] b->h_clips; delete b;
      return fail(MST_ERR_INVALID, "clip %d: %lld frames is too short (hop*(T-1) must exceed n_fft/2)", c, (long long)T);
    }
    b->h_clips[c].sample_offset = off;
    b->h_clips[c].length = len;
    off += len;
  }
  rc = batch_finish(b);
  if (!rc) rc = batch_upload_window(b);
  if (rc) { mst_batch_destroy(b); return rc; }
  // Griffin-Lim needs 1 / window-sum-square per accumulator position (librosa.istft's normalisation,
  // accumulated in float32 frame by frame like librosa's window_sumsquare).  Clips with equal T share one envelope.
  {
    const std::vector<double> wd = padded_hann(n_fft, win_length);
    std::vector<double> wsq(n_fft);
    for (int n = 0; n < n_fft; ++n) wsq[n] = wd[n] * wd[n];
    std::vector<int64_t> wss_off((size_t)n_clips);
    std::vector<float> env;
    std::vector<std::pair<int32_t, int64_t>> seen;  // (T, offset)
    for (int c = 0; c < n_clips; ++c) {
      const int32_t T = b->h_clips[c].frames;
      int64_t found = -1;
      for (auto& s : seen) if (s.first == T) { found = s.second; break; }
      if (found < 0) {
        found = (int64_t)env.size();
        const int64_t n = n_fft + (int64_t)hop * (T - 1);
        const int64_t n_pad = (n + 3) & ~(int64_t)3;
        env.resize(env.size() + (size_t)n_pad, 0.0f);
        float* x = env.data() + found;
        for (int32_t t = 0; t < T; ++t) {
          float* p = x + (int64_t)t * hop;
          for (int j = 0; j < n_fft; ++j) p[j] = (float)((double)p[j] + wsq[j]);
        }
        const float tiny = 1.17549435e-38f;
        for (int64_t i = 0; i < n; ++i) x[i] = x[i] > tiny ? 1.0f / x[i] : 1.0f;
        seen.push_back({T, found});
      }
      wss_off[c] = found;
    }
    // Interior frames (all n_fft/hop overlapping neighbours present) see an envelope that is periodic in hop; fold it
    // into the analysis window once.  Computed like the envelopes above (float32 accumulation in frame order).
    std::vector<float> wq(n_fft, 0.0f);
    {
      const int q = (n_fft - 1) / hop;
      const int32_t Tq = 2 * q + 1;  // smallest clip with one interior frame (frame q)
      std::vector<float> x((size_t)n_fft + (size_t)hop * (Tq - 1), 0.0f);
      for (int32_t t = 0; t < Tq; ++t) {
        float* p = x.data() + (int64_t)t * hop;
        for (int j = 0; j < n_fft; ++j) p[j] = (float)((double)p[j] + wsq[j]);
      }
      const float tiny = 1.17549435e-38f;
      for (int j = 0; j < n_fft; ++j) {
        const float e = x[(size_t)q * hop + j];
        const float w = (float)wd[j];
        wq[j] = w * (e > tiny ? 1.0f / e : 1.0f);
      }
    }
    if (cudaMalloc(&b->d_wq, sizeof(float) * n_fft) != cudaSuccess ||
        cudaMemcpy(b->d_wq, wq.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&b->d_inv_wss, sizeof(float) * env.size()) != cudaSuccess ||
        cudaMemcpy(b->d_inv_wss, env.data(), sizeof(float) * env.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&b->d_wss_offset, sizeof(int64_t) * (size_t)n_clips) != cudaSuccess ||
        cudaMemcpy(b->d_wss_offset, wss_off.data(), sizeof(int64_t) * (size_t)n_clips, cudaMemcpyHostToDevice) != cudaSuccess) {
      mst_batch_destroy(b);
      return fail(MST_ERR_CUDA, "allocating the window-sum-square envelope failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  *out = b;
  return MST_OK;
}

void mst_batch_destroy(mst_batch_t* b) {
  if (!b) return;
  if (b->d_clips) cudaFree(b->d_clips);
  if (b->d_tile_clip) cudaFree(b->d_tile_clip);
  if (b->d_inv_wss) cudaFree(b->d_inv_wss);
  if (b->d_wq) cudaFree(b->d_wq);
  if (b->d_window) cudaFree(b->d_window);
  if (b->d_wsyn) cudaFree(b->d_wsyn);
  if (b->d_wss_offset) cudaFree(b->d_wss_offset);
  delete[] b->h_clips;
  delete b;
}
int mst_batch_n_clips(const mst_batch_t* b) { return b ? b->n_clips : 0; }
int64_t mst_batch_total_frames(const mst_batch_t* b) { return b ? b->total_frames : 0; }
int64_t mst_batch_total_samples(const mst_batch_t* b) { return b ? b->total_samples : 0; }
int64_t mst_batch_audio_extent(const mst_batch_t* b) { return b ? b->audio_extent : 0; }
int mst_batch_device(const mst_batch_t* b) { return b ? b->device : -1; }
int mst_batch_n_fft(const mst_batch_t* b) { return b ? b->n_fft : 0; }
int64_t mst_batch_clip_frames(const mst_batch_t* b, int c) { return (b && c >= 0 && c < b->n_clips) ? b->h_clips[c].frames : -1; }
int64_t mst_batch_frame_offset(const mst_batch_t* b, int c) {
  if (!b || c < 0 || c > b->n_clips) return -1;
  return c == b->n_clips ? b->total_frames : b->h_clips[c].frame_offset;
}

// ---- Slaney mel filterbank (librosa.filters.mel, htk=False, norm='slaney') --------------------------
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

int mst_mel_filterbank_f32(int sr, int n_fft, int n_mels, double fmin, double fmax, float* W) {
  if (!W || sr <= 0 || n_fft < 2 || n_mels < 1) return fail(MST_ERR_INVALID, "bad mel filterbank arguments");
  if (fmax <= 0) fmax = 0.5 * (double)sr;
  const int K = 1 + n_fft / 2;
  std::vector<double> mel_f((size_t)n_mels + 2), fftf((size_t)K);
  // np.linspace(a, b, n): a + i*step, last point forced to b
  const double mmin = hz_to_mel(fmin), mmax = hz_to_mel(fmax);
  for (int i = 0; i < n_mels + 2; ++i) {
    const double m = (i == n_mels + 1) ? mmax : mmin + (double)i * ((mmax - mmin) / (double)(n_mels + 1));
    mel_f[i] = mel_to_hz(m);
  }
  for (int k = 0; k < K; ++k) fftf[k] = (k == K - 1) ? 0.5 * (double)sr : (double)k * ((0.5 * (double)sr) / (double)(K - 1));
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < K; ++k) {
      const double lower = -(mel_f[i] - fftf[k]) / fd0;
      const double upper = (mel_f[i + 2] - fftf[k]) / fd1;
      double w = lower < upper ? lower : upper;
      if (!(w > 0.0)) w = 0.0;
      const float w32 = (float)w;                       // weights[i] = ... (float32 store)
      W[(size_t)i * K + k] = (float)((double)w32 * enorm);  // weights *= enorm[:, None] (f64 product, f32 store)
    }
  }
  return MST_OK;
}

}  // extern "C"
