// Mel inversion: librosa.feature.inverse.mel_to_stft (the first half of mel_to_audio, the alternative kept as a
// comment at reference tests/test_griffinlim.py:24):  S = nnls(mel_basis, M) ** (1 / power).
//
// librosa solves  min_X 0.5 ||A X - M||^2, X >= 0  (A = mel filterbank, n_mels x n_bins) column block by column block
// with SciPy's L-BFGS-B, started from the clipped least-squares solution max(pinv(A) M, 0).  The problem is separable:
// every frame (column) is an independent n_bins-variable NNLS problem, and A is extremely sparse -- a frequency bin lies
// under at most two overlapping triangular filters.  One WARP owns one frame: its 1025 unknowns live in registers (33
// per lane), the residual r = A y - m (n_mels floats) in shared memory.  Start point: the same clipped least-squares
// solution (pinv(A) = A^T (A A^T)^-1, factorised once per plan on the host in double precision).  Iteration:
// accelerated projected gradient (FISTA, step 1 / lambda_max(A A^T), restart when the residual grows); cond(A A^T) is
// ~30 for the Slaney bank, so ~60 iterations reach float32 precision.  Where L-BFGS-B stops early on its scaled
// projected-gradient test, this solver runs to convergence: same objective, same start, residual <= librosa's.
#include <algorithm>
#include <cmath>
#include <vector>
#include "mst_common.cuh"

namespace mst {
constexpr int kMaxPerBin = 4;     // filters overlapping one frequency bin (2 for triangular mel banks)
constexpr int kInvWarps = 8;
}  // namespace mst

struct mst_mel_inverse_plan {
  int n_mels = 0, n_bins = 0;
  float* d_pinv_t = nullptr;      // [n_mels][n_bins]: transpose of pinv(A), coalesced over bins
  int16_t* d_col_rows = nullptr;  // [n_bins][kMaxPerBin]: mel rows with a non-zero weight at this bin (-1 = none)
  float* d_col_vals = nullptr;    // [n_bins][kMaxPerBin]
  float inv_lipschitz = 0.0f;     // 1 / lambda_max(A A^T)
};

namespace mst {

__global__ void __launch_bounds__(kInvWarps * 32)
mel_nnls_kernel(const float* __restrict__ mel, int mel_layout, const ClipDesc* __restrict__ clips, int n_clips,
                int64_t total_frames, int n_mels, const float* __restrict__ pinv_t, const int16_t* __restrict__ col_rows,
                const float* __restrict__ col_vals, float inv_L, float inv_power, int max_iter, float tol,
                float* __restrict__ S_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int16_t* s_rows = reinterpret_cast<int16_t*>(smem_raw);                                   // [kBins][kMaxPerBin]
  float* s_vals = reinterpret_cast<float*>(smem_raw + ((kBins * kMaxPerBin * 2 + 15) & ~15));  // [kBins][kMaxPerBin]
  float* s_r_all = s_vals + kBins * kMaxPerBin;                                             // [kInvWarps][n_mels]
  float* s_b_all = s_r_all + kInvWarps * n_mels;                                            // [kInvWarps][n_mels]
  for (int i = threadIdx.x; i < kBins * kMaxPerBin; i += blockDim.x) {
    s_rows[i] = col_rows[i];
    s_vals[i] = col_vals[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_r = s_r_all + warp * n_mels;
  float* s_b = s_b_all + warp * n_mels;
  for (int64_t g = (int64_t)blockIdx.x * kInvWarps + warp; g < total_frames; g += (int64_t)gridDim.x * kInvWarps) {
    // clip of global frame g (clips are few and sorted: binary search)
    int lo = 0, hi = n_clips;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (clips[mid].frame_offset <= g) lo = mid; else hi = mid;
    }
    const int64_t f0 = clips[lo].frame_offset;
    const int T = clips[lo].frames, t = (int)(g - f0);
    float bb = 0.0f;
    for (int m = lane; m < n_mels; m += 32) {
      const float v = mel_layout == MST_LAYOUT_FRAME_MAJOR ? mel[g * n_mels + m] : mel[f0 * n_mels + (int64_t)m * T + t];
      s_b[m] = v;
      bb = fmaf(v, v, bb);
    }
    for (int o = 16; o; o >>= 1) bb += __shfl_xor_sync(0xffffffffu, bb, o);
    __syncwarp();
    // start point: max(pinv(A) m, 0)
    float x[33], xo[33];
#pragma unroll
    for (int j = 0; j < 33; ++j) x[j] = 0.0f;
    for (int m = 0; m < n_mels; ++m) {
      const float bm = s_b[m];
      const float* row = pinv_t + (size_t)m * kBins + lane;
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = fmaf(__ldg(row + 32 * j), bm, x[j]);
      if (lane == 0) x[32] = fmaf(__ldg(row + 1024), bm, x[32]);
    }
#pragma unroll
    for (int j = 0; j < 33; ++j) { x[j] = fmaxf(x[j], 0.0f); xo[j] = x[j]; }
    float tk = 1.0f, beta = 0.0f, prev_res = 3.4e38f;
    for (int it = 0; it < max_iter; ++it) {
      for (int m = lane; m < n_mels; m += 32) s_r[m] = -s_b[m];
      __syncwarp();
      // r = A y - m with y = x + beta (x - x_old); every lane scatters the contributions of its own bins
#pragma unroll
      for (int j = 0; j < 33; ++j) {
        if (j == 32 && lane != 0) break;
        const int k = j < 32 ? lane + 32 * j : 1024;
        const float y = fmaf(beta, x[j] - xo[j], x[j]);
#pragma unroll
        for (int e = 0; e < kMaxPerBin; ++e) {
          const int m = s_rows[k * kMaxPerBin + e];
          if (m >= 0) atomicAdd(s_r + m, s_vals[k * kMaxPerBin + e] * y);
        }
      }
      __syncwarp();
      float res = 0.0f;
      for (int m = lane; m < n_mels; m += 32) res = fmaf(s_r[m], s_r[m], res);
      for (int o = 16; o; o >>= 1) res += __shfl_xor_sync(0xffffffffu, res, o);
      if (res <= tol * tol * bb) break;        // converged (warp-uniform: res is the same in every lane)
      const bool restart = res > prev_res;     // adaptive restart of the momentum
      prev_res = res;
      const float tn = restart ? 1.0f : 0.5f * (1.0f + sqrtf(fmaf(4.0f * tk, tk, 1.0f)));
      const float beta_next = restart ? 0.0f : (tk - 1.0f) / tn;
      // x_new = max(y - (A^T r) / L, 0)
#pragma unroll
      for (int j = 0; j < 33; ++j) {
        if (j == 32 && lane != 0) break;
        const int k = j < 32 ? lane + 32 * j : 1024;
        const float y = fmaf(beta, x[j] - xo[j], x[j]);
        float gk = 0.0f;
#pragma unroll
        for (int e = 0; e < kMaxPerBin; ++e) {
          const int m = s_rows[k * kMaxPerBin + e];
          if (m >= 0) gk = fmaf(s_vals[k * kMaxPerBin + e], s_r[m], gk);
        }
        xo[j] = x[j];
        x[j] = fmaxf(fmaf(-inv_L, gk, y), 0.0f);
      }
      tk = tn;
      beta = beta_next;
      __syncwarp();
    }
    float* out = S_out + g * kBins;
#pragma unroll
    for (int j = 0; j < 33; ++j) {
      if (j == 32 && lane != 0) break;
      const int k = j < 32 ? lane + 32 * j : 1024;
      out[k] = inv_power == 0.5f ? sqrtf(x[j]) : (inv_power == 1.0f ? x[j] : powf(x[j], inv_power));
    }
    __syncwarp();
  }
}

}  // namespace mst

using namespace mst;

extern "C" {

int mst_mel_inverse_plan_create(const float* W, int n_mels, int n_bins, mst_mel_inverse_plan_t** out) {
  if (!W || !out) return fail(MST_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n_bins != kBins) return fail(MST_ERR_UNSUPPORTED, "mel inversion needs n_bins=1025 (n_fft=2048), got %d", n_bins);
  if (n_mels < 1 || n_mels > 256) return fail(MST_ERR_UNSUPPORTED, "n_mels=%d outside [1,256]", n_mels);
  // column lists
  std::vector<int16_t> rows((size_t)n_bins * kMaxPerBin, (int16_t)-1);
  std::vector<float> vals((size_t)n_bins * kMaxPerBin, 0.0f);
  for (int k = 0; k < n_bins; ++k) {
    int e = 0;
    for (int m = 0; m < n_mels; ++m) {
      const float w = W[(size_t)m * n_bins + k];
      if (w == 0.0f) continue;
      if (e == kMaxPerBin) return fail(MST_ERR_UNSUPPORTED, "bin %d lies under more than %d filters", k, kMaxPerBin);
      rows[(size_t)k * kMaxPerBin + e] = (int16_t)m;
      vals[(size_t)k * kMaxPerBin + e] = w;
      ++e;
    }
  }
  // G = A A^T over the non-empty filters (double), Cholesky, pinv(A) = A^T G^-1 (the minimum-norm least-squares map;
  // an empty filter -- librosa warns about those -- gets a zero column)
  std::vector<int> live;
  for (int m = 0; m < n_mels; ++m) {
    bool any = false;
    for (int k = 0; k < n_bins && !any; ++k) any = W[(size_t)m * n_bins + k] != 0.0f;
    if (any) live.push_back(m);
  }
  const int n = (int)live.size();
  std::vector<double> G((size_t)n * n, 0.0);
  for (int a = 0; a < n; ++a)
    for (int b2 = a; b2 < n; ++b2) {
      double acc = 0.0;
      const float* ra = W + (size_t)live[a] * n_bins;
      const float* rb = W + (size_t)live[b2] * n_bins;
      for (int k = 0; k < n_bins; ++k) acc += (double)ra[k] * (double)rb[k];
      G[(size_t)a * n + b2] = G[(size_t)b2 * n + a] = acc;
    }
  // lambda_max by power iteration
  double lam = 0.0;
  {
    std::vector<double> v((size_t)n, 1.0), w2((size_t)n);
    for (int it = 0; it < 500; ++it) {
      double nrm = 0.0;
      for (int a = 0; a < n; ++a) {
        double acc = 0.0;
        for (int b2 = 0; b2 < n; ++b2) acc += G[(size_t)a * n + b2] * v[b2];
        w2[a] = acc;
        nrm += acc * acc;
      }
      nrm = std::sqrt(nrm);
      if (nrm == 0.0) break;
      for (int a = 0; a < n; ++a) v[a] = w2[a] / nrm;
      lam = nrm;
    }
  }
  if (!(lam > 0.0)) return fail(MST_ERR_INVALID, "filterbank is all zeros");
  std::vector<double> Lc(G);  // in-place Cholesky (lower)
  for (int j = 0; j < n; ++j) {
    for (int k2 = 0; k2 < j; ++k2)
      for (int i = j; i < n; ++i) Lc[(size_t)i * n + j] -= Lc[(size_t)i * n + k2] * Lc[(size_t)j * n + k2];
    const double d = Lc[(size_t)j * n + j];
    if (!(d > 0.0)) return fail(MST_ERR_INVALID, "A A^T is not positive definite (linearly dependent filters)");
    const double sd = std::sqrt(d);
    for (int i = j; i < n; ++i) Lc[(size_t)i * n + j] /= sd;
  }
  // solve G Z = A (n x n_bins) column by column of A^T: pinv^T = G^-1 A  ->  stored as [n_mels][n_bins]
  std::vector<float> pinv_t((size_t)n_mels * n_bins, 0.0f);
  std::vector<double> col((size_t)n);
  for (int k = 0; k < n_bins; ++k) {
    for (int a = 0; a < n; ++a) col[a] = (double)W[(size_t)live[a] * n_bins + k];
    for (int i = 0; i < n; ++i) {  // forward
      double acc = col[i];
      for (int j = 0; j < i; ++j) acc -= Lc[(size_t)i * n + j] * col[j];
      col[i] = acc / Lc[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {  // backward (L^T)
      double acc = col[i];
      for (int j = i + 1; j < n; ++j) acc -= Lc[(size_t)j * n + i] * col[j];
      col[i] = acc / Lc[(size_t)i * n + i];
    }
    for (int a = 0; a < n; ++a) pinv_t[(size_t)live[a] * n_bins + k] = (float)col[a];
  }
  mst_mel_inverse_plan* p = new mst_mel_inverse_plan();
  p->n_mels = n_mels; p->n_bins = n_bins; p->inv_lipschitz = (float)(1.0 / lam);
  if (cudaMalloc(&p->d_pinv_t, pinv_t.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&p->d_col_rows, rows.size() * sizeof(int16_t)) != cudaSuccess ||
      cudaMalloc(&p->d_col_vals, vals.size() * sizeof(float)) != cudaSuccess ||
      cudaMemcpy(p->d_pinv_t, pinv_t.data(), pinv_t.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_col_rows, rows.data(), rows.size() * sizeof(int16_t), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_col_vals, vals.data(), vals.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    mst_mel_inverse_plan_destroy(p);
    return fail(MST_ERR_CUDA, "mel inverse plan upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = p;
  return MST_OK;
}

void mst_mel_inverse_plan_destroy(mst_mel_inverse_plan_t* p) {
  if (!p) return;
  if (p->d_pinv_t) cudaFree(p->d_pinv_t);
  if (p->d_col_rows) cudaFree(p->d_col_rows);
  if (p->d_col_vals) cudaFree(p->d_col_vals);
  delete p;
}

int mst_mel_to_stft_f32(const float* d_mel, int mel_layout, const mst_batch_t* b, const mst_mel_inverse_plan_t* plan,
                        float power, int max_iter, float tol, float* d_S_out, mst_stream_t stream) {
  if (!d_mel || !b || !plan || !d_S_out) return fail(MST_ERR_INVALID, "mst_mel_to_stft_f32: null argument");
  if (mel_layout != MST_LAYOUT_FRAME_MAJOR && mel_layout != MST_LAYOUT_BIN_MAJOR) return fail(MST_ERR_INVALID, "bad layout %d", mel_layout);
  if (!(power > 0.0f) || max_iter < 0 || !(tol >= 0.0f)) return fail(MST_ERR_INVALID, "bad power / max_iter / tol");
  if (b->n_fft != kNfft) return fail(MST_ERR_UNSUPPORTED, "mel inversion is built for n_fft=2048 (batch has n_fft=%d)", b->n_fft);
  if (b->total_frames == 0) return MST_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = ((kBins * kMaxPerBin * 2 + 15) & ~15) + sizeof(float) * kBins * kMaxPerBin +
                      2 * sizeof(float) * kInvWarps * (size_t)plan->n_mels;
  static_assert(((1025 * 4 * 2 + 15) & ~15) + 4 * 1025 * 4 + 2 * 4 * 8 * 256 <= 48 * 1024, "static smem budget");
  int dev = 0, sms = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  MST_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t want = (b->total_frames + kInvWarps - 1) / kInvWarps;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sms * 4);
  mel_nnls_kernel<<<grid, kInvWarps * 32, smem, s>>>(d_mel, mel_layout, b->d_clips, b->n_clips, b->total_frames, plan->n_mels,
                                                    plan->d_pinv_t, plan->d_col_rows, plan->d_col_vals, plan->inv_lipschitz,
                                                    1.0f / power, max_iter, tol, d_S_out);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

}  // extern "C"
