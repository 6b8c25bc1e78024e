// P1 / P2: framing + Hann window + 2048-point real FFT with the magnitude / power / log1p / mel epilogue
// fused in.  One warp per frame, persistent CTAs of 8 warps walking tiles of 8 consecutive frames of one clip.
//
// Replaces librosa.stft + np.log1p(np.abs(.)**2) (reference preprocessing/preprocess.py:47-57) and
// librosa.feature.melspectrogram (reference tests/plot_spec.py:20).
#include <algorithm>
#include <vector>
#include "fft_warp.cuh"
#include "mst_common.cuh"

namespace mst {

constexpr int kModeMel = 4;           // internal epilogue id (after the public MST_OUT_* ids 0..3)
constexpr int kTileStride = 1028;     // floats per frame row of the bin-major staging tile (== 4 mod 32: conflict-free)

struct MelDev {            // device view of a mel plan (banded-compact filterbank)
  const float* w;          // compact weights, row m starts at woff[m]
  const int32_t* klo;      // first non-zero bin of row m
  const int32_t* kcnt;     // number of bins in the band of row m
  const int32_t* woff;
  int n_mels;
  int apply_log1p;
};

// Load one frame (centre-padded, reflect or zero) as z[m] = x[2m] + i*x[2m+1], m = 32*r + lane, times the window.
__device__ __forceinline__ void load_frame(float2 (&v)[32], const float* __restrict__ clip, int64_t len, int64_t base,
                                           int pad_mode, const float* s_window, float2* scratch, int lane) {
  const float2* w2 = reinterpret_cast<const float2*>(s_window);
  const bool interior = base >= 0 && base + kNfft <= len;
  if (interior && ((reinterpret_cast<uintptr_t>(clip + base) & 7) == 0)) {
    const float2* src = reinterpret_cast<const float2*>(clip + base);
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = __ldg(src + 32 * r + lane);
  } else {
    // edge or unaligned frame: stage through this warp's scratch with a compact (not unrolled) loop
    float* sf = reinterpret_cast<float*>(scratch);
#pragma unroll 1
    for (int jj = lane; jj < kNfft; jj += 32) {
      int64_t i = base + jj;
      bool ok = true;
      if (i < 0) {
        if (pad_mode == MST_PAD_REFLECT) i = -i; else ok = false;
      } else if (i >= len) {
        if (pad_mode == MST_PAD_REFLECT) i = 2 * (len - 1) - i; else ok = false;
      }
      sf[jj] = ok ? __ldg(clip + i) : 0.0f;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = scratch[32 * r + lane];
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const float2 w = w2[32 * r + lane];
    v[r].x *= w.x;
    v[r].y *= w.y;
  }
}

template <int MODE>
__device__ __forceinline__ float epilogue_value(float2 x) {
  const float p = fmaf(x.x, x.x, x.y * x.y);
  if (MODE == MST_OUT_MAGNITUDE) return sqrtf(p);
  if (MODE == MST_OUT_LOG1P_POWER) return fast_log1p(p);
  return p;
}

constexpr size_t kStftSmemBytes = 8192 + sizeof(float2) * kTwpCount + 8192 + sizeof(float2) * kScratchPerWarp * kWarpsPerCta;

template <int MODE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
stft_kernel(const float* __restrict__ audio, const ClipDesc* __restrict__ clips, const int32_t* __restrict__ tile_clip,
            int total_tiles, int hop, int pad_mode, Tables tabs, int layout, void* __restrict__ out_v, MelDev mel) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tw1024 = reinterpret_cast<float2*>(smem_raw);
  float2* s_twp = s_tw1024 + 1024;
  float* s_window = reinterpret_cast<float*>(s_twp + kTwpCount);
  float2* s_scratch_all = reinterpret_cast<float2*>(s_window + kNfft);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = mirror_base(lane);
  float2* scratch = s_scratch_all + warp * kScratchPerWarp;
  float* s_tile = reinterpret_cast<float*>(s_scratch_all);  // bin-major staging tile aliases the scratch region

  stage_table(s_tw1024, tabs.tw1024, 512);
  stage_table(s_twp, tabs.twp, kTwpCount / 2);
  stage_table(s_window, tabs.window, 512);
  __syncthreads();

  const int n_out = (MODE == kModeMel) ? mel.n_mels : kBins;  // values per frame in the output

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int c = __ldg(tile_clip + tile);
    const ClipDesc cd = clips[c];
    const int t0 = (tile - cd.tile_offset) * kWarpsPerCta;
    const int t = t0 + warp;
    const bool active = t < cd.frames;
    {  // pull the audio of this CTA's next tile towards L2 while this one computes
      const int nt = tile + gridDim.x;
      if (nt < total_tiles && warp == 0) {
        const int c2 = __ldg(tile_clip + nt);
        const ClipDesc cn = clips[c2];
        const int64_t b0 = (int64_t)(nt - cn.tile_offset) * kWarpsPerCta * hop - kHalf;
        const int64_t span = (int64_t)(kWarpsPerCta - 1) * hop + kNfft;
        for (int64_t i = (int64_t)lane * 32; i < span; i += 32 * 32) {
          const int64_t idx = b0 + i;
          if (idx >= 0 && idx < cn.length) prefetch_l2(audio + cn.sample_offset + idx);
        }
      }
    }
    float2 o[32];
    float2 mid = make_float2(0.0f, 0.0f);
    if (active) {
      float2 v[32];
      load_frame(v, audio + cd.sample_offset, cd.length, (int64_t)t * hop - kHalf, pad_mode, s_window, scratch, lane);
      rfft2048_warp(v, o, &mid, scratch, s_tw1024, s_twp, lane);
    }
    const int64_t g = cd.frame_offset + t;  // global frame id

    if (MODE == MST_OUT_COMPLEX) {
      if (active) {
        float2* out = reinterpret_cast<float2*>(out_v) + g * kBins;
#pragma unroll
        for (int j = 0; j < 32; ++j) out[mirror_bin(lane, kb, j)] = o[j];
        if (lane == 0) out[512] = mid;
      }
      continue;
    }

    // real-valued epilogues ------------------------------------------------------------------
    float* out = reinterpret_cast<float*>(out_v);
    if (MODE == kModeMel) {
      // power spectrum of this frame -> this warp's scratch (linear [k]), then banded mel rows per lane
      float* pw = reinterpret_cast<float*>(scratch);
      __syncwarp();
      if (active) {
#pragma unroll
        for (int j = 0; j < 32; ++j) pw[mirror_bin(lane, kb, j)] = fmaf(o[j].x, o[j].x, o[j].y * o[j].y);
        if (lane == 0) pw[512] = fmaf(mid.x, mid.x, mid.y * mid.y);
      }
      __syncwarp();
      float melv[8];  // n_mels <= 256
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        melv[i] = 0.0f;
        const int m = lane + 32 * i;
        if (active && m < mel.n_mels) {
          const int k0 = __ldg(mel.klo + m), cnt = __ldg(mel.kcnt + m);
          const float* wr = mel.w + __ldg(mel.woff + m);
          float acc = 0.0f;
          for (int j = 0; j < cnt; ++j) acc = fmaf(__ldg(wr + j), pw[k0 + j], acc);
          melv[i] = mel.apply_log1p ? fast_log1p(acc) : acc;
        }
      }
      __syncwarp();
      if (layout == MST_LAYOUT_FRAME_MAJOR) {
        if (active) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = lane + 32 * i;
            if (m < mel.n_mels) out[g * mel.n_mels + m] = melv[i];
          }
        }
      } else {
        __syncthreads();  // every warp is done with its scratch before the tile overwrites it
        if (active) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = lane + 32 * i;
            if (m < mel.n_mels) s_tile[warp * kTileStride + m] = melv[i];
          }
        }
      }
    } else {
      if (layout == MST_LAYOUT_FRAME_MAJOR) {
        if (active) {
          float* row = out + g * kBins;
#pragma unroll
          for (int j = 0; j < 32; ++j) row[mirror_bin(lane, kb, j)] = epilogue_value<MODE>(o[j]);
          if (lane == 0) row[512] = epilogue_value<MODE>(mid);
        }
      } else {
        __syncthreads();
        if (active) {
#pragma unroll
          for (int j = 0; j < 32; ++j) s_tile[warp * kTileStride + mirror_bin(lane, kb, j)] = epilogue_value<MODE>(o[j]);
          if (lane == 0) s_tile[warp * kTileStride + 512] = epilogue_value<MODE>(mid);
        }
      }
    }
    if (layout == MST_LAYOUT_BIN_MAJOR) {
      // transposed store: clip block is [n_out][T]; 8 consecutive frames give 32-byte row segments
      __syncthreads();
      const int nvalid = min(kWarpsPerCta, cd.frames - t0);
      const int f = threadIdx.x & 7, kk = threadIdx.x >> 3;
      float* blk = out + cd.frame_offset * n_out;
      if (f < nvalid) {
        for (int k = kk; k < n_out; k += 32) blk[(int64_t)k * cd.frames + t0 + f] = s_tile[f * kTileStride + k];
      }
      __syncthreads();
    }
  }
}

template <int MODE>
static int launch_stft(const float* d_audio, const mst_batch* b, int layout, void* d_out, const MelDev& mel,
                       cudaStream_t stream) {
  Tables tabs;
  int rc = get_tables(&tabs);
  if (rc) return rc;
  const size_t smem = kStftSmemBytes;
  static bool attr_set[64] = {false};
  int dev = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  if (!attr_set[dev]) {
    MST_CUDA_OK(cudaFuncSetAttribute(stft_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev] = true;
  }
  int sms = 0;
  MST_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(b->total_tiles, 2 * sms);
  stft_kernel<MODE><<<grid, kWarpsPerCta * 32, smem, stream>>>(d_audio, b->d_clips, b->d_tile_clip, b->total_tiles,
                                                              b->hop, b->pad_mode, tabs, layout, d_out, mel);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

}  // namespace mst

using namespace mst;

struct mst_mel_plan {
  int n_mels = 0, n_bins = 0;
  float* d_w = nullptr;      // compact band weights
  int32_t* d_meta = nullptr; // klo[n_mels], kcnt[n_mels], woff[n_mels]
  float* d_dense = nullptr;  // dense [n_mels][n_bins] copy (kept for later tensor-core forms)
};

extern "C" {

int mst_stft_f32(const float* d_audio, const mst_batch_t* b, int out_mode, int layout, void* d_out,
                 mst_stream_t stream) {
  if (!d_audio || !b || !d_out) return fail(MST_ERR_INVALID, "mst_stft_f32: null argument");
  if (layout != MST_LAYOUT_FRAME_MAJOR && layout != MST_LAYOUT_BIN_MAJOR) return fail(MST_ERR_INVALID, "bad layout %d", layout);
  MelDev none{};
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (out_mode) {
    case MST_OUT_COMPLEX:
      if (layout != MST_LAYOUT_FRAME_MAJOR)
        return fail(MST_ERR_UNSUPPORTED, "complex STFT output is frame-major only (librosa's native Fortran order)");
      return launch_stft<MST_OUT_COMPLEX>(d_audio, b, layout, d_out, none, s);
    case MST_OUT_MAGNITUDE: return launch_stft<MST_OUT_MAGNITUDE>(d_audio, b, layout, d_out, none, s);
    case MST_OUT_POWER: return launch_stft<MST_OUT_POWER>(d_audio, b, layout, d_out, none, s);
    case MST_OUT_LOG1P_POWER: return launch_stft<MST_OUT_LOG1P_POWER>(d_audio, b, layout, d_out, none, s);
    default: return fail(MST_ERR_INVALID, "bad out_mode %d", out_mode);
  }
}

int mst_mel_plan_create(const float* W, int n_mels, int n_bins, mst_mel_plan_t** out) {
  if (!W || !out) return fail(MST_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n_bins != kBins) return fail(MST_ERR_UNSUPPORTED, "mel plan needs n_bins=1025 (n_fft=2048), got %d", n_bins);
  if (n_mels < 1 || n_mels > 256) return fail(MST_ERR_UNSUPPORTED, "n_mels=%d outside [1,256]", n_mels);
  std::vector<int32_t> meta((size_t)3 * n_mels);
  std::vector<float> compact;
  for (int m = 0; m < n_mels; ++m) {
    int lo = n_bins, hi = -1;
    for (int k = 0; k < n_bins; ++k)
      if (W[(size_t)m * n_bins + k] != 0.0f) { lo = std::min(lo, k); hi = std::max(hi, k); }
    const int cnt = hi >= lo ? hi - lo + 1 : 0;
    meta[m] = cnt ? lo : 0;
    meta[n_mels + m] = cnt;
    meta[2 * n_mels + m] = (int32_t)compact.size();
    for (int k = 0; k < cnt; ++k) compact.push_back(W[(size_t)m * n_bins + lo + k]);
  }
  mst_mel_plan* p = new mst_mel_plan();
  p->n_mels = n_mels; p->n_bins = n_bins;
  if (cudaMalloc(&p->d_w, sizeof(float) * std::max<size_t>(1, compact.size())) != cudaSuccess ||
      cudaMalloc(&p->d_meta, sizeof(int32_t) * meta.size()) != cudaSuccess ||
      cudaMalloc(&p->d_dense, sizeof(float) * (size_t)n_mels * n_bins) != cudaSuccess ||
      cudaMemcpy(p->d_w, compact.data(), sizeof(float) * compact.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_meta, meta.data(), sizeof(int32_t) * meta.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_dense, W, sizeof(float) * (size_t)n_mels * n_bins, cudaMemcpyHostToDevice) != cudaSuccess) {
    mst_mel_plan_destroy(p);
    return fail(MST_ERR_CUDA, "mel plan upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = p;
  return MST_OK;
}

void mst_mel_plan_destroy(mst_mel_plan_t* p) {
  if (!p) return;
  if (p->d_w) cudaFree(p->d_w);
  if (p->d_meta) cudaFree(p->d_meta);
  if (p->d_dense) cudaFree(p->d_dense);
  delete p;
}

size_t mst_stft_mel_workspace_bytes(const mst_batch_t*, const mst_mel_plan_t*) { return 0; }

int mst_stft_mel_f32(const float* d_audio, const mst_batch_t* b, const mst_mel_plan_t* plan, int apply_log1p, int layout,
                     float* d_out, void*, size_t, mst_stream_t stream) {
  if (!d_audio || !b || !plan || !d_out) return fail(MST_ERR_INVALID, "mst_stft_mel_f32: null argument");
  if (layout != MST_LAYOUT_FRAME_MAJOR && layout != MST_LAYOUT_BIN_MAJOR) return fail(MST_ERR_INVALID, "bad layout %d", layout);
  MelDev mel;
  mel.w = plan->d_w;
  mel.klo = plan->d_meta;
  mel.kcnt = plan->d_meta + plan->n_mels;
  mel.woff = plan->d_meta + 2 * plan->n_mels;
  mel.n_mels = plan->n_mels;
  mel.apply_log1p = apply_log1p;
  return launch_stft<kModeMel>(d_audio, b, layout, d_out, mel, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
