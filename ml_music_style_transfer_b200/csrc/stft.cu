// P1 / P2: framing + Hann window + 2048-point real FFT with the epilogue fused in: complex spectrum, magnitude, power,
// log1p(power), or -- for the mel path -- the power spectrum as split-precision bf16 (hi, lo) rows handed to the
// tcgen05 projection kernel in mel_gemm.cu through an L2-resident ring.  One warp per frame, persistent CTAs of
// kWarpsPerCta warps walking tiles of kWarpsPerCta consecutive frames of one clip.
//
// Replaces librosa.stft + np.log1p(np.abs(.)**2) (reference preprocessing/preprocess.py:47-57) and, together with
// mel_gemm.cu, librosa.feature.melspectrogram (reference tests/plot_spec.py:20).
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <vector>
#include "fft_warp.cuh"
#include "mel_plan.cuh"
#include "mst_common.cuh"

namespace mst {

constexpr int kModeConv = 5;          // internal epilogue: per-clip sums of (|S| - S_target)^2 and S_target^2 (spectral convergence)
constexpr int kModeSplit = 4;         // internal epilogue: |S|^2 as split bf16 (hi, lo) rows for the tensor-core mel projection
// Bin-major staging: every warp parks its frame's 1025 epilogue values inside ITS OWN scratch tile (so no barrier is
// needed between the FFT and the parking), row f starting 4 f floats into the tile: with a row pitch of 2112 floats
// (== 0 mod 32) the shift makes bank = (4 f + k) % 32 distinct over the 8 frames x 4 bins one warp reads per trip.
constexpr int kTileStride = 2 * kScratchPerWarp + 4;   // floats from row f to row f + 1 (scratch pitch + the 4-float shift)
// (the bin-major epilogue below maps 32 lanes to 8 frames x 4 bins: builds with another tile size refuse that layout)
static_assert(kBins + 4 * (kWarpsPerCta - 1) <= 2 * kScratchPerWarp, "staging row must fit the warp's scratch tile");

struct SplitOut {          // ring of split-precision power-spectrum rows (kSpecPad bf16 each), consumed by mel_gemm.cu
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  int64_t g0;              // global frame id stored in ring row 0
  // kModeConv only: target magnitudes (layout `layout`) and per-clip accumulators
  const float* target;
  double* num;             // [n_clips] sum (|STFT(y)| - S)^2
  double* den;             // [n_clips] sum S^2
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
    v[r] = pk_mul(v[r], w2[32 * r + lane]);
  }
}

template <int MODE>
__device__ __forceinline__ float epilogue_value(float2 x) {
  const float p = fmaf(x.x, x.x, x.y * x.y);
  if (MODE == MST_OUT_MAGNITUDE) return sqrtf(p);
  if (MODE == MST_OUT_LOG1P_POWER) return fast_log1p(p);
  return p;
}

// Epilogue of two bins at once (packed arithmetic where the operation allows it).
template <int MODE>
__device__ __forceinline__ float2 epilogue_pair(float2 a, float2 b) {
  const float2 p = make_float2(fmaf(a.x, a.x, a.y * a.y), fmaf(b.x, b.x, b.y * b.y));
  if (MODE == MST_OUT_MAGNITUDE) return make_float2(sqrtf(p.x), sqrtf(p.y));
  if (MODE == MST_OUT_LOG1P_POWER) return fast_log1p2(p);
  return p;
}

// GROUPS = tiles (groups of 8 warps) one CTA works on at a time.  The frame-major / complex / split / convergence paths
// have no cross-warp communication, so they run as ONE CTA of 16 warps per SM (GROUPS = 2): the tables are staged once
// per SM instead of twice, 152 KB instead of 176 KB of shared memory leave the L1 twice the room for the 75 %-overlapping
// frame loads, and half as many CTAs are launched per (short) mel chunk -- same-box A/B: log-mel 6.71 -> 6.32 ms,
// frame-major log1p-power 5.10 -> 4.75 ms per 8 192 clips.  The bin-major epilogue synchronises the 8 warps of a tile
// with CTA barriers and keeps GROUPS = 1 (2 CTAs x 8 warps per SM), like the Griffin-Lim kernel, which a 16-warp barrier
// slows down by 19 %.
constexpr size_t stft_smem_bytes(int groups) {
  return 8192 + sizeof(float2) * kTwpCount + 8192 + sizeof(float2) * kScratchPerWarp * kWarpsPerCta * groups;
}

template <int MODE, int GROUPS>
__global__ void __launch_bounds__(GROUPS * kWarpsPerCta * 32, GROUPS == 1 ? kCtasPerSm : 1)
stft_kernel(const float* __restrict__ audio, const ClipDesc* __restrict__ clips, const int32_t* __restrict__ tile_clip,
            int tile_begin, int tile_end, int hop, int pad_mode, Tables tabs, int layout, void* __restrict__ out_v,
            SplitOut split) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tw1024 = reinterpret_cast<float2*>(smem_raw);
  float2* s_twp = s_tw1024 + 1024;
  float* s_window = reinterpret_cast<float*>(s_twp + kTwpCount);
  float2* s_scratch_all = reinterpret_cast<float2*>(s_window + kNfft);
  const int warp_all = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp_all / kWarpsPerCta, warp = warp_all % kWarpsPerCta;   // tile group of this warp, frame inside the tile
  const int kb = mirror_base(lane);
  float2* scratch = s_scratch_all + warp_all * kScratchPerWarp;
  float* s_tile = reinterpret_cast<float*>(s_scratch_all);  // bin-major staging tile aliases the scratch region (GROUPS == 1)

  stage_table(s_tw1024, tabs.tw1024, 512);
  stage_table(s_twp, tabs.twp, kTwpCount / 2);
  stage_table(s_window, tabs.window, 512);
  __syncthreads();

  const int n_out = kBins;  // values per frame in the output

  const int tile_step = gridDim.x * GROUPS;
  for (int tile = tile_begin + blockIdx.x * GROUPS + grp; tile < tile_end; tile += tile_step) {
    const int c = __ldg(tile_clip + tile);
    const ClipDesc cd = clips[c];
    const int t0 = (tile - cd.tile_offset) * kWarpsPerCta;
    const int t = t0 + warp;
    const bool active = t < cd.frames;
    {  // pull the audio of this CTA's next tile towards L2 while this one computes
      const int nt = tile + tile_step;
      if (nt < tile_end && warp == 0) {
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

    if (MODE == kModeConv) {
      // spectral convergence: || |STFT(y)| - S ||_F^2 and || S ||_F^2 per clip (model/inference.py:149-150, normalised)
      double num = 0.0, den = 0.0;
      if (active) {
#pragma unroll
        for (int j = 0; j < 33; ++j) {
          if (j == 32 && lane != 0) break;
          const int k = j < 32 ? mirror_bin(lane, kb, j) : 512;
          const float2 xv = j < 32 ? o[j] : mid;
          const float mag = sqrtf(fmaf(xv.x, xv.x, xv.y * xv.y));
          const int64_t idx = layout == MST_LAYOUT_FRAME_MAJOR ? g * kBins + k
                                                                : cd.frame_offset * kBins + (int64_t)k * cd.frames + t;
          const float sref = __ldg(split.target + idx);
          const float d = mag - sref;
          num += (double)d * (double)d;
          den += (double)sref * (double)sref;
        }
      }
      for (int off = 16; off; off >>= 1) {
        num += __shfl_xor_sync(MST_FULL_MASK, num, off);
        den += __shfl_xor_sync(MST_FULL_MASK, den, off);
      }
      if (active && lane == 0) {
        atomicAdd(split.num + c, num);
        atomicAdd(split.den + c, den);
      }
      continue;
    }

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
    if (MODE == kModeSplit) {
      // |S|^2 -> (hi, lo) bf16 rows staged in this warp's scratch, then copied out with 128-bit stores
      __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(scratch);
      __nv_bfloat16* sl = sh + kSpecPad;
      __syncwarp();
      if (active) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float p = fmaf(o[j].x, o[j].x, o[j].y * o[j].y);
          const __nv_bfloat16 h = __float2bfloat16_rn(p);
          const int k = mirror_bin(lane, kb, j);
          sh[k] = h;
          sl[k] = __float2bfloat16_rn(p - __bfloat162float(h));
        }
        if (lane == 0) {
          const float p = fmaf(mid.x, mid.x, mid.y * mid.y);
          const __nv_bfloat16 h = __float2bfloat16_rn(p);
          sh[512] = h;
          sl[512] = __float2bfloat16_rn(p - __bfloat162float(h));
        }
        for (int k = kBins + lane; k < kSpecPad; k += 32) {  // zero padding of the last K-slice
          sh[k] = __float2bfloat16_rn(0.0f);
          sl[k] = __float2bfloat16_rn(0.0f);
        }
      }
      __syncwarp();
      if (active) {
        const int64_t r = g - split.g0;
        uint4* dh = reinterpret_cast<uint4*>(split.hi + r * kSpecPad);
        uint4* dl = reinterpret_cast<uint4*>(split.lo + r * kSpecPad);
        const uint4* s4h = reinterpret_cast<const uint4*>(sh);
        const uint4* s4l = reinterpret_cast<const uint4*>(sl);
        for (int i = lane; i < kSpecPad / 8; i += 32) {
          dh[i] = s4h[i];
          dl[i] = s4l[i];
        }
      }
      __syncwarp();
      continue;
    } else {
      if (layout == MST_LAYOUT_FRAME_MAJOR) {
        if (active) {
          float* row = out + g * kBins;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 e = epilogue_pair<MODE>(o[j], o[j + 1]);
            row[mirror_bin(lane, kb, j)] = e.x;
            row[mirror_bin(lane, kb, j + 1)] = e.y;
          }
          if (lane == 0) row[512] = epilogue_value<MODE>(mid);
        }
      } else {
        __syncwarp();  // this warp is done with its scratch tile (split butterflies read registers only)
        if (active) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 e = epilogue_pair<MODE>(o[j], o[j + 1]);
            s_tile[warp * kTileStride + mirror_bin(lane, kb, j)] = e.x;
            s_tile[warp * kTileStride + mirror_bin(lane, kb, j + 1)] = e.y;
          }
          if (lane == 0) s_tile[warp * kTileStride + 512] = epilogue_value<MODE>(mid);
        }
      }
    }
    if (layout == MST_LAYOUT_BIN_MAJOR) {
      // transposed store: clip block is [n_out][T]; the tile's 8 consecutive frames give one 32-byte segment per bin row
      __syncthreads();
      if (kWarpsPerCta != 8 || GROUPS != 1) __trap();   // the host never launches this layout with another geometry
      const int nvalid = min(kWarpsPerCta, cd.frames - t0);
      float* blk = out + cd.frame_offset * n_out;
      const bool vec4 = (cd.frames & 3) == 0 && ((cd.frame_offset * n_out) & 3) == 0 && nvalid == kWarpsPerCta &&
                        (reinterpret_cast<uintptr_t>(out) & 15) == 0;
      if (vec4) {
        // T % 4 == 0 (e.g. the reference's 860-frame chunks): every row segment is two aligned float4 stores
        const int h = threadIdx.x & 1;
        for (int k = threadIdx.x >> 1; k < n_out; k += kWarpsPerCta * 16) {
          const float* sp = s_tile + (4 * h) * kTileStride + k;
          const float4 v = make_float4(sp[0], sp[kTileStride], sp[2 * kTileStride], sp[3 * kTileStride]);
          *reinterpret_cast<float4*>(blk + (int64_t)k * cd.frames + t0 + 4 * h) = v;
        }
      } else {
        const int f = threadIdx.x & 7, kk = threadIdx.x >> 3;
        if (f < nvalid) {
          for (int k = kk; k < n_out; k += 32) blk[(int64_t)k * cd.frames + t0 + f] = s_tile[f * kTileStride + k];
        }
      }
      __syncthreads();
    }
  }
}

template <int MODE, int GROUPS>
static int launch_stft_g(const float* d_audio, const mst_batch* b, int layout, void* d_out, const SplitOut& split,
                         int tile_begin, int tile_end, const Tables& tabs, cudaStream_t stream) {
  const size_t smem = stft_smem_bytes(GROUPS);
  static std::atomic<bool> attr_set[64];  // one flag per device and per (MODE, GROUPS) instantiation
  int dev = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(MST_ERR_INVALID, "device index %d out of range", dev);
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    MST_CUDA_OK(cudaFuncSetAttribute(stft_kernel<MODE, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev].store(true, std::memory_order_release);
  }
  int sms = 0;
  MST_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n_tiles = tile_end - tile_begin;
  const int grid = std::min((n_tiles + GROUPS - 1) / GROUPS, (GROUPS == 1 ? kCtasPerSm : 1) * sms);
  stft_kernel<MODE, GROUPS><<<grid, GROUPS * kWarpsPerCta * 32, smem, stream>>>(
      d_audio, b->d_clips, b->d_tile_clip, tile_begin, tile_end, b->hop, b->pad_mode, tabs, layout, d_out, split);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

template <int MODE>
static int launch_stft(const float* d_audio, const mst_batch* b, int layout, void* d_out, const SplitOut& split,
                       int tile_begin, int tile_end, cudaStream_t stream) {
  Tables tabs;
  int rc = get_tables(&tabs);
  if (rc) return rc;
  if (b->d_window) tabs.window = b->d_window;  // win_length < n_fft: this batch's centre-padded window
  // real-valued epilogues in bin-major order synchronise a tile's 8 warps with CTA barriers: 2 CTAs x 8 warps per SM;
  // everything else: 1 CTA x 16 warps per SM
  const bool barriers = layout == MST_LAYOUT_BIN_MAJOR &&
                        (MODE == MST_OUT_MAGNITUDE || MODE == MST_OUT_POWER || MODE == MST_OUT_LOG1P_POWER);
#ifdef MST_STFT_FORCE_G1   // A/B switch (tools/build_variant.sh): 2 CTAs x 8 warps for every layout
  const bool force_g1 = true;
#else
  const bool force_g1 = false;
#endif
  if (barriers || force_g1 || kWarpsPerCta != 8)
    return launch_stft_g<MODE, 1>(d_audio, b, layout, d_out, split, tile_begin, tile_end, tabs, stream);
  return launch_stft_g<MODE, 2>(d_audio, b, layout, d_out, split, tile_begin, tile_end, tabs, stream);
}

}  // namespace mst

using namespace mst;

extern "C" {

int mst_stft_f32(const float* d_audio, const mst_batch_t* b, int out_mode, int layout, void* d_out,
                 mst_stream_t stream) {
  if (!d_audio || !b || !d_out) return fail(MST_ERR_INVALID, "mst_stft_f32: null argument");
  if (layout != MST_LAYOUT_FRAME_MAJOR && layout != MST_LAYOUT_BIN_MAJOR) return fail(MST_ERR_INVALID, "bad layout %d", layout);
  SplitOut none{};
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (b->n_fft != kNfft) return generic_stft(d_audio, b, out_mode, layout, d_out, s);  // general path (generic_fft.cu)
  const int nt = b->total_tiles;
  switch (out_mode) {
    case MST_OUT_COMPLEX:
      if (layout != MST_LAYOUT_FRAME_MAJOR)
        return fail(MST_ERR_UNSUPPORTED, "complex STFT output is frame-major only (librosa's native Fortran order)");
      return launch_stft<MST_OUT_COMPLEX>(d_audio, b, layout, d_out, none, 0, nt, s);
    case MST_OUT_MAGNITUDE: return launch_stft<MST_OUT_MAGNITUDE>(d_audio, b, layout, d_out, none, 0, nt, s);
    case MST_OUT_POWER: return launch_stft<MST_OUT_POWER>(d_audio, b, layout, d_out, none, 0, nt, s);
    case MST_OUT_LOG1P_POWER: return launch_stft<MST_OUT_LOG1P_POWER>(d_audio, b, layout, d_out, none, 0, nt, s);
    default: return fail(MST_ERR_INVALID, "bad out_mode %d", out_mode);
  }
}

int mst_spectral_convergence_f32(const float* d_y, const mst_batch_t* b, const float* d_S, int s_layout, double* d_num,
                                 double* d_den, mst_stream_t stream) {
  if (!d_y || !b || !d_S || !d_num || !d_den) return fail(MST_ERR_INVALID, "mst_spectral_convergence_f32: null argument");
  if (s_layout != MST_LAYOUT_FRAME_MAJOR && s_layout != MST_LAYOUT_BIN_MAJOR) return fail(MST_ERR_INVALID, "bad layout %d", s_layout);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MST_CUDA_OK(cudaMemsetAsync(d_num, 0, sizeof(double) * (size_t)b->n_clips, s));
  MST_CUDA_OK(cudaMemsetAsync(d_den, 0, sizeof(double) * (size_t)b->n_clips, s));
  if (b->n_fft != kNfft) return generic_spectral_convergence(d_y, b, d_S, s_layout, d_num, d_den, s);
  SplitOut so{};
  so.target = d_S;
  so.num = d_num;
  so.den = d_den;
  return launch_stft<kModeConv>(d_y, b, s_layout, nullptr, so, 0, b->total_tiles, s);
}

int mst_mel_plan_create(const float* W, int n_mels, int n_bins, mst_mel_plan_t** out) {
  if (!W || !out) return fail(MST_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n_mels < 1 || n_mels > 256) return fail(MST_ERR_UNSUPPORTED, "n_mels=%d outside [1,256]", n_mels);
  if (n_bins != kBins) {
    // another n_fft: the general path projects with the dense float32 filterbank, row by row over its non-zero band
    if (n_bins < 2 || !generic_n_fft_ok(2 * (n_bins - 1)))
      return fail(MST_ERR_UNSUPPORTED, "mel plan: n_bins=%d is not 1 + n_fft/2 of a supported n_fft", n_bins);
    mst_mel_plan* p = new mst_mel_plan();
    p->n_mels = n_mels; p->n_bins = n_bins;
    std::vector<int32_t> k_lo((size_t)n_mels, 0), k_hi((size_t)n_mels, 0);
    for (int m = 0; m < n_mels; ++m) {
      int lo_k = n_bins, hi_k = 0;
      for (int k = 0; k < n_bins; ++k)
        if (W[(size_t)m * n_bins + k] != 0.0f) { lo_k = std::min(lo_k, k); hi_k = k + 1; }
      k_lo[(size_t)m] = std::min(lo_k, hi_k); k_hi[(size_t)m] = hi_k;
    }
    const size_t wb = sizeof(float) * (size_t)n_mels * (size_t)n_bins, ib = sizeof(int32_t) * (size_t)n_mels;
    if (cudaMalloc(&p->d_dense_w, wb) != cudaSuccess || cudaMalloc(&p->d_k_lo, ib) != cudaSuccess ||
        cudaMalloc(&p->d_k_hi, ib) != cudaSuccess ||
        cudaMemcpy(p->d_dense_w, W, wb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->d_k_lo, k_lo.data(), ib, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->d_k_hi, k_hi.data(), ib, cudaMemcpyHostToDevice) != cudaSuccess) {
      mst_mel_plan_destroy(p);
      return fail(MST_ERR_CUDA, "mel plan upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = p;
    return MST_OK;
  }
  mst_mel_plan* p = new mst_mel_plan();
  p->n_mels = n_mels; p->n_bins = n_bins;
  // band of non-zero mel rows per 64-bin K-slice, padded to multiples of 16 rows (UMMA N granularity at M=128)
  std::vector<__nv_bfloat16> hi, lo;
  p->slices.n_slices = kMaxSlices;
  for (int s = 0; s < kMaxSlices; ++s) {
    int lo_m = n_mels, hi_m = -1;
    for (int m = 0; m < n_mels; ++m)
      for (int k = 64 * s; k < std::min(64 * s + 64, n_bins); ++k)
        if (W[(size_t)m * n_bins + k] != 0.0f) { lo_m = std::min(lo_m, m); hi_m = std::max(hi_m, m); }
    p->slices.row[s] = (int)(hi.size() / 64);
    if (hi_m < 0) { p->slices.n0[s] = 0; p->slices.n[s] = 0; continue; }
    const int n0 = (lo_m / 16) * 16, n1 = (hi_m / 16 + 1) * 16;
    p->slices.n0[s] = n0; p->slices.n[s] = n1 - n0;
    for (int m = n0; m < n1; ++m)
      for (int k = 64 * s; k < 64 * s + 64; ++k) {
        const float w = (m < n_mels && k < n_bins) ? W[(size_t)m * n_bins + k] : 0.0f;
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        hi.push_back(h);
        lo.push_back(__float2bfloat16_rn(w - __bfloat162float(h)));
      }
  }
  p->w_rows = (int)(hi.size() / 64);
  {
    // store the exact shared-memory image the projection kernel wants: K-major SWIZZLE_128B, i.e. the 16-byte
    // chunk c (8 bf16) of 128-byte row r lives at chunk position c ^ (r & 7)
    std::vector<__nv_bfloat16> ih(hi.size()), il(lo.size());
    for (int r = 0; r < p->w_rows; ++r)
      for (int c = 0; c < 8; ++c)
        for (int e = 0; e < 8; ++e) {
          ih[(size_t)r * 64 + ((c ^ (r & 7)) * 8) + e] = hi[(size_t)r * 64 + c * 8 + e];
          il[(size_t)r * 64 + ((c ^ (r & 7)) * 8) + e] = lo[(size_t)r * 64 + c * 8 + e];
        }
    hi.swap(ih);
    lo.swap(il);
  }
  const size_t band_bytes = std::max<size_t>(1, hi.size()) * sizeof(__nv_bfloat16);
  if (cudaMalloc(&p->d_band_hi, band_bytes) != cudaSuccess || cudaMalloc(&p->d_band_lo, band_bytes) != cudaSuccess ||
      cudaMemcpy(p->d_band_hi, hi.data(), hi.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(p->d_band_lo, lo.data(), lo.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice) != cudaSuccess) {
    mst_mel_plan_destroy(p);
    return fail(MST_ERR_CUDA, "mel plan upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (mel_gemm_smem_bytes(p) > 227 * 1024) {
    mst_mel_plan_destroy(p);
    return fail(MST_ERR_UNSUPPORTED, "banded filterbank (%d rows) does not fit in shared memory", p->w_rows);
  }
  *out = p;
  return MST_OK;
}

void mst_mel_plan_destroy(mst_mel_plan_t* p) {
  if (!p) return;
  if (p->d_band_hi) cudaFree(p->d_band_hi);
  if (p->d_band_lo) cudaFree(p->d_band_lo);
  if (p->d_dense_w) cudaFree(p->d_dense_w);
  if (p->d_k_lo) cudaFree(p->d_k_lo);
  if (p->d_k_hi) cudaFree(p->d_k_hi);
  delete p;
}

// Ring of split-precision power-spectrum rows between the STFT kernel and the projection kernel.  Round 1 sized it as ONE
// wave of 128-frame projection tiles (n_SM x 128 rows x 2 x 2176 B ~ 82 MB) so that it stays in the 126 MB L2; measured in
// round 2 (tools/ab_mel.py, same box, 8 192 clips): 1 / 2 / 4 / 8 waves per chunk = 6.28 / 5.68 / 5.39 / 5.27 ms.  With one
// tile per CTA the projection kernel never overlaps a tile's epilogue with the next tile's MMAs (its two TMEM accumulators
// exist for exactly that) and pays its filterbank load and pipeline fill per 128 frames; the extra HBM round trip of a
// ring that no longer fits in L2 (2 x 4.3 KB per frame) is far below what either kernel needs in time.  Default: 8 waves,
// never more rows than the batch has frames.
#ifndef MST_MEL_RING_WAVES
#define MST_MEL_RING_WAVES 8
#endif
static int64_t ring_rows_for_device() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static const int waves = [] {   // A/B knob: MST_MEL_RING_WAVES (env) = projection tiles per CTA and ring chunk
    const char* e = getenv("MST_MEL_RING_WAVES");
    const int w = e ? atoi(e) : MST_MEL_RING_WAVES;
    return w >= 1 && w <= 16 ? w : MST_MEL_RING_WAVES;
  }();
  return (int64_t)sms * 128 * waves;
}

static int64_t ring_rows_for_batch(const mst_batch_t* b) {
  const int64_t need = b ? ((b->total_frames + 127) / 128) * 128 : 0;  // whole projection tiles
  return std::max<int64_t>(128, std::min(ring_rows_for_device(), need));
}

size_t mst_stft_mel_workspace_bytes(const mst_batch_t* b, const mst_mel_plan_t*) {
  return 2 * (size_t)ring_rows_for_batch(b) * kSpecPad * sizeof(__nv_bfloat16) + 1024;
}

int mst_stft_mel_f32(const float* d_audio, const mst_batch_t* b, const mst_mel_plan_t* plan, int apply_log1p, int layout,
                     float* d_out, void* d_workspace, size_t workspace_bytes, mst_stream_t stream) {
  if (!d_audio || !b || !plan || !d_out || !d_workspace) return fail(MST_ERR_INVALID, "mst_stft_mel_f32: null argument");
  if (layout != MST_LAYOUT_FRAME_MAJOR && layout != MST_LAYOUT_BIN_MAJOR) return fail(MST_ERR_INVALID, "bad layout %d", layout);
  if (workspace_bytes < mst_stft_mel_workspace_bytes(b, plan))
    return fail(MST_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, mst_stft_mel_workspace_bytes(b, plan));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (b->n_fft != kNfft) return generic_stft_mel(d_audio, b, plan, apply_log1p, layout, d_out, s);
  if (plan->n_bins != kBins) return fail(MST_ERR_INVALID, "mel plan has %d bins, the batch (n_fft=2048) needs 1025", plan->n_bins);
  const int64_t ring_rows = ring_rows_for_batch(b);
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_workspace) + 1023) & ~(uintptr_t)1023);
  SplitOut split;
  split.hi = reinterpret_cast<__nv_bfloat16*>(ws);
  split.lo = split.hi + ring_rows * kSpecPad;
  const int chunk_tiles = (int)(ring_rows / kWarpsPerCta);  // a tile holds at most kWarpsPerCta frames
  int c = 0;                                                // clip cursor (tiles are ordered by clip)
  for (int tile0 = 0; tile0 < b->total_tiles; tile0 += chunk_tiles) {
    const int tile1 = std::min(b->total_tiles, tile0 + chunk_tiles);
    while (c + 1 < b->n_clips && b->h_clips[c + 1].tile_offset <= tile0) ++c;
    const int64_t g0 = b->h_clips[c].frame_offset + (int64_t)(tile0 - b->h_clips[c].tile_offset) * kWarpsPerCta;
    int64_t g1 = b->total_frames;
    if (tile1 < b->total_tiles) {
      int c1 = c;
      while (c1 + 1 < b->n_clips && b->h_clips[c1 + 1].tile_offset <= tile1) ++c1;
      g1 = b->h_clips[c1].frame_offset + (int64_t)(tile1 - b->h_clips[c1].tile_offset) * kWarpsPerCta;
    }
    split.g0 = g0;
    int rc = launch_stft<kModeSplit>(d_audio, b, layout, nullptr, split, tile0, tile1, s);
    if (rc) return rc;
    rc = launch_mel_gemm(plan, b, split.hi, split.lo, ring_rows, (int)(g1 - g0), g0, apply_log1p, layout, d_out, s);
    if (rc) return rc;
  }
  return MST_OK;
}

}  // extern "C"
