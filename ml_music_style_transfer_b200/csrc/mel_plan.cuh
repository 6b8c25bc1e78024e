// Mel plan: device-resident forms of one filterbank, shared by the STFT kernel (split-precision power-spectrum
// writer) and the tcgen05 projection kernel.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "mst_common.cuh"

namespace mst {

constexpr int kSpecPad = 1088;                 // 1025 bins padded to 17 K-slices of 64
constexpr int kMaxSlices = kSpecPad / 64;      // 17

struct MelSlices {      // per 64-bin K-slice: the band of mel rows that is non-zero there
  int n_slices;
  int n0[kMaxSlices];   // first mel row (multiple of 16) == TMEM column offset of the slice's accumulator window
  int n[kMaxSlices];    // rows in the band (multiple of 16; 0 = slice skipped)
  int row[kMaxSlices];  // first row of the slice inside the packed band arrays
};

}  // namespace mst

struct mst_mel_plan {
  int n_mels = 0, n_bins = 0;
  mst::MelSlices slices{};
  int w_rows = 0;                      // sum of slices.n
  __nv_bfloat16* d_band_hi = nullptr;  // [w_rows][64] bf16(W)
  __nv_bfloat16* d_band_lo = nullptr;  // [w_rows][64] bf16(W - bf16(W))
  // n_bins != 1025 (n_fft != 2048): dense float32 filterbank + the non-zero band of every row, for generic_fft.cu
  float* d_dense_w = nullptr;          // [n_mels][n_bins]
  int32_t* d_k_lo = nullptr;           // [n_mels]
  int32_t* d_k_hi = nullptr;           // [n_mels]
};

namespace mst {
size_t mel_gemm_smem_bytes(const mst_mel_plan* plan);
int launch_mel_gemm(const mst_mel_plan* plan, const mst_batch* b, void* ring_hi, void* ring_lo, int64_t ring_rows,
                    int n_rows, int64_t g0, int apply_log1p, int layout, float* out, cudaStream_t stream);
int generic_stft_mel(const float* d_audio, const mst_batch* b, const mst_mel_plan* plan, int apply_log1p, int layout,
                     float* d_out, cudaStream_t s);
}  // namespace mst
