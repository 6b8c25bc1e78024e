// Shared host/device helpers for the mst_b200 kernels (error plumbing, device tables, batch layout).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/mst_b200.h"

namespace mst {

constexpr int kNfft = 2048;           // preprocess.py:25 / inference.py:105 -- the only n_fft the reference uses
constexpr int kBins = kNfft / 2 + 1;  // 1025
constexpr int kHalf = kNfft / 2;      // complex FFT length (even/odd packing) and centre padding
// Warps (= frames of one tile) per CTA.  8 warps x 2 CTAs = 16 resident warps per SM at 128 registers per thread.  Measured
// alternative (round 2): 10 warps x 2 CTAs = 20 warps at 96 registers (36-88 B of spills, 221 KB of shared memory so the L1
// shrinks to its minimum): Griffin-Lim iteration 19.7 ms vs 18.75 ms, log-mel 14.4 vs 13.7 ms -- slower, so 8 it stays.
#ifndef MST_WARPS_PER_CTA
#define MST_WARPS_PER_CTA 8
#endif
constexpr int kWarpsPerCta = MST_WARPS_PER_CTA;
constexpr int kCtasPerSm = kWarpsPerCta <= 10 ? 2 : 1;  // resident CTAs per SM the FFT kernels are sized for
constexpr int kScratchPerWarp = 32 * 33;  // float2 elements: padded 32x32 transpose tile == one 2048-sample frame slot

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
#define MST_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::mst::fail(MST_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                 \
  } while (0)

// ---- constant tables (per device, built once on the host in double precision) ----------------
struct Tables {
  const float2* tw1024;  // [32][32]: exp(-2*pi*i*k1*n2/1024)
  const float2* twp;     // [520]:    -0.5i * exp(-2*pi*i*k/2048), k = 0..512 (real-FFT split twiddles, padded)
  const float* window;   // [2048]:   periodic Hann (scipy get_window('hann', 2048, fftbins=True))
  const float* wsyn;     // [2048]:   window / 1024 (synthesis window with the inverse-FFT scale folded in)
};
int get_tables(Tables* out);  // for the current device
constexpr int kTwpCount = 520;

// ---- batch descriptor -------------------------------------------------------------------------
struct ClipDesc {        // one per clip, device resident
  int64_t sample_offset; // first sample of the clip in the audio buffer
  int64_t length;        // samples
  int64_t frame_offset;  // prefix sum of frames
  int64_t acc_offset;    // Griffin-Lim overlap-add accumulator offset (floats)
  int32_t frames;        // T = 1 + length / hop
  int32_t tile_offset;   // prefix sum of ceil(T / kWarpsPerCta)
};

// ---- general path for n_fft != 2048 (generic_fft.cu) ---------------------------------------------------------------------
bool generic_n_fft_ok(int n_fft);  // power of two in [64, 16384]
}  // namespace mst
struct mst_batch;
namespace mst {
int generic_stft(const float* d_audio, const mst_batch* b, int out_mode, int layout, void* d_out, cudaStream_t s);
int generic_spectral_convergence(const float* d_y, const mst_batch* b, const float* d_S, int s_layout, double* d_num,
                                 double* d_den, cudaStream_t s);
size_t generic_gl_workspace_bytes(const mst_batch* b);
int generic_griffinlim(const float* d_S, int s_layout, int s_is_log1p_power, const mst_batch* b, int n_iter, float momentum,
                       const float* d_init_phase, int init_mode, uint64_t seed, float* d_y_out, void* d_workspace,
                       cudaStream_t s);

}  // namespace mst

struct mst_batch {
  int n_clips = 0;
  int n_fft = 0, hop = 0, pad_mode = 0;
  int win_length = 0;                // <= n_fft; the periodic Hann window of this length is centre-padded to n_fft (librosa)
  float* d_window = nullptr;         // [n_fft] padded analysis window, NULL = the default table (n_fft == win_length == 2048)
  float* d_wsyn = nullptr;           // [n_fft] padded window / (n_fft / 2): synthesis window with the inverse-FFT scale
  int64_t total_frames = 0, total_samples = 0, total_acc = 0;
  int64_t audio_extent = 0;          // max over clips of sample_offset + length: elements the audio buffer must hold
  int total_tiles = 0;
  int uniform_frames = 0;            // frames per clip when every clip has the same length, else 0
  int device = 0;
  bool from_frames = false;
  mst::ClipDesc* h_clips = nullptr;  // host copy
  mst::ClipDesc* d_clips = nullptr;  // device copy
  int32_t* d_tile_clip = nullptr;    // [total_tiles] clip id of each frame tile
  float* d_wq = nullptr;             // Griffin-Lim: analysis window * interior (hop-periodic) 1/window-sum-square
  float* d_inv_wss = nullptr;        // Griffin-Lim: 1 / window-sum-square per accumulator position (shared by equal-T clips)
  int64_t* d_wss_offset = nullptr;   // [n_clips] offset of the clip's envelope inside d_inv_wss
};
