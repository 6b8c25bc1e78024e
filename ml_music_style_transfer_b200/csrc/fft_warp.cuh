// Warp-level 1024-point complex FFT (and the 2048-point real transforms built on it).
//
// One warp owns one frame.  N = 1024 = 32 x 32 (Cooley-Tukey): lane l, register r hold element
// 32*r + l.  Pass 1 is a radix-2 DIF FFT-32 entirely in registers (compile-time twiddles), then the
// W_1024^(k1*n2) twiddles, a 32x32 transpose through a padded per-warp shared-memory tile, and pass 2
// (another in-register FFT-32).  Output element k = l + 32*r sits in register slot BR5(r) of lane l,
// i.e. the output layout equals the input layout, so forward and inverse chain without reshuffles.
// The real-FFT split / merge butterflies exchange bin k with bin 1024-k through the same scratch tile.
#pragma once
#include <cuda_runtime.h>

namespace mst {

#define MST_FULL_MASK 0xffffffffu

__host__ __device__ __forceinline__ constexpr int br5(int k) {
  return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// cos / sin of 2*pi*q/32 for q in [0,16)
__device__ __forceinline__ constexpr float cos32(int q) {
  return q == 0   ? 1.0f
         : q == 1 ? 0.98078528040323044913f
         : q == 2 ? 0.92387953251128675613f
         : q == 3 ? 0.83146961230254523708f
         : q == 4 ? 0.70710678118654752440f
         : q == 5 ? 0.55557023301960222474f
         : q == 6 ? 0.38268343236508977173f
         : q == 7 ? 0.19509032201612826785f
         : q == 8 ? 0.0f
                  : -cos32(16 - q);
}
__device__ __forceinline__ constexpr float sin32(int q) { return q <= 8 ? cos32(8 - q) : cos32(q - 8); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// In-register radix-2 DIF FFT-32.  SIGN = -1 forward, +1 inverse (unnormalised).
// Output X[k] is left in v[br5(k)].
template <int SIGN>
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
#pragma unroll
    for (int g = 0; g < 32; g += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const int q = j * (16 / half);
        const float2 a = v[g + j], b = v[g + j + half];
        v[g + j] = make_float2(a.x + b.x, a.y + b.y);
        const float dx = a.x - b.x, dy = a.y - b.y;
        if (q == 0) {
          v[g + j + half] = make_float2(dx, dy);
        } else if (q == 8) {  // multiply by SIGN * i
          v[g + j + half] = SIGN < 0 ? make_float2(dy, -dx) : make_float2(-dy, dx);
        } else {
          const float wr = cos32(q), wi = SIGN * sin32(q);
          v[g + j + half] = make_float2(fmaf(dx, wr, -dy * wi), fmaf(dx, wi, dy * wr));
        }
      }
    }
  }
}

// 1024-point complex FFT across one warp.  `scratch` is this warp's 32x33 float2 tile,
// `tw` the shared-memory table exp(-2*pi*i*k1*n2/1024) laid out [k1][n2].
template <int SIGN>
__device__ __forceinline__ void fft1024_warp(float2 (&v)[32], float2* scratch, const float2* tw, int lane) {
  fft32<SIGN>(v);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    float2 a = v[br5(k1)];
    if (k1 != 0) {
      float2 w = tw[k1 * 32 + lane];
      if (SIGN > 0) w.y = -w.y;
      a = cmul(a, w);
    }
    scratch[k1 * 33 + lane] = a;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = scratch[lane * 33 + j];
  __syncwarp();
  fft32<SIGN>(v);
}

// Forward real FFT of a 2048-sample frame held as z[m] = x[2m] + i*x[2m+1], m = 32*r + lane in v[r].
// On return X[k], k = lane + 32*r, is in x[r] (natural register order); *nyq = X[1024].x (valid on lane 0).
// The split butterfly pairs bin k with bin 1024-k; the pair is exchanged through the warp's scratch tile
// (linear [k] layout, mirrored read) so that no second register copy of the spectrum stays live.
__device__ __forceinline__ void rfft2048_warp(float2 (&v)[32], float2 (&x)[32], float* nyq, float2* scratch,
                                              const float2* tw1024, const float2* tw2048, int lane) {
  fft1024_warp<-1>(v, scratch, tw1024, lane);
#pragma unroll
  for (int r = 0; r < 32; ++r) scratch[r * 32 + lane] = v[br5(r)];
  __syncwarp();
  *nyq = v[0].x - v[0].y;  // Z[0] lives in register slot br5(0) = 0 of lane 0
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const float2 z = v[br5(r)];
    const float2 p = scratch[(1024 - (lane + 32 * r)) & 1023];
    const float2 w = tw2048[lane + 32 * r];
    const float er = 0.5f * (z.x + p.x), ei = 0.5f * (z.y - p.y);
    const float orr = 0.5f * (z.y + p.y), oi = -0.5f * (z.x - p.x);
    x[r] = make_float2(er + fmaf(w.x, orr, -w.y * oi), ei + fmaf(w.x, oi, w.y * orr));
  }
  __syncwarp();
}

// Inverse real FFT: Y[k], k = lane + 32*r in y[r] (natural order) plus Y[1024].x in `nyq` (lane 0).
// Returns z[m] = x[2m] + i*x[2m+1], m = lane + 32*r, UNSCALED by 1/1024, in v[br5(r)].
// Imaginary parts of the DC and Nyquist bins are ignored, as pocketfft's c2r does.
__device__ __forceinline__ void irfft2048_warp(float2 (&y)[32], float nyq, float2 (&v)[32], float2* scratch,
                                               const float2* tw1024, const float2* tw2048, int lane) {
#pragma unroll
  for (int r = 0; r < 32; ++r) scratch[r * 32 + lane] = y[r];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const float2 a = y[r];
    const float2 p = scratch[(1024 - (lane + 32 * r)) & 1023];
    const float2 w = tw2048[lane + 32 * r];  // exp(-i*theta); the merge needs exp(+i*theta) = conj(w)
    const float er = 0.5f * (a.x + p.x), ei = 0.5f * (a.y - p.y);
    const float dr = 0.5f * (a.x - p.x), di = 0.5f * (a.y + p.y);
    const float opr = fmaf(dr, w.x, di * w.y), opi = fmaf(di, w.x, -dr * w.y);  // (dr + i*di) * conj(w)
    v[r] = make_float2(er - opi, ei + opr);
  }
  if (lane == 0) v[0] = make_float2(0.5f * (y[0].x + nyq), 0.5f * (y[0].x - nyq));
  __syncwarp();
  fft1024_warp<+1>(v, scratch, tw1024, lane);
}

// Copy the 24 KB of constant tables into shared memory (all threads of the CTA participate).
__device__ __forceinline__ void stage_tables(float2* s_tw1024, float2* s_tw2048, float* s_window,
                                             const float2* g_tw1024, const float2* g_tw2048,
                                             const float* g_window) {
  const float4* a = reinterpret_cast<const float4*>(g_tw1024);
  const float4* b = reinterpret_cast<const float4*>(g_tw2048);
  const float4* c = reinterpret_cast<const float4*>(g_window);
  float4* sa = reinterpret_cast<float4*>(s_tw1024);
  float4* sb = reinterpret_cast<float4*>(s_tw2048);
  float4* sc = reinterpret_cast<float4*>(s_window);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    sa[i] = __ldg(a + i);
    sb[i] = __ldg(b + i);
    sc[i] = __ldg(c + i);
  }
}

}  // namespace mst
