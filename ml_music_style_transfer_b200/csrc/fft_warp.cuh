// Warp-level 1024-point complex FFT (and the 2048-point real transforms built on it).
//
// One warp owns one frame.  N = 1024 = 32 x 32 (Cooley-Tukey): lane l, register r hold element
// 32*r + l.  Pass 1 is a radix-2 DIT FFT-32 entirely in registers (compile-time twiddles), then the
// W_1024^(k1*n2) twiddles, a 32x32 transpose through a padded per-warp shared-memory tile, and pass 2
// (another in-register FFT-32).  Output element k = l + 32*r sits in register slot BR5(r) of lane l,
// i.e. the output layout equals the input layout, so forward and inverse chain without reshuffles.
// The real-FFT split / merge butterflies exchange bin k with bin 1024-k by warp shuffle (mirror layout below).
#pragma once
#include <cuda_runtime.h>

namespace mst {

#define MST_FULL_MASK 0xffffffffu

__host__ __device__ __forceinline__ constexpr int br5(int k) {
  return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// cos / sin of 2*pi*q/32 for q in [0,16]: plain constant tables so that, after full unrolling, every twiddle folds
// into an immediate operand (a recursive constexpr helper was NOT folded by nvcc and ended up as real calls).
#define MST_C1 0.98078528040323044913f
#define MST_C2 0.92387953251128675613f
#define MST_C3 0.83146961230254523708f
#define MST_C4 0.70710678118654752440f
#define MST_C5 0.55557023301960222474f
#define MST_C6 0.38268343236508977173f
#define MST_C7 0.19509032201612826785f
template <int Q>
struct Tw32 {
  static_assert(Q >= 0 && Q <= 16, "twiddle index");
  __host__ __device__ static constexpr float cosv() {
    constexpr float t[17] = {1.0f, MST_C1, MST_C2, MST_C3, MST_C4, MST_C5, MST_C6, MST_C7, 0.0f,
                             -MST_C7, -MST_C6, -MST_C5, -MST_C4, -MST_C3, -MST_C2, -MST_C1, -1.0f};
    return t[Q];
  }
  __host__ __device__ static constexpr float sinv() {
    constexpr float t[17] = {0.0f, MST_C7, MST_C6, MST_C5, MST_C4, MST_C3, MST_C2, MST_C1, 1.0f,
                             MST_C1, MST_C2, MST_C3, MST_C4, MST_C5, MST_C6, MST_C7, 0.0f};
    return t[Q];
  }
};

// ---- packed fp32x2 arithmetic (sm_100: PTX fma/add/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2) ------------------------
// One instruction does the same IEEE operation on both halves of a 64-bit register pair, i.e. on (re, im) of a float2.
// ptxas folds half swaps, per-half negation, scalar broadcast and immediates into operand modifiers
// (R4.F32x2.LO_HI.NP, R0.F32, 0.92387...), so a complex radix-2 butterfly is 3 issue slots instead of 6 and every
// result is bit-identical to the scalar formulation.  The FFT kernels are issue-bound, not FMA-pipe-bound.
#ifndef MST_PACKED
#define MST_PACKED 1
#endif
#if MST_PACKED
__device__ __forceinline__ unsigned long long pk_pack(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 pk_unpack(unsigned long long r) {
  float2 a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk_pack(a)), "l"(pk_pack(b)), "l"(pk_pack(c)));
  return pk_unpack(d);
}
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_pack(a)), "l"(pk_pack(b)));
  return pk_unpack(d);
}
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_pack(a)), "l"(pk_pack(b)));
  return pk_unpack(d);
}
#else
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#endif
__device__ __forceinline__ float2 pk_bcast(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 pk_neg(float2 a) { return make_float2(-a.x, -a.y); }

// Shared-memory table reads in groups (MST_LDS_BATCH per group, issued back to back as volatile loads so that ptxas keeps
// them together): with 4 warps per scheduler a 29-cycle LDS latency per table entry cannot be hidden by other warps, and
// ptxas otherwise schedules each load right before its single use (ncu: short_scoreboard 17 % of the stall samples).
#ifndef MST_LDS_BATCH
#define MST_LDS_BATCH 4
#endif
__device__ __forceinline__ float2 lds64(const float2* p) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

// a * b (complex):  (a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x) = a.x * b + a.y * (-b.y, b.x).  The half-swapped,
// half-negated pair must be the FIRST operand of the FFMA2: there ptxas folds swap and sign into operand modifiers
// (-R4.F32x2.LO_HI.NP); as second operand, or in an FMUL2, it is materialised with extra FADD / MOV instructions.
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return pk_fma(make_float2(-b.y, b.x), pk_bcast(a.y), pk_mul(pk_bcast(a.x), b));
}
// a * conj(b):  (a.x*b.x + a.y*b.y, -a.x*b.y + a.y*b.x) = a.x * (b.x, -b.y) + a.y * (b.y, b.x)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
  return pk_fma(make_float2(b.x, -b.y), pk_bcast(a.x), pk_mul(pk_bcast(a.y), make_float2(b.y, b.x)));
}

// One radix-2 decimation-in-time butterfly with the compile-time twiddle w = W_32^Q (forward) / conj (inverse):
//   t = w*b;  a' = a + t;  b' = a - t = 2a - a'.
// Written so that the general case is 6 FMAs (a' by two nested FMAs per component, b' = fma(2, a, -a')) and the
// sqrt(1/2) twiddles are 2 adds + 4 FMAs; multiplications by 1 and -+i are 4 adds.
template <int SIGN, int Q>
__device__ __forceinline__ void bfly(float2& a, float2& b) {
  if (Q == 0) {
    const float2 t = b;
    b = pk_add(a, pk_neg(t));
    a = pk_add(a, t);
  } else if (Q == 8) {  // w = SIGN * i:  t = SIGN * (-b.y, b.x)
    const float2 t = SIGN < 0 ? make_float2(b.y, -b.x) : make_float2(-b.y, b.x);
    b = pk_add(a, pk_neg(t));
    a = pk_add(a, t);
  } else if (Q == 4 || Q == 12) {
    // Q=4:  w = c*(1 + SIGN*i)  -> t = c*(b.x - SIGN*b.y, b.y + SIGN*b.x)
    // Q=12: w = c*(-1 + SIGN*i) -> t = c*(-b.x - SIGN*b.y, -b.y + SIGN*b.x)
    constexpr float c = MST_C4;
    const float2 rot = SIGN < 0 ? make_float2(b.y, -b.x) : make_float2(-b.y, b.x);  // SIGN * i * b
    const float2 u = pk_add(Q == 4 ? b : pk_neg(b), rot);
    const float2 a2 = pk_fma(pk_bcast(c), u, a);
    b = pk_fma(pk_bcast(2.0f), a, pk_neg(a2));
    a = a2;
  } else {
    constexpr float wr = Tw32<Q>::cosv();
    constexpr float wi = SIGN * Tw32<Q>::sinv();
    // a' = a + w*b = a + wi*(-b.y, b.x) + wr*b ;  b' = 2a - a'
    const float2 u = pk_fma(pk_bcast(wi), make_float2(-b.y, b.x), a);
    const float2 a2 = pk_fma(pk_bcast(wr), b, u);
    b = pk_fma(pk_bcast(2.0f), a, pk_neg(a2));
    a = a2;
  }
}

// Stage with butterfly span HALF over the conceptual array u[i] = v[br5(i)] (bit-reversed input == natural v[]).
template <int SIGN, int HALF, int G, int J>
struct StageLoop {
  __device__ __forceinline__ static void run(float2 (&v)[32]) {
    bfly<SIGN, J*(16 / HALF)>(v[br5(G + J)], v[br5(G + J + HALF)]);
    if constexpr (J + 1 < HALF) StageLoop<SIGN, HALF, G, J + 1>::run(v);
    else if constexpr (G + 2 * HALF < 32) StageLoop<SIGN, HALF, G + 2 * HALF, 0>::run(v);
  }
};

// In-register radix-2 DIT FFT-32.  SIGN = -1 forward, +1 inverse (unnormalised).
// Input x[n] in v[n]; output X[k] is left in v[br5(k)].
template <int SIGN>
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
  StageLoop<SIGN, 1, 0, 0>::run(v);
  StageLoop<SIGN, 2, 0, 0>::run(v);
  StageLoop<SIGN, 4, 0, 0>::run(v);
  StageLoop<SIGN, 8, 0, 0>::run(v);
  StageLoop<SIGN, 16, 0, 0>::run(v);
}

// Between the two FFT-32 passes: multiply by the W_1024^(k1*n2) twiddles and park the column in the transpose tile.
template <int SIGN>
__device__ __forceinline__ void twiddle_and_park(float2 (&v)[32], float2* scratch, const float2* tw, int lane) {
#pragma unroll
  for (int g = 0; g < 32; g += MST_LDS_BATCH) {
    float2 w[MST_LDS_BATCH];
#pragma unroll
    for (int i = 0; i < MST_LDS_BATCH; ++i)
      if (g + i != 0) w[i] = MST_LDS_BATCH > 1 ? lds64(tw + (g + i) * 32 + lane) : tw[(g + i) * 32 + lane];
#pragma unroll
    for (int i = 0; i < MST_LDS_BATCH; ++i) {
      const int k1 = g + i;
      float2 a = v[br5(k1)];
      if (k1 != 0) a = SIGN > 0 ? cmul_conj(a, w[i]) : cmul(a, w[i]);
      scratch[k1 * 33 + lane] = a;
    }
  }
}

// 1024-point complex FFT across one warp.  `scratch` is this warp's 32x33 float2 tile,
// `tw` the shared-memory table exp(-2*pi*i*k1*n2/1024) laid out [k1][n2].
// fft1024_front leaves the transposed intermediate in v[] and the scratch tile FREE (callers may start asynchronous
// copies into it); the transform is completed by fft32<SIGN>(v).
template <int SIGN>
__device__ __forceinline__ void fft1024_front(float2 (&v)[32], float2* scratch, const float2* tw, int lane) {
  fft32<SIGN>(v);
  twiddle_and_park<SIGN>(v, scratch, tw, lane);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = scratch[lane * 33 + j];
  __syncwarp();
}

// Both FFT-32 passes share ONE copy of the butterfly code (a 2-trip loop that is deliberately not unrolled): the
// unrolled butterflies are ~600 instructions per instance and the kernels were paying instruction-cache misses.
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};

// `hook` runs right after the transpose, when the scratch tile is free again and the second pass is still to come
// (Griffin-Lim starts its asynchronous copy of the previous iterate there).
template <int SIGN, typename Hook = NoHook>
__device__ __forceinline__ void fft1024_warp(float2 (&v)[32], float2* scratch, const float2* tw, int lane,
                                             Hook hook = Hook()) {
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    fft32<SIGN>(v);
    if (pass == 0) {
      twiddle_and_park<SIGN>(v, scratch, tw, lane);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = scratch[lane * 33 + j];
      __syncwarp();
      hook();
    }
  }
}

// ---- real <-> half-length complex conversion ("mirror layout") ------------------------------------------------
// The split / merge butterflies of a real FFT pair bin k with bin 1024-k.  After the complex FFT lane l holds
// Z[l + 32*r]; bin 1024-k of its bins k = l + 32*r, r < 16, sits in lane (32-l)&31, register 31-r.  One warp shuffle
// per pair moves it over, after which lane l owns BOTH bins of 16 pairs and one complex multiply serves two outputs.
// Resulting register layout of a 1025-bin spectrum ("mirror layout"), j = 0..31:
//     j <  16 : bin l + 32*j
//     j >= 16 : bin (32 - l) + 32*j          (lane 0: bin 32*(j+1), so j = 31 is the Nyquist bin 1024)
//     lane 0 additionally holds bin 512 in a separate register ("mid").
// All elementwise spectrum work (epilogues, Griffin-Lim re-projection) runs in this layout; global accesses stay
// coalesced because consecutive lanes still touch consecutive (ascending or descending) bins.
// kb = mirror_base(lane) is computed once per thread so that every bin index is (lane | kb) + compile-time constant.
__device__ __forceinline__ int mirror_base(int lane) { return lane == 0 ? 32 : 32 - lane; }
__device__ __forceinline__ int mirror_bin(int lane, int kb, int j) { return j < 16 ? lane + 32 * j : kb + 32 * j; }

// One pair butterfly.  z = value at bin k, p = value at bin 1024-k, w = -0.5i*exp(-2*pi*i*k/2048) (CONJ_W: its conjugate,
// the inverse direction).  Returns out_k = 0.5*(z + conj p) + w*(z - conj p) and
// out_m = conj(0.5*(z + conj p) - w*(z - conj p)).
template <bool CONJ_W>
__device__ __forceinline__ void pair_butterfly(float2 z, float2 p, float2 w, float2& out_k, float2& out_m) {
  const float2 s = pk_add(z, make_float2(p.x, -p.y));   // z + conj p
  const float2 d = pk_add(z, make_float2(-p.x, p.y));   // z - conj p
  const float2 t = CONJ_W ? cmul_conj(d, w) : cmul(d, w);   // w * d  (conj(w) * d for the inverse direction)
  out_k = pk_fma(pk_bcast(0.5f), s, t);
  out_m = pk_fma(make_float2(0.5f, -0.5f), s, make_float2(-t.x, t.y));
}

// Forward real FFT of a 2048-sample frame held as z[m] = x[2m] + i*x[2m+1], m = 32*r + lane in v[r].
// Result: spectrum in mirror layout in o[0..31], bin 512 in *mid (lane 0 only).
// twp[k], k = 0..512: -0.5i * exp(-2*pi*i*k/2048).
__device__ __forceinline__ void rfft_split(float2 (&v)[32], float2 (&o)[32], float2* mid, const float2* twp, int lane) {
  const int src = (32 - lane) & 31;
#pragma unroll
  for (int g = 0; g < 16; g += MST_LDS_BATCH) {
    float2 w[MST_LDS_BATCH];
#pragma unroll
    for (int i = 0; i < MST_LDS_BATCH; ++i) w[i] = MST_LDS_BATCH > 1 ? lds64(twp + lane + 32 * (g + i)) : twp[lane + 32 * (g + i)];
#pragma unroll
    for (int i = 0; i < MST_LDS_BATCH; ++i) {
      const int r = g + i;
      const float2 z = v[br5(r)];
      float2 p;
      p.x = __shfl_sync(MST_FULL_MASK, v[br5(31 - r)].x, src);
      p.y = __shfl_sync(MST_FULL_MASK, v[br5(31 - r)].y, src);
      if (lane == 0) p = v[br5((32 - r) & 31)];  // lane 0 pairs with itself: bin 32*r <-> bin 32*(32-r); r = 0 -> Z[0]
      pair_butterfly<false>(z, p, w[i], o[r], o[31 - r]);
    }
  }
  {
    const float2 z = v[br5(16)];  // lane 0: Z[512], its own partner
    float2 dummy;
    pair_butterfly<false>(z, z, twp[512], *mid, dummy);
  }
}

__device__ __forceinline__ void rfft2048_warp(float2 (&v)[32], float2 (&o)[32], float2* mid, float2* scratch,
                                              const float2* tw1024, const float2* twp, int lane) {
  fft1024_warp<-1>(v, scratch, tw1024, lane);
  rfft_split(v, o, mid, twp, lane);
}

// Inverse real FFT from a mirror-layout spectrum y[0..31] (+ bin 512 in `mid` on lane 0).  Imaginary parts of the DC
// and Nyquist bins are ignored (pocketfft c2r semantics).  Returns z[m] = x[2m] + i*x[2m+1], m = lane + 32*r, UNSCALED
// by 1/1024, in v[br5(r)].
__device__ __forceinline__ void irfft2048_warp(float2 (&y)[32], float2 mid, float2 (&v)[32], float2* scratch,
                                               const float2* tw1024, const float2* twp, int lane) {
  const int src = (32 - lane) & 31;
  float2 t[16];  // merged values for bins 1024-k, to be handed to the mirrored lane
#pragma unroll
  for (int g = 0; g < 16; g += MST_LDS_BATCH) {
    float2 w[MST_LDS_BATCH];
#pragma unroll
    for (int i = 0; i < MST_LDS_BATCH; ++i) w[i] = MST_LDS_BATCH > 1 ? lds64(twp + lane + 32 * (g + i)) : twp[lane + 32 * (g + i)];
#pragma unroll
    for (int i = 0; i < MST_LDS_BATCH; ++i) {
      const int r = g + i;
      float2 a = y[r], b = y[31 - r];
      if (r == 0 && lane == 0) { a.y = 0.0f; b.y = 0.0f; }  // DC / Nyquist
      pair_butterfly<true>(a, b, w[i], v[r], t[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    float2 got;
    got.x = __shfl_sync(MST_FULL_MASK, t[r].x, src);
    got.y = __shfl_sync(MST_FULL_MASK, t[r].y, src);
    // lane 0 keeps its own: bin 32*(32-q) is element 32*(32-q) -> register 32-q, i.e. register 31-r takes q = r+1
    if (lane == 0) got = (r < 15) ? t[r + 1] : make_float2(mid.x, -mid.y);  // register 16 = Z[512] = conj(Y[512])
    v[31 - r] = got;
  }
  fft1024_warp<+1>(v, scratch, tw1024, lane);
}

// Fast, accurate log1p for x >= 0: 2*atanh(x/(2+x)) series below 0.25, hardware log2 above.
__device__ __forceinline__ float fast_log1p(float x) {
  const float s = __fdividef(x, 2.0f + x);
  const float s2 = s * s;
  float poly = fmaf(s2, 1.0f / 9.0f, 1.0f / 7.0f);
  poly = fmaf(s2, poly, 0.2f);
  poly = fmaf(s2, poly, 1.0f / 3.0f);
  poly = fmaf(s2, poly, 1.0f);
  const float small = 2.0f * s * poly;
  const float big = __log2f(1.0f + x) * 0.69314718055994530942f;
  return x < 0.25f ? small : big;
}

// The same for two values at once with packed fp32x2 arithmetic (the two special-function lookups per value stay
// scalar): 9.5 issue slots per value instead of 16; the same formula per component.
__device__ __forceinline__ float2 fast_log1p2(float2 x) {
  const float2 den = pk_add(x, pk_bcast(2.0f));
  const float2 s = pk_mul(x, make_float2(__fdividef(1.0f, den.x), __fdividef(1.0f, den.y)));
  const float2 s2 = pk_mul(s, s);
  float2 poly = pk_fma(s2, pk_bcast(1.0f / 9.0f), pk_bcast(1.0f / 7.0f));
  poly = pk_fma(s2, poly, pk_bcast(0.2f));
  poly = pk_fma(s2, poly, pk_bcast(1.0f / 3.0f));
  poly = pk_fma(s2, poly, pk_bcast(1.0f));
  const float2 small = pk_mul(pk_mul(pk_bcast(2.0f), s), poly);
  const float2 u = pk_add(x, pk_bcast(1.0f));
  const float2 big = pk_mul(make_float2(__log2f(u.x), __log2f(u.y)), pk_bcast(0.69314718055994530942f));
  return make_float2(x.x < 0.25f ? small.x : big.x, x.y < 0.25f ? small.y : big.y);
}

// Cooperative copy of a 16-byte-aligned table into shared memory (n_vec4 float4 elements).
__device__ __forceinline__ void stage_table(void* dst, const void* src, int n_vec4) {
  const float4* a = reinterpret_cast<const float4*>(src);
  float4* d = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < n_vec4; i += blockDim.x) d[i] = __ldg(a + i);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 16-byte asynchronous global -> shared copy (LDGSTS), bypassing registers and L1.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

}  // namespace mst
